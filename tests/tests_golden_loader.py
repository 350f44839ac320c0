"""Loads tests/golden/reference_api.json (the reference's struct and method definitions, extracted by
tests/golden/make_reference_api.py)."""
import json
import os


def load_reference_api():
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "reference_api.json")) as fh:
        return json.load(fh)
