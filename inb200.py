"""Loader: `import inb200` -> the package in ./invertiblenetworks.jl_b200/ (a directory name with a
dot cannot be imported by name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "invertiblenetworks.jl_b200")
_spec = importlib.util.spec_from_file_location("inb200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["inb200"] = _mod
_spec.loader.exec_module(_mod)
