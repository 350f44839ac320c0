"""Extracts the part of the reference's Julia API the shim (julia/InvertibleNetworksB200.jl) binds to: every struct
definition (field names and declared types) and every method definition of the overloaded generic functions
(positional argument types, keyword names, file:line).  Run in the build container, where /root/reference exists:

    python tests/golden/make_reference_api.py            # rewrites tests/golden/reference_api.json

tests/test_julia_shim.py checks the shim against the committed JSON (and, when /root/reference is present, that the
JSON is up to date).  No reference source text is copied: only names, types and line numbers."""
import json
import os
import re
import sys

REF = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_api.json")
GENERICS = ["forward", "inverse", "backward", "squeeze", "unsqueeze", "wavelet_squeeze", "wavelet_unsqueeze",
            "Haar_squeeze", "invHaar_unsqueeze", "get_params", "clear_grad!", "set_params!"]


def split_top(s, sep=","):
    """split on `sep` at bracket depth 0"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def parse_signature(text):
    """`name(args; kwargs)` -> (positional type strings, keyword names)"""
    depth, i0 = 0, text.index("(")
    for i in range(i0, len(text)):
        if text[i] == "(":
            depth += 1
        elif text[i] == ")":
            depth -= 1
            if depth == 0:
                inner = text[i0 + 1:i]
                break
    else:
        raise ValueError(text)
    pos, kw = inner, ""
    # the first ';' at depth 0 separates keywords
    d = 0
    for j, ch in enumerate(inner):
        if ch in "([{":
            d += 1
        elif ch in ")]}":
            d -= 1
        elif ch == ";" and d == 0:
            pos, kw = inner[:j], inner[j + 1:]
            break
    types = []
    for a in split_top(pos):
        if not a:
            continue
        a = a.split("=")[0].strip()
        types.append(a.split("::", 1)[1].strip() if "::" in a else "Any")
    kws = []
    for a in split_top(kw):
        if a:
            kws.append(re.split(r"::|=", a)[0].strip())
    return types, kws


def scan(ref=REF):
    structs, methods = {}, {g: [] for g in GENERICS}
    for root, _, files in os.walk(ref):
        for fn in sorted(files):
            if not fn.endswith(".jl"):
                continue
            path = os.path.join(root, fn)
            rel = os.path.relpath(path, os.path.dirname(ref))
            lines = open(path, encoding="utf-8").read().split("\n")
            i = 0
            while i < len(lines):
                ln = lines[i]
                m = re.match(r"^(?:mutable\s+)?struct\s+(\w+)", ln)
                if m:
                    name, fields, j = m.group(1), [], i + 1
                    while j < len(lines) and not re.match(r"^end\b", lines[j]):
                        f = lines[j].split("#")[0].strip()
                        if f and re.match(r"^\w+(::.*)?$", f):
                            parts = f.split("::", 1)
                            fields.append([parts[0], parts[1].strip() if len(parts) > 1 else "Any"])
                        j += 1
                    structs[name] = {"fields": fields, "where": f"{rel}:{i + 1}"}
                    i = j
                    continue
                m = re.match(r"^function\s+([\w!]+)\s*\(", ln)
                if m and m.group(1) in methods:
                    sig = ln
                    j = i
                    while sig.count("(") > sig.count(")") and j + 1 < len(lines):
                        j += 1
                        sig += " " + lines[j].strip()
                    types, kws = parse_signature(sig[sig.index(m.group(1)):])
                    methods[m.group(1)].append({"args": types, "kwargs": kws, "where": f"{rel}:{i + 1}"})
                i += 1
    return {"structs": structs, "methods": methods}


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("the reference tree is not present")
    api = scan()
    with open(OUT, "w") as fh:
        json.dump(api, fh, indent=1, sort_keys=True)
    print(f"{len(api['structs'])} structs, {sum(len(v) for v in api['methods'].values())} methods -> {OUT}")
