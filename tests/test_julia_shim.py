"""Static checks of julia/InvertibleNetworksB200.jl - the reference-side binding, which this image cannot execute
(no Julia toolchain).  Three things are verified on every CPU test run:

  1. every `obj.field` chain rooted at a typed argument (or a local derived from one) resolves against the reference's
     struct definitions (tests/golden/reference_api.json, extracted from /root/reference/src by
     tests/golden/make_reference_api.py) - the class of bug the round-1 shim had (`act.low` on an ActivationFunction,
     whose fields are forward / inverse / backward only);
  2. every overloaded method (forward / inverse / backward / squeeze / ...) has a reference method of the same name,
     arity, object type and keyword set, and its `invoke` fall-through names that method's signature;
  3. every `ccall` names a symbol declared in include/inb200.h with the right number of arguments and compatible
     argument types, and the GlowDesc / HintDesc mirrors match the C structs field by field.
"""
import json
import os
import re

import pytest

from tests_golden_loader import load_reference_api  # noqa: E402  (tests/tests_golden_loader.py)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "julia", "InvertibleNetworksB200.jl")
HEADER = os.path.join(ROOT, "include", "inb200.h")
GENERICS = {"forward", "inverse", "backward", "squeeze", "unsqueeze", "wavelet_squeeze", "wavelet_unsqueeze",
            "Haar_squeeze", "invHaar_unsqueeze"}
ALIASES = {"GlowLayer": ["CouplingLayerGlow", "ConditionalLayerGlow"]}


def strip_comments(src):
    return "\n".join(re.sub(r"#.*$", "", ln) if '"' not in ln.split("#")[0][-1:] else ln for ln in
                     (re.sub(r'(^|[^"])#[^"\n]*$', r"\1", l) for l in src.split("\n")))


def split_top(s, sep=","):
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


def balanced(text, i0, open_ch="(", close_ch=")"):
    """index just past the bracket that closes the one at i0"""
    depth = 0
    for i in range(i0, len(text)):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced")


def parse_args(inner):
    pos, kw, d = inner, "", 0
    for j, ch in enumerate(inner):
        if ch in "([{":
            d += 1
        elif ch in ")]}":
            d -= 1
        elif ch == ";" and d == 0:
            pos, kw = inner[:j], inner[j + 1:]
            break
    args = []
    for a in split_top(pos):
        if a:
            a = a.split("=")[0].strip()
            nm, ty = (a.split("::", 1) + ["Any"])[:2] if "::" in a else (a, "Any")
            args.append((nm.strip(), ty.strip()))
    kws = [re.split(r"::|=", a)[0].strip() for a in split_top(kw) if a]
    return args, kws


def shim_functions(src):
    """[(name, [(arg, type)], [kw], body, line)] for every `function name(...) ... end` at top level"""
    out = []
    for m in re.finditer(r"^(?:    @eval )?function\s+(\$?[\w!]+)\s*\(", src, re.M):
        i0 = m.end() - 1
        i1 = balanced(src, i0)
        args, kws = parse_args(src[i0 + 1:i1 - 1])
        # body: up to the matching top-level `end` (functions here start at column 0 or inside an @eval loop)
        indent = len(m.group(0)) - len(m.group(0).lstrip())
        tail = src[i1:]
        e = re.search(r"^" + " " * indent + r"end\b", tail, re.M)
        body = tail[:e.start()] if e else tail
        out.append((m.group(1), args, kws, body, src[:m.start()].count("\n") + 1))
    return out


def base_types(ty):
    """Julia type annotation -> list of candidate struct names"""
    ty = ty.strip()
    if ty in ALIASES:
        return ALIASES[ty]
    m = re.match(r"^Union\{(.*)\}$", ty)
    if m:
        out = []
        for t in split_top(m.group(1)):
            out += base_types(t)
        return out
    return [re.sub(r"\{.*$", "", ty)]


def elem_type(ty):
    m = re.match(r"^(?:Abstract)?Array\{\s*([\w.]+)\s*,", ty.strip())
    return m.group(1) if m else None


def resolve_chain(api, types, chain):
    """walk `.field` / `[...]` steps from the candidate struct types; returns (ok, message, final types)"""
    cur = list(types)
    cur_decl = None
    for step in chain:
        if step.startswith("["):
            if cur_decl is not None and elem_type(cur_decl):
                cur, cur_decl = [elem_type(cur_decl)], None
                continue
            return True, "", []  # indexing something we do not model (tuples, arrays of numbers)
        known = [t for t in cur if t in api["structs"]]
        if not known:
            return True, "", []  # Function / Bool / untyped field: nothing to check below it
        nxt, decl = [], None
        for t in known:
            for f, fty in api["structs"][t]["fields"]:
                if f == step:
                    decl = fty
                    nxt += base_types(fty) if not elem_type(fty) else []
        if decl is None:
            return False, f"`{step}` is not a field of {' / '.join(known)} " \
                          f"(fields: {[f for f, _ in api['structs'][known[0]]['fields']]})", []
        cur, cur_decl = nxt, decl
    return True, "", cur


CHAIN = re.compile(r"(?<![\w.\]\)])([A-Za-zΔ∇θ_][\wΔ∇θ]*)((?:\.[A-Za-z_]\w*|\[[^\[\]]*\])+)")


@pytest.fixture(scope="module")
def api():
    return load_reference_api()


@pytest.fixture(scope="module")
def shim():
    return strip_comments(open(SHIM, encoding="utf-8").read())


def test_reference_api_fixture_is_current(api):
    """When the reference tree is available (the build container) the committed fixture must equal a fresh extraction."""
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not present")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "tests", "golden", "make_reference_api.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    assert mk.scan() == api


def test_every_field_access_exists_in_the_reference_structs(api, shim):
    checked, problems = 0, []
    for name, args, kws, body, line in shim_functions(shim):
        env = {}
        for a, ty in args:
            bt = [t for t in base_types(ty) if t in api["structs"]]
            if bt:
                env[a] = bt
        # locals bound to a chain (`rb = L1.RB`, `L1 = G.CL[1, 1]`) and lambda / loop variables over typed arrays
        for _ in range(3):
            for m in re.finditer(r"(?:^|[\s(;])([A-Za-z_]\w*)\s*=\s*([A-Za-z_]\w*)((?:\.[A-Za-z_]\w*|\[[^\[\]]*\])+)\s*(?:$|[;\n])",
                                 body, re.M):
                var, root, ch = m.group(1), m.group(2), m.group(3)
                if root in env:
                    ok, _, fin = resolve_chain(api, env[root], re.findall(r"\.(\w+)|(\[[^\[\]]*\])", ch) and
                                               [a or b for a, b in re.findall(r"\.(\w+)|(\[[^\[\]]*\])", ch)])
                    if ok and fin and all(t in api["structs"] for t in fin):
                        env[var] = fin
            for m in re.finditer(r"\b([A-Za-z_]\w*)\s*->[^,\n]*?,\s*([A-Za-z_]\w*)\.(\w+)\)", body):  # f(x -> ..., G.CL)
                var, root, fld = m.groups()
                if root in env:
                    for t in env[root]:
                        for f, fty in api["structs"].get(t, {"fields": []})["fields"]:
                            if f == fld and elem_type(fty):
                                env[var] = [elem_type(fty)]
            for m in re.finditer(r"\bfor\s+([A-Za-z_]\w*)\s+in\s+([A-Za-z_]\w*)\.(\w+)", body):  # for L in G.CL
                var, root, fld = m.groups()
                if root in env:
                    for t in env[root]:
                        for f, fty in api["structs"].get(t, {"fields": []})["fields"]:
                            if f == fld and elem_type(fty):
                                env[var] = [elem_type(fty)]
        for m in CHAIN.finditer(body):
            root, ch = m.group(1), m.group(2)
            if root not in env:
                continue
            steps = [a or b for a, b in re.findall(r"\.(\w+)|(\[[^\[\]]*\])", ch)]
            ok, msg, _ = resolve_chain(api, env[root], steps)
            checked += 1
            if not ok:
                problems.append(f"{name} (shim line {line}): {root}{ch}: {msg}")
    assert not problems, "\n".join(problems)
    assert checked > 100, f"only {checked} field accesses were resolved - the checker lost track of the shim"


def test_known_bad_accesses_are_caught(api):
    """The checker itself: the round-1 bugs must be flagged."""
    ok, msg, _ = resolve_chain(api, ["CouplingLayerGlow"], ["activation", "low"])
    assert not ok and "low" in msg
    ok, _, _ = resolve_chain(api, ["NetworkGlow"], ["CL", "[1, 1]", "RB", "W1", "data"])
    assert ok
    ok, msg, _ = resolve_chain(api, ["NetworkMultiScaleHINT"], ["Z_dims"])
    assert not ok
    ok, msg, _ = resolve_chain(api, ["NetworkConditionalGlow"], ["logdet"])
    assert not ok


def test_every_overload_matches_a_reference_method(api, shim):
    seen = 0
    for name, args, kws, body, line in shim_functions(shim):
        names = [name]
        if name == "$fn":  # @eval loop over (forward, inverse) of Conv1x1
            names = ["forward", "inverse"]
        for nm in names:
            if nm not in GENERICS:
                continue
            seen += 1
            obj = base_types(args[-1][1])
            cands = [m for m in api["methods"][nm] if len(m["args"]) == len(args)]
            match = []
            for m in cands:
                mobj = base_types(m["args"][-1])
                if set(obj) & set(mobj) or (not set(obj) & set(api["structs"]) and mobj[0] in ("AbstractArray", "Any")):
                    # tuple-taking methods: the reference has forward/inverse(::Tuple, ::Conv1x1)
                    if all(("Tuple" in a[1]) == ("Tuple" in b) for a, b in zip(args[:-1], m["args"][:-1])):
                        match.append(m)
            assert match, f"{nm}{[a[1] for a in args]} (shim line {line}) has no reference method of that arity / object type"
            assert any(sorted(m["kwargs"]) == sorted(kws) for m in match), \
                f"{nm}(..., ::{args[-1][1]}) (shim line {line}): keywords {kws} vs reference {[m['kwargs'] for m in match]}"
            # the fall-through must name a signature of the same arity ending in the same object type
            for iv in re.finditer(r"invoke\(\s*" + re.escape(nm) + r"\s*,\s*Tuple\{", body):
                j = balanced(body, iv.end() - 1, "{", "}")
                tys = split_top(body[iv.end():j - 1])
                assert len(tys) == len(args), f"{nm} (shim line {line}): invoke signature has {len(tys)} types for {len(args)} arguments"
                assert set(base_types(tys[-1])) & set(obj) or not set(obj) & set(api["structs"]), \
                    f"{nm} (shim line {line}): invoke falls through to {tys[-1]}, the method is for {args[-1][1]}"
    assert seen >= 30, seen


# ---------------------------------------------------------------- ccall vs include/inb200.h
def c_prototypes():
    h = open(HEADER).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|long long|const char\*)\s+(inb_\w+)\s*\(([^;{]*)\)\s*;", h):
        params = [p.strip() for p in m.group(3).replace("\n", " ").split(",")]
        if params == ["void"]:
            params = []
        protos[m.group(2)] = (m.group(1), params)
    structs = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\}\s*\w+;", h, re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ty, names = decl.split(None, 1)
            fields += [(n.strip(), ty) for n in names.split(",")]
        structs[m.group(1)] = fields
    return protos, structs


def c_kind(param):
    p = re.sub(r"/\*.*?\*/", "", param).strip()
    p = re.sub(r"\[\d*\]$", "*", re.sub(r"\s+\w+(\[\d*\])?$", lambda m: m.group(1) or "", p)).strip() if not p.endswith("*") else p
    p = p.replace("const ", "").replace(" const", "").replace(" ", "")
    if p in ("int",):
        return "int"
    if p == "longlong":
        return "ll"
    if p == "float":
        return "float"
    if p == "float*":
        return "pf"
    if p == "float**":
        return "ppf"
    if p in ("int*",):
        return "pint"
    if p == "longlong*":
        return "pll"
    if p in ("char*",):
        return "pchar"
    if p.endswith("**"):
        return "pphandle"
    if p.endswith("*"):
        return "phandle"  # void*, inb_plan*, inb_comm*, const inb_glow_desc*, double*
    raise AssertionError(f"unclassified C parameter `{param}`")


JL_KIND = {"Cint": {"int"}, "Clonglong": {"ll"}, "Cfloat": {"float"}, "Ptr{Cfloat}": {"pf"}, "Ptr{Ptr{Cfloat}}": {"ppf"},
           "Ptr{Cint}": {"pint"}, "Ptr{Clonglong}": {"pll"}, "Ref{Clonglong}": {"pll"}, "Ptr{UInt8}": {"pchar"},
           "Ptr{Cvoid}": {"phandle"}, "Ref{Ptr{Cvoid}}": {"pphandle"}, "Ref{GlowDesc}": {"phandle"},
           "Ref{HintDesc}": {"phandle"}, "Cstring": {"pchar"}}


def test_every_ccall_matches_the_header(shim):
    protos, _ = c_prototypes()
    consts = {m.group(1): split_top(m.group(2)) for m in re.finditer(r"^const (\w+_ARGT) = \(([^)]*)\)", shim, re.M)}
    n = 0
    for m in re.finditer(r"ccall\(\(\s*(?::(\w+)|\$\(QuoteNode\(sym\)\))\s*,\s*LIB\)\s*,\s*(\w+)\s*,\s*\(", shim):
        sym = m.group(1)
        i0 = m.end() - 1
        i1 = balanced(shim, i0)
        tys = []
        for t in split_top(shim[i0 + 1:i1 - 1]):
            if t.endswith("..."):
                tys += consts[t[:-3]]
            elif t:
                tys.append(t)
        call_end = balanced(shim, shim.rfind("ccall(", 0, m.end()) + 5)
        nargs = len(split_top(shim[i1:call_end - 1].lstrip(", \n")))
        syms = [sym] if sym else ["inb_conv1x1_forward", "inb_conv1x1_inverse"]
        for s in syms:
            n += 1
            assert s in protos, f"ccall of `{s}`, which include/inb200.h does not declare"
            ret, params = protos[s]
            assert len(tys) == len(params), f"{s}: {len(tys)} Julia argument types for {len(params)} C parameters"
            for jt, cp in zip(tys, params):
                assert jt in JL_KIND, f"{s}: unclassified Julia type {jt}"
                assert c_kind(cp) in JL_KIND[jt], f"{s}: Julia {jt} passed for C `{cp}`"
            assert {"int": "Cint", "const char*": "Cstring", "long long": "Clonglong"}[ret] == m.group(2)
        # splatted argument tuples (geom(X)..., rb_ints(...)...) cannot be counted statically
        if "..." not in shim[i1:call_end]:
            assert nargs == len(tys), f"{syms[0]}: {nargs} arguments for {len(tys)} declared types"
    assert n >= 40, n


def test_struct_mirrors_match_the_header(shim):
    _, cstructs = c_prototypes()
    for jl, c in (("GlowDesc", "inb_glow_desc"), ("HintDesc", "inb_hint_desc")):
        m = re.search(r"^struct " + jl + r"\n(.*?)^end", shim, re.S | re.M)
        fields = []
        for decl in re.split(r"[;\n]", m.group(1)):
            decl = decl.strip()
            if decl:
                nm, ty = decl.split("::")
                fields.append((nm, {"Cint": "int", "Cfloat": "float"}[ty]))
        assert fields == cstructs[c], f"{jl} does not mirror {c}"


def test_integration_md_lists_only_bound_symbols(shim):
    """INTEGRATION.md's table rows name C symbols; each must be declared in the header and ccall'ed by the shim."""
    protos, _ = c_prototypes()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    named = set(re.findall(r"`(inb_\w+)`", doc))
    called = set(re.findall(r"ccall\(\(:(inb_\w+)", shim)) | {"inb_conv1x1_forward", "inb_conv1x1_inverse"}
    assert named, "INTEGRATION.md names no symbols"
    for s in sorted(named):
        if s.endswith("_"):
            continue  # a family prefix such as `inb_prof_`
        assert s in protos, f"INTEGRATION.md names `{s}`, which the header does not declare"
    bound_claims = set(re.findall(r"^\|[^|]*\|[^|]*`(inb_\w+)`[^|]*\|\s*yes", doc, re.M))
    for s in sorted(bound_claims):
        assert s in called, f"INTEGRATION.md says the shim binds `{s}`, but no ccall names it"
