"""CPU-side checks of the drop-in boundary: libinb200.so loads without a GPU, exports every symbol
include/inb200.h declares, the ctypes table mirrors the header, and the plan bookkeeping
(get_params order / sizes, Z_dims, errors) matches the oracle's restatement of the reference."""
import ctypes
import os
import re

import pytest
import torch

from _util import O, ROOT

import inb200
from inb200 import lib as L

HEADER = os.path.join(ROOT, "include", "inb200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|long long|const char\*)\s+(inb_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("void", "") else len([a for a in args.split(",")])
        out[m.group(1)] = n
    return out


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.load()
    fns = header_functions()
    assert len(fns) >= 30
    for name in fns:
        assert hasattr(lib, name), f"{name} declared in inb200.h but not exported"
    assert lib.inb_version() >= 100


def test_ctypes_table_mirrors_header():
    fns = header_functions()
    assert set(fns) == set(L.SIGNATURES), set(fns) ^ set(L.SIGNATURES)
    for name, n in fns.items():
        assert len(L.SIGNATURES[name][1]) == n, name


def make_plan(**kw):
    base = dict(ndims=2, nx=16, ny=16, nz=1, n_in=2, n_cond=0, n_hidden=4, L=2, K=2, batch=2, split_scales=1,
                logdet=1, k1=3, k2=1, p1=1, p2=0, sig_low=0.0, sig_high=1.0, freeze_conv=0, precision=0)
    base.update(kw)
    d = L.GlowDesc(*[base[f[0]] for f in L.GlowDesc._fields_])
    plan = L.P()
    L.call("inb_glow_plan_create", ctypes.byref(d), ctypes.byref(plan))
    return plan


def plan_numels(plan):
    lib = L.load()
    n = ctypes.c_longlong()
    out = []
    for i in range(lib.inb_glow_num_params(plan)):
        assert lib.inb_glow_param_numel(plan, i, ctypes.byref(n)) == 0
        out.append(n.value)
    return out


@pytest.mark.parametrize("cfg", [
    dict(n_in=1, n_hidden=32, L=2, K=2, nx=64, ny=64),                       # BASELINE cfg1
    dict(n_in=3, n_hidden=256, L=3, K=16, nx=256, ny=256, batch=64),          # BASELINE cfg2
    dict(n_in=2, n_hidden=4, L=2, K=2, split_scales=0),
    dict(n_in=2, n_hidden=4, L=2, K=2, ndims=3, nx=8, ny=8, nz=8),
    dict(n_in=3, n_hidden=8, L=1, K=3, split_scales=1),
])
def test_param_order_and_sizes_match_reference_get_params(cfg):
    plan = make_plan(**cfg)
    G = O.NetworkGlow(cfg["n_in"], cfg["n_hidden"], cfg["L"], cfg["K"], split_scales=bool(cfg.get("split_scales", 1)),
                      ndims=cfg.get("ndims", 2))
    # ActNorm parameters are unset until the first forward: take their sizes from k
    want = []
    for row in G.AN:
        for an in row:
            want += [an.k, an.k]
    for row in G.CL:
        for cl in row:
            want += [p.data.numel() for p in cl.params()]
    assert plan_numels(plan) == want
    assert len(want) == 10 * cfg["L"] * cfg["K"]  # test_glow.jl:58-59
    L.call("inb_glow_plan_destroy", plan)


def test_param_order_conditional():
    plan = make_plan(n_in=1, n_cond=1, n_hidden=8, L=2, K=3, nx=16, ny=16)
    G = O.NetworkConditionalGlow(1, 1, 8, 2, 3, split_scales=True)
    want = []
    for row in G.AN:
        for an in row:
            want += [an.k, an.k]
    want += [1, 1]
    for row in G.CL:
        for cl in row:
            want += [p.data.numel() for p in cl.params()]
    assert plan_numels(plan) == want
    assert len(want) == 10 * 2 * 3 + 2  # test_conditional_glow_network.jl:46
    L.call("inb_glow_plan_destroy", plan)


@pytest.mark.parametrize("nd", [2, 3])
def test_zdims_match_reference(nd):
    sp = (16, 16) if nd == 2 else (8, 8, 8)
    G = O.NetworkGlow(2, 4, 3, 1, split_scales=True, ndims=nd)
    G.forward(torch.rand(2, 2, *sp))
    plan = make_plan(n_in=2, L=3, K=1, ndims=nd, nx=sp[-1], ny=sp[-2], nz=sp[0] if nd == 3 else 1)
    dims = (ctypes.c_int * 5)()
    for i, zd in enumerate(G.Z_dims):
        n = L.load().inb_glow_zdims(plan, 2, i, dims)
        assert tuple(dims[j] for j in range(n)) == tuple(zd)
    L.call("inb_glow_plan_destroy", plan)


def test_errors_are_reported_not_thrown():
    with pytest.raises(L.InbError, match="multiple of 2"):      # dimensionality_operations.jl:82-84
        make_plan(nx=15, ny=16)
    with pytest.raises(L.InbError, match="kernel sizes"):
        make_plan(k1=5, p1=2)
    with pytest.raises(L.InbError, match="ndims"):
        make_plan(ndims=4)
    with pytest.raises(L.InbError, match="precision"):
        make_plan(precision=7)


def test_workspace_is_sized_without_a_gpu():
    plan = make_plan(n_in=3, n_hidden=256, L=3, K=16, nx=256, ny=256, batch=8)
    gb = L.load().inb_glow_workspace_bytes(plan) / 2 ** 30
    assert 0.1 < gb < 8
    L.call("inb_glow_plan_destroy", plan)


def test_host_mirror_refuses_cpu_tensors():
    with pytest.raises(L.InbError, match="CUDA tensors only"):
        inb200.squeeze(torch.rand(1, 1, 4, 4))
    G = inb200.NetworkGlow(2, 4, 2, 2, split_scales=True, device="cpu")
    assert len(G.get_params()) == 40
    with pytest.raises(L.InbError, match="CUDA tensors only"):
        G.forward(torch.rand(2, 2, 16, 16))


def test_host_mirror_param_shapes_match_oracle():
    G = inb200.NetworkGlow(3, 8, 2, 2, split_scales=True, device="cpu")
    Og = O.NetworkGlow(3, 8, 2, 2, split_scales=True)
    for p, q in zip(G.get_params()[8:], Og.get_params()[8:]):
        assert tuple(p.data.shape) == tuple(q.data.shape)
    Gc = inb200.NetworkConditionalGlow(1, 2, 8, 2, 2, split_scales=True, device="cpu")
    Oc = O.NetworkConditionalGlow(1, 2, 8, 2, 2, split_scales=True)
    for p, q in zip(Gc.get_params()[10:], Oc.get_params()[10:]):
        assert tuple(p.data.shape) == tuple(q.data.shape)


def test_product_path_never_touches_the_oracle_and_has_no_fallback():
    """The oracle is test infrastructure: nothing under invertiblenetworks.jl_b200/ may import, call or mention it, and
    a missing libinb200.so is a hard error (no CPU / PyTorch fallback)."""
    pkg = os.path.join(ROOT, "invertiblenetworks.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "/root/reference" not in src, f"{f} reads the reference tree"
                if f.endswith(".py"):
                    assert not re.search(r"^\s*(from|import)\s+\.*oracle|oracle\.|import_module\([^)]*oracle", src, flags=re.M), \
                        f"{f} imports or calls the oracle"
    saved, L._lib = L._lib, None
    real = L.LIB_PATH
    try:
        L.LIB_PATH = os.path.join(pkg, "no_such_library.so")
        with pytest.raises(L.InbError, match="no CPU or PyTorch fallback"):
            L.load()
    finally:
        L.LIB_PATH, L._lib = real, saved


def test_hint_host_mirror_refuses_cpu_tensors():
    H = inb200.CouplingLayerHINT(8, 4, device="cpu")
    with pytest.raises(L.InbError, match="CUDA tensors only"):
        H.forward(torch.randn(1, 8, 4, 4))
    with pytest.raises(L.InbError, match="CUDA tensors only"):
        inb200.wavelet_squeeze(torch.randn(1, 1, 4, 4))
