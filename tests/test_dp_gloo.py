"""Data-parallel host logic on CPU: two gloo ranks, each with half of the batch, must reproduce the
single-process gradient of the global batch after ONE averaged all-reduce of the flat gradient buffer
(SURVEY 8e).  The compute on each rank is the oracle (this is a test: the oracle is the checker's stand-in
for the per-GPU library call); what is under test is invertiblenetworks.jl_b200/dp.py."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_dp():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import inb200  # the shared library is only dlopen'ed by a compute call; none is made here
    return inb200.dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _net():
    sys.path.insert(0, ROOT)
    from oracle import glow_oracle as O
    return O, O.NetworkGlow(2, 8, 2, 2, split_scales=True, seed=5, faithful=False, dtype=torch.float64)


def _worker(rank, world, port, X, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dp = _load_dp()
        O, G = _net()
        torch.set_num_threads(1)
        lo, hi = dp.shard_bounds(X.shape[0], rank, world)
        # rank 0 runs the data-dependent ActNorm init on ITS shard, everyone receives its parameters
        if rank == 0:
            G.forward(X[lo:hi])
        params = [p for p in G.get_params()]
        if rank != 0:  # allocate s, b (None before the first forward) so the broadcast has a destination
            G.forward(X[lo:hi])
        flat = dp.flatten([p.data for p in params])
        dp.broadcast_params(flat, src=0)
        dp.unflatten_into(flat, [p.data for p in params])
        f, _ = O.glow_train_step(G, X[lo:hi])
        g = dp.flatten([p.grad for p in params])
        dp.allreduce_grads(g)
        fm = dp.allreduce_mean_scalar(f.reshape(1))
        if rank == 0:
            out["flat_params"] = flat
            out["grads"] = g
            out["f"] = fm
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gradient_matches_global_batch():
    dp = _load_dp()
    torch.manual_seed(0)
    X = torch.rand(4, 2, 8, 8, dtype=torch.float64)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), X, out), nprocs=2, join=True)
    # single process, global batch, same parameters (rank 0's init)
    O, G = _net()
    G.forward(X[:2])  # allocates s, b
    dp.unflatten_into(out["flat_params"], [p.data for p in G.get_params()])
    f, _ = O.glow_train_step(G, X)
    g = dp.flatten([p.grad for p in G.get_params()])
    assert torch.allclose(out["grads"], g, rtol=1e-9, atol=1e-11), (out["grads"] - g).abs().max()
    assert torch.allclose(out["f"], f.reshape(1), rtol=1e-10)


def test_shard_bounds():
    dp = _load_dp()
    assert [dp.shard_bounds(64, r, 8) for r in (0, 7)] == [(0, 8), (56, 64)]
    with pytest.raises(ValueError):
        dp.shard_bounds(10, 0, 4)


def test_flat_layout_of_the_library_is_the_mirrors_layout():
    """inb_glow_flat_layout (the canonical layout a caller keeps its gradients in to get one collective per contiguous
    range) equals the offsets the Python mirror gives its flat parameter / gradient buffers."""
    import ctypes
    import inb200
    L = inb200.lib
    for cond in (0, 2):
        d = L.GlowDesc(2, 32, 32, 1, 3, cond, 16, 3, 2, 2, 1, 1, 3, 1, 1, 0, 0.0, 1.0, 0, 0)
        plan = L.P()
        L.call("inb_glow_plan_create", ctypes.byref(d), ctypes.byref(plan))
        n = L.load().inb_glow_num_params(plan)
        offs = (ctypes.c_longlong * n)()
        tot = ctypes.c_longlong()
        L.call("inb_glow_flat_layout", plan, offs, ctypes.byref(tot))
        o, want = 0, []
        for i in range(n):
            want.append(o)
            m = ctypes.c_longlong()
            L.call("inb_glow_param_numel", plan, i, ctypes.byref(m))
            o += (m.value + 63) // 64 * 64
        assert list(offs) == want and tot.value == o
        L.call("inb_glow_plan_destroy", plan)
