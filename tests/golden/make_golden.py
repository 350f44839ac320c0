"""Generates tests/golden/glow_small.npz and hint_small.npz: seeded inputs, parameters and float64 outputs of the ORACLE
(oracle/glow_oracle.py) for a small NetworkGlow and a small NetworkConditionalGlow.

These vectors are NOT produced by the Julia reference (no Julia in the image; the reference's tests hold no golden
vectors for this path) - the oracle stays "parity unpinned".  They pin the oracle itself: any later edit of
the restatement that changes its numbers fails tests/test_oracle.py::test_oracle_matches_golden, and the CUDA path is
checked against the same committed numbers in tests/test_gpu_networks.py::test_cuda_matches_golden.

    python tests/golden/make_golden.py        # rewrites glow_small.npz (deterministic)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import glow_oracle as O  # noqa: E402
from oracle import hint_oracle as H  # noqa: E402


def run_glow():
    torch.manual_seed(123)
    n_in, nh, L, K, shape = 2, 8, 2, 2, (3, 2, 16, 16)
    G = O.NetworkGlow(n_in, nh, L, K, split_scales=True, seed=5, dtype=torch.float32, faithful=False)
    X = torch.rand(*shape)
    G.forward(X)  # data-dependent ActNorm init in float32 fixes the parameters
    G64 = O.NetworkGlow(n_in, nh, L, K, split_scales=True, seed=5, dtype=torch.float64, faithful=False)
    for p, q in zip(G64.get_params(), G.get_params()):
        p.data = q.data.double()
    Z, ld = G64.forward(X.double())
    dZ = Z / shape[0]
    dX, Xr = G64.backward(dZ, Z)
    out = {"glow_X": X.numpy(), "glow_Z": Z.numpy(), "glow_logdet": np.float64(ld), "glow_dX": dX.numpy(),
           "glow_cfg": np.array([n_in, nh, L, K])}
    for i, p in enumerate(G64.get_params()):
        out[f"glow_p{i:03d}"] = p.data.float().numpy()
        out[f"glow_g{i:03d}"] = p.grad.numpy()
    return out


def run_hint():
    """hint_small.npz: NetworkMultiScaleHINT(2, 6, L=2, K=2; k2=1) of test_multiscale_hint_network.jl:10-18 at 16x16 with
    split_scales (so both the wavelet squeeze, the recursion at 8 and 16 channels and the latent split are in)."""
    torch.manual_seed(321)
    n_in, nh, L, K, shape = 2, 6, 2, 2, (3, 2, 16, 16)
    mk = lambda dt: H.NetworkMultiScaleHINT(n_in, nh, L, K, split_scales=True, k2=1, p2=0, seed=7, dtype=dt)
    N = mk(torch.float32)
    X = torch.randn(*shape)
    N.forward(X)
    N64 = mk(torch.float64)
    for p, q in zip(N64.get_params(), N.get_params()):
        p.data = q.data.double()
    Z, ld = N64.forward(X.double())
    dX, Xr = N64.backward(Z / shape[0], Z)
    out = {"hint_X": X.numpy(), "hint_Z": Z.numpy(), "hint_logdet": np.float64(ld), "hint_dX": dX.numpy(),
           "hint_cfg": np.array([n_in, nh, L, K])}
    for i, p in enumerate(N64.get_params()):
        out[f"hint_p{i:03d}"] = p.data.float().numpy()
        out[f"hint_g{i:03d}"] = p.grad.numpy()
    return out


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name, fn in (("glow_small.npz", run_glow), ("hint_small.npz", run_hint)):
        out = fn()
        np.savez_compressed(os.path.join(here, name), **out)
        print("wrote", name, len(out), "arrays,", sum(v.nbytes for v in out.values()), "bytes")


if __name__ == "__main__":
    main()
