"""Shared helpers of the parity tests: the oracle is the checker, libinb200 the thing checked."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import glow_oracle as O  # noqa: E402

# float32 parity tolerances (north_star: "within a stated float32 tolerance, e.g. 1e-4 relative").
# Every comparison is ||a - b||_2 / ||b||_2 against the oracle evaluated in float64 on the same
# float32 inputs and weights ("truth"), beside the error of the float32 oracle itself.
TOL_OUT = 1e-4      # outputs Z, X, dX
TOL_GRAD = 1e-4     # parameter gradients
TOL_LOGDET = 1e-5   # |dlogdet| / |logdet|


def rel(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    den = torch.linalg.norm(b).item()
    if den == 0.0:
        return torch.linalg.norm(a).item()
    return (torch.linalg.norm(a - b) / den).item()


def to64(net_params):
    return [None if p.data is None else p.data.double() for p in net_params]


def copy_oracle_to(net, oracle_net, device="cuda"):
    """Load the oracle's float32 parameters (get_params order) into an inb200 object."""
    import inb200
    inb200.set_params(net, [None if p.data is None else p.data.float() for p in oracle_net.get_params()])


def clone_oracle(make, dtype):
    """Build the same seeded oracle network in another dtype (weights are generated in float64 and
    rounded to float32 first, so both copies hold identical float32 values)."""
    n32 = make(torch.float32)
    n = make(dtype)
    for p, q in zip(n.get_params(), n32.get_params()):
        if q.data is not None:
            p.data = q.data.to(dtype)
    return n
