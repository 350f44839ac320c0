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


class FragileUnits:
    """Counts ReLU units whose pre-activation is within `thr` (relative to the tensor's rms) of zero
    while the float64 oracle runs its backward.  At such a unit _relugrad (activation_functions.jl:84)
    is discontinuous: ANY arithmetic that is not bit-identical (the reference's own fp32 on another
    BLAS included) may take the other branch, and the gradient then differs by O(1) around that unit.
    Parity of gradients is therefore asserted strictly when no unit is fragile, and with the bound
    TOL_GRAD_FRAGILE when the oracle itself reports fragile units (assert_grad_close)."""

    def __init__(self, thr):
        self.thr, self.count, self.total = thr, 0, 0

    def __enter__(self):
        self._orig = O.relu_grad

        def counted(dy, x):
            rms = x.pow(2).mean().sqrt()
            self.count += int((x.abs() < self.thr * rms).sum())
            self.total += x.numel()
            return self._orig(dy, x)
        O.relu_grad = counted
        return self

    def __exit__(self, *a):
        O.relu_grad = self._orig


# gradient bound that applies ONLY when the float64 oracle reports fragile ReLU units for the arithmetic
# under test: a unit that takes the other branch of _relugrad perturbs the gradients it feeds by O(1)
# (a flipped unit behind the 1x1 conv reaches every hidden channel of its pixel), so the L2 error is then
# set by the number of flips, not by the arithmetic.  Strict, flip-free gradient parity of the same
# kernels is asserted by test_resblock_exact_lattice and by every float32-path test.
TOL_GRAD_FRAGILE = 2e-2


def assert_grad_close(a, b, tol, fragile, what=""):
    r = rel(a, b)
    if r < tol:
        return
    assert fragile is not None and fragile.count > 0, f"{what}: {r} >= {tol} and no fragile ReLU unit explains it"
    assert r < max(TOL_GRAD_FRAGILE, tol), f"{what}: rel {r} (fragile ReLU units: {fragile.count}/{fragile.total})"
