"""CPU oracle for the HINT family (BASELINE configs[3]) of
slimgroup/InvertibleNetworks.jl: CouplingLayerBasic, the recursive CouplingLayerHINT,
NetworkMultiScaleHINT and the Haar / wavelet squeezes.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as glow_oracle.py: only tests/, smoke()
and bench.py's CPU baseline may import it; the product path never does).

PARITY UNPINNED, like glow_oracle.py (no Julia here, no golden vectors in the reference's tests).
Pinned by the properties the reference's own tests assert (test_coupling_layer_hint.jl:17-31,
test_multiscale_hint_network.jl:25-66, test_squeeze.jl:15-31) and by float64 autograd.

Conventions restated from un-vendored packages:
  * Wavelets.jl (compat 0.9 / 0.10, Project.toml:26) `dwt(x, wavelet(WT.db1), 1)` of a matrix: separable
    one-level Haar transform, approximation in the leading half of each axis, detail in the trailing
    half, detail coefficient = (x[2k] - x[2k+1]) / sqrt(2) (0-based).  The sign of the detail
    coefficients is the package's filter convention and cannot be checked here; the reference's own
    lifting implementation `HaarLift` (dimensionality_operations.jl:264-284, restated literally in
    `haar_squeeze`) uses the same sign, and both are orthonormal, so invertibility, logdet (= 0) and
    the adjoint property hold for either sign.

Two reference behaviours that a user will notice, both kept selectable:
  * CouplingLayerHINT.backward with set_grad=true overwrites the gradients of the coupling layers
    shared by the two recursive calls (`RB.W1.grad = ...`, layer_residual_block.jl:168-172, called
    twice per level from invertible_layer_hint.jl:208-210), so for n_in > 4 only the LAST visit's
    contribution survives, whereas the set_grad=false path sums them (`Δθa+Δθb`, :222).
    `shared_grads="sum"` (default, the true gradient, what autograd gives) or `"last"` (set_grad=true).
  * NetworkMultiScaleHINT with split_scales=true calls `split_states(H.X_dims, Z)`
    (invertible_network_hint_multiscale.jl:123,138-139) with the arguments swapped relative to the
    only method (dimensionality_operations.jl:480); the evident intent (identical to NetworkGlow) is
    restated here.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

from .glow_oracle import (ActNorm, Conv1x1, Parameter, ResidualBlock, Tensor, cat_states, glorot_uniform,
                          sigmoid, sigmoid_grad, tensor_cat, tensor_split)


# ----------------------------------------------------------------------------------------
# Haar squeezes (2-D)
# ----------------------------------------------------------------------------------------
def _quads(X: Tensor):
    """p[ix][iy] = X[2x'+ix, 2y'+iy] in the reference's (nx, ny) order; torch tensors are (.., ny, nx)."""
    return [[X[..., iy::2, ix::2] for iy in (0, 1)] for ix in (0, 1)]


def wavelet_squeeze(X: Tensor) -> Tensor:
    """dimensionality_operations.jl:199-216 with type = WT.db1: per (channel j, sample) one-level 2-D dwt, then the
    four quadrants as channels 4j + q in `patch` order q = qx + 2 qy (patch_inds :30-36): q=0 approximation,
    q=1 detail along x, q=2 detail along y, q=3 diagonal."""
    if X.dim() != 4:
        raise ValueError("2-D wavelet squeeze only")
    if X.shape[2] % 2 or X.shape[3] % 2:
        raise ValueError("Input dimensions must be multiple of 2")
    p = _quads(X)
    a = (p[0][0] + p[1][0] + p[0][1] + p[1][1]) / 2
    dx = (p[0][0] - p[1][0] + p[0][1] - p[1][1]) / 2
    dy = (p[0][0] + p[1][0] - p[0][1] - p[1][1]) / 2
    dd = (p[0][0] - p[1][0] - p[0][1] + p[1][1]) / 2
    return torch.stack((a, dx, dy, dd), dim=2).reshape(X.shape[0], 4 * X.shape[1], X.shape[2] // 2, X.shape[3] // 2)


def wavelet_unsqueeze(Y: Tensor) -> Tensor:
    """dimensionality_operations.jl:243-258 (the transform is orthonormal: inverse = transpose)."""
    B, C4, h, w = Y.shape
    if C4 % 4:
        raise ValueError("number of channels must be divisible by 4")
    Yq = Y.reshape(B, C4 // 4, 4, h, w)
    a, dx, dy, dd = Yq[:, :, 0], Yq[:, :, 1], Yq[:, :, 2], Yq[:, :, 3]
    X = Y.new_zeros(B, C4 // 4, 2 * h, 2 * w)
    X[..., 0::2, 0::2] = (a + dx + dy + dd) / 2
    X[..., 0::2, 1::2] = (a - dx + dy - dd) / 2  # ix = 1
    X[..., 1::2, 0::2] = (a + dx - dy - dd) / 2  # iy = 1
    X[..., 1::2, 1::2] = (a - dx - dy + dd) / 2
    return X


def _haar_lift(x: Tensor, dim: int):
    """HaarLift, dimensionality_operations.jl:264-284, literal; dim 1 = x (torch last axis), dim 2 = y."""
    ax = -1 if dim == 1 else -2
    n = x.shape[ax]
    H = x.index_select(ax, torch.arange(0, n, 2)).clone()  # 1:2:end
    L = x.index_select(ax, torch.arange(1, n, 2)).clone()  # 2:2:end
    H -= L
    L += H / 2.0
    H /= math.sqrt(2.0)
    L *= math.sqrt(2.0)
    return L, H


def haar_squeeze(X: Tensor) -> Tensor:
    """Haar_squeeze, dimensionality_operations.jl:318-331 (2-D): channels cat(a, v, h, d), each of C channels."""
    L, H = _haar_lift(X, 2)
    a, h = _haar_lift(L, 1)
    v, d = _haar_lift(H, 1)
    return torch.cat((a, v, h, d), dim=1)


def _inv_haar_lift(L: Tensor, H: Tensor, dim: int) -> Tensor:
    """invHaarLift, dimensionality_operations.jl:291-312."""
    ax = -1 if dim == 1 else -2
    H = H * math.sqrt(2.0)
    L = L / math.sqrt(2.0)
    L = L - H / 2.0
    H = H + L
    shape = list(L.shape)
    shape[ax] *= 2
    x = L.new_zeros(shape)
    if dim == 1:
        x[..., 0::2] = H
        x[..., 1::2] = L
    else:
        x[..., 0::2, :] = H
        x[..., 1::2, :] = L
    return x


def inv_haar_unsqueeze(Y: Tensor) -> Tensor:
    """invHaar_unsqueeze, dimensionality_operations.jl:354-371 (2-D)."""
    a, h = tensor_split(Y)
    (a, v), (h, d) = tensor_split(a), tensor_split(h)
    L = _inv_haar_lift(a, h, 1)
    H = _inv_haar_lift(v, d, 1)
    return _inv_haar_lift(L, H, 2)


# ----------------------------------------------------------------------------------------
# CouplingLayerBasic  (src/layers/invertible_layer_basic.jl:62-166)
# ----------------------------------------------------------------------------------------
def coupling_logdet_forward(S: Tensor) -> Tensor:
    return torch.sum(torch.log(torch.abs(S))) / S.shape[0]  # :211


def coupling_logdet_backward(S: Tensor) -> Tensor:
    return 1.0 / S / S.shape[0]  # :212


class CouplingLayerBasic:
    def __init__(self, RB: ResidualBlock, logdet: bool = False, low: float = 0.0, high: float = 1.0):
        self.RB, self.logdet, self.low, self.high = RB, logdet, low, high

    def params(self) -> List[Parameter]:
        return self.RB.params()

    def _st(self, X1: Tensor):
        logS, T = tensor_split(self.RB.forward(X1))  # :96
        return sigmoid(logS, self.low, self.high), T  # :97

    def forward(self, X1: Tensor, X2: Tensor):
        S, T = self._st(X1)
        Y2 = S * X2 + T  # :98
        return X1, Y2, (coupling_logdet_forward(S) if self.logdet else 0.0)

    def inverse(self, Y1: Tensor, Y2: Tensor):
        S, T = self._st(Y1)  # :112-113
        eps = torch.finfo(torch.float32).eps
        X2 = (Y2 - T) / (S + eps)  # :114, eps(Float32) also in the float64 "truth" run
        return Y1, X2, S

    def backward(self, dY1: Tensor, dY2: Tensor, Y1: Tensor, Y2: Tensor):
        X1, X2, S = self.inverse(Y1, Y2)  # :127
        dT = dY2.clone()  # :130
        dS = dY2 * X2  # :131
        if self.logdet:
            dS = dS - coupling_logdet_backward(S)  # :133
        dX2 = dY2 * S  # :135
        dX1 = self.RB.backward(tensor_cat(sigmoid_grad(dS, S, self.low, self.high), dT), X1) + dY1  # :137
        return dX1, dX2, X1, X2


# ----------------------------------------------------------------------------------------
# CouplingLayerHINT  (src/layers/invertible_layer_hint.jl:52-300)
# ----------------------------------------------------------------------------------------
def get_depth(n_in: int) -> int:
    """invertible_layer_hint.jl:63-71"""
    count, nc = 0, float(n_in)
    while nc > 4:
        nc /= 2
        count += 1
    return count + 1


class CouplingLayerHINT:
    def __init__(self, CL: Sequence[CouplingLayerBasic], C: Optional[Conv1x1], logdet: bool = False,
                 permute: str = "none", shared_grads: str = "sum"):
        assert permute in ("none", "full", "lower", "both")
        assert shared_grads in ("sum", "last")
        self.CL, self.C, self.logdet, self.permute, self.shared_grads = list(CL), C, logdet, permute, shared_grads

    def params(self) -> List[Parameter]:
        """get_params field order (neuralnet.jl:72-84): CL[1..n].RB.(W1,W2,W3,b1,b2) then C.(v1,v2,v3)."""
        out: List[Parameter] = []
        for cl in self.CL:
            out += cl.params()
        if self.C is not None:
            out += self.C.params()
        return out

    # :105-156
    def forward(self, X: Tensor, scale: int = 1, permute: Optional[str] = None):
        permute = self.permute if permute is None else permute
        if permute in ("full", "both"):  # :111-113
            X = self.C.forward(X)
        Xa, Xb = tensor_split(X)
        if permute == "lower":
            Xb = self.C.forward(Xb)
        cl = self.CL[scale - 1]
        if X.shape[1] > 4:  # :120-124
            Ya, ld1 = self.forward(Xa, scale + 1, "none")
            Yt, ld2 = self.forward(Xb, scale + 1, "none")
            _, Yb, ld3 = cl.forward(Xa, Yt)
            ld = ld1 + ld2 + ld3
        else:
            Ya = Xa.clone()
            _, Yb, ld = cl.forward(Xa, Xb)
        Y = tensor_cat(Ya, Yb)
        if permute == "both":  # :149
            Y = self.C.inverse(Y)
        return Y, ld

    # :159-204
    def inverse(self, Y: Tensor, scale: int = 1, permute: Optional[str] = None) -> Tensor:
        permute = self.permute if permute is None else permute
        if permute == "both":  # :164
            Y = self.C.forward(Y)
        Ya, Yb = tensor_split(Y)
        cl = self.CL[scale - 1]
        if Y.shape[1] > 4:
            Xa = self.inverse(Ya, scale + 1, "none")
            Yt = cl.inverse(Xa, Yb)[1]
            Xb = self.inverse(Yt, scale + 1, "none")
        else:
            Xa = Ya.clone()
            Xb = cl.inverse(Ya, Yb)[1]
        if permute == "lower":
            Xb = self.C.inverse(Xb)
        X = tensor_cat(Xa, Xb)
        if permute in ("full", "both"):  # :195-197
            X = self.C.inverse(X)
        return X

    def _cl_backward(self, cl, dY1, dY2, Y1, Y2, seen):
        """CL[scale].backward with the shared-gradient rule of the module docstring."""
        old = [p.grad for p in cl.params()] if (id(cl) in seen and self.shared_grads == "sum") else None
        out = cl.backward(dY1, dY2, Y1, Y2)
        if old is not None:
            for p, g in zip(cl.params(), old):
                p.grad = p.grad + g
        seen.add(id(cl))
        return out

    # :207-297 (set_grad = true)
    def backward(self, dY: Tensor, Y: Tensor, scale: int = 1, permute: Optional[str] = None, seen=None):
        permute = self.permute if permute is None else permute
        seen = set() if seen is None else seen
        if permute == "both":  # :219-221
            dY, Y = self.C.forward_tuple(dY, Y, faithful_batch_loop=False)
        Ya, Yb = tensor_split(Y)
        dYa, dYb = tensor_split(dY)
        cl = self.CL[scale - 1]
        if Y.shape[1] > 4:
            dXa, Xa = self.backward(dYa, Ya, scale + 1, "none", seen)  # :230
            dXa_t, dXb_t, _, Xt = self._cl_backward(cl, dXa * 0, dYb, Xa, Yb, seen)  # :231
            dXb, Xb = self.backward(dXb_t, Xt, scale + 1, "none", seen)  # :232
            dXa = dXa + dXa_t  # :248
        else:
            Xa = Ya.clone()
            dXa = dYa.clone()
            dXa_, dXb, _, Xb = self._cl_backward(cl, dYa * 0, dYb, Ya, Yb, seen)  # :253
            dXa = dXa + dXa_  # :264
        if permute == "lower":
            dXb, Xb = self.C.inverse_tuple(dXb, Xb, faithful_batch_loop=False)  # :268
        dX, X = tensor_cat(dXa, dXb), tensor_cat(Xa, Xb)
        if permute in ("full", "both"):
            dX, X = self.C.inverse_tuple(dX, X, faithful_batch_loop=False)  # :278
        return dX, X


def make_hint_coupling(gen, n_in: int, n_hidden: int, *, logdet=False, permute="none", k1=3, k2=3, p1=1, p2=1,
                       low=0.0, high=1.0, dtype=torch.float32, shared_grads="sum") -> CouplingLayerHINT:
    """invertible_layer_hint.jl:78-101"""
    n = get_depth(n_in)
    cls = []
    for j in range(1, n + 1):
        c = n_in // 2 ** j
        if c * 2 ** j != n_in:
            raise ValueError("InexactError: n_in/2^j is not an integer")  # Int(n_in/2^j) :86
        RB = ResidualBlock(glorot_uniform(gen, n_hidden, c, k1, k1, dtype=dtype),
                           glorot_uniform(gen, n_hidden, n_hidden, k2, k2, dtype=dtype),
                           glorot_uniform(gen, n_hidden, 2 * c, k1, k1, dtype=dtype),
                           torch.zeros(n_hidden, dtype=dtype), torch.zeros(n_hidden, dtype=dtype), p1=p1, p2=p2)
        cls.append(CouplingLayerBasic(RB, logdet=logdet, low=low, high=high))
    if permute in ("full", "both"):
        C = Conv1x1(*(glorot_uniform(gen, n_in, dtype=dtype) for _ in range(3)))
    elif permute == "lower":
        C = Conv1x1(*(glorot_uniform(gen, n_in // 2, dtype=dtype) for _ in range(3)))
    else:
        C = None
    return CouplingLayerHINT(cls, C, logdet=logdet, permute=permute, shared_grads=shared_grads)


# ----------------------------------------------------------------------------------------
# NetworkMultiScaleHINT  (src/networks/invertible_network_hint_multiscale.jl:55-174)
# ----------------------------------------------------------------------------------------
class NetworkMultiScaleHINT:
    def __init__(self, n_in: int, n_hidden: int, L: int, K: int, *, split_scales=False, k1=3, k2=3, p1=1, p2=1,
                 low=0.0, high=1.0, seed: int = 0, dtype=torch.float32, shared_grads="sum", squeezer="wavelet"):
        gen = torch.Generator().manual_seed(seed)
        self.L, self.K, self.split_scales = L, K, split_scales
        self.sq, self.unsq = ((wavelet_squeeze, wavelet_unsqueeze) if squeezer == "wavelet"
                              else (haar_squeeze, inv_haar_unsqueeze))
        self.AN = [[None] * K for _ in range(L)]
        self.CL = [[None] * K for _ in range(L)]
        cf = 2 if split_scales else 4  # :76-82
        for i in range(L):
            for j in range(K):
                self.AN[i][j] = ActNorm(n_in * 4, logdet=True)  # :87
                self.CL[i][j] = make_hint_coupling(gen, n_in * 4, n_hidden, logdet=True, permute="full", k1=k1,
                                                   k2=k2, p1=p1, p2=p2, low=low, high=high, dtype=dtype,
                                                   shared_grads=shared_grads)  # :88-89
            n_in *= cf

    def get_params(self) -> List[Parameter]:
        out: List[Parameter] = []
        for row in self.AN:
            for an in row:
                out += an.params()
        for row in self.CL:
            for cl in row:
                out += cl.params()
        return out

    def forward(self, X: Tensor):  # :98-117
        saved, logdet = [], 0
        for i in range(self.L):
            X = self.sq(X)
            for j in range(self.K):
                X, ld1 = self.AN[i][j].forward(X)
                X, ld2 = self.CL[i][j].forward(X)
                logdet = logdet + (ld1 + ld2)
            if self.split_scales and i < self.L - 1:
                X, Z = tensor_split(X)
                saved.append(Z)
        if self.split_scales:
            self._tail_shape = tuple(X.shape)
            self._saved_shapes = [tuple(z.shape) for z in saved]
            X = cat_states(saved, X)
        return X, logdet

    def _split(self, Z: Tensor):
        outs, o = [], 0
        for s in self._saved_shapes:
            n = math.prod(s)
            outs.append(Z[o:o + n].reshape(s))
            o += n
        return outs, Z[o:].reshape(self._tail_shape)

    def inverse(self, Z: Tensor) -> Tensor:  # :120-133
        if self.split_scales:
            saved, Z = self._split(Z)
        for i in reversed(range(self.L)):
            if self.split_scales and i < self.L - 1:
                Z = tensor_cat(Z, saved[i])
            for j in reversed(range(self.K)):
                Z = self.CL[i][j].inverse(Z)
                Z = self.AN[i][j].inverse(Z)
            Z = self.unsq(Z)
        return Z

    def backward(self, dZ: Tensor, Z: Tensor):  # :136-174 (set_grad = true)
        if self.split_scales:
            dsaved, dZ = self._split(dZ)
            saved, Z = self._split(Z)
        for i in reversed(range(self.L)):
            if self.split_scales and i < self.L - 1:
                dZ = tensor_cat(dZ, dsaved[i])
                Z = tensor_cat(Z, saved[i])
            for j in reversed(range(self.K)):
                dZ, Z = self.CL[i][j].backward(dZ, Z)
                dZ, Z = self.AN[i][j].backward(dZ, Z)
            dZ = self.unsq(dZ)
            Z = self.unsq(Z)
        return dZ, Z


def hint_train_step(H: NetworkMultiScaleHINT, X: Tensor):
    """loss of test_multiscale_hint_network.jl:37-43: f = -log_likelihood(Z) - logdet, ΔZ = Z / B."""
    B = X.shape[0]
    Z, logdet = H.forward(X)
    f = 0.5 * torch.sum(Z * Z) / B - logdet  # objective_functions.jl:54
    dX, _ = H.backward(Z / B, Z)  # :65
    return f, dX
