"""CPU oracle for the Glow training hot path of slimgroup/InvertibleNetworks.jl (v2.3.1).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product
path (`invertiblenetworks.jl_b200/`) never imports, calls or falls back to anything here.

PARITY UNPINNED: the reference is pure Julia and there is no Julia toolchain in this image,
and the reference's own tests hold no golden vectors / known-answer files for this path
 - they are property tests only.  This file therefore restates the
reference's algorithm op for op (each function cites the file:line it follows) and is
pinned against (i) every *property* the reference tests assert (tests/test_oracle_*.py) and
(ii) torch.autograd in float64 for every hand-derived gradient.  Two conventions live in
un-vendored third-party packages and cannot be verified here; they are restated from the
packages' published behaviour:
  * NNlib (compat 0.7/0.8/0.9, Project.toml:25) `conv` is a true convolution (kernel
    flipped, `flipped=false` default): conv(x,w) == F.conv2d(x, w.flip(spatial)),
    ∇conv_data(y,w) == F.conv_transpose2d(y, w.flip(spatial)),
    ∇conv_filter(x,dy) == d/dw of that conv.
  * Statistics.var is the unbiased (n-1) estimator.

Memory layout: a Julia array (nx, ny[, nz], C, B) in column-major order is byte-identical to
a C-order torch tensor (B, C[, nz], ny, nx).  All tensors here use the torch order.  Conv
weights (kx, ky[, kz], Cin, Cout) in Julia == torch (Cout, Cin[, kz], ky, kx).

All arithmetic runs in the dtype of the inputs (float32 like the reference, or float64 for the
"truth" column of the parity report).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------
# Parameter container  (src/utils/parameter.jl:7-10, 53-57)
# ----------------------------------------------------------------------------------------
class Parameter:
    __slots__ = ("data", "grad")

    def __init__(self, data: Optional[Tensor], grad: Optional[Tensor] = None):
        self.data = data
        self.grad = grad


def clear_grad(params: Sequence[Parameter]) -> None:
    """src/utils/parameter.jl:53-57 : grads become `nothing`."""
    for p in params:
        p.grad = None


# ----------------------------------------------------------------------------------------
# Dimensionality operations  (src/utils/dimensionality_operations.jl)
# ----------------------------------------------------------------------------------------
def _julia_round_half(c: int) -> int:
    """Int(round(c/2)) with Julia's ties-to-even rounding (dimensionality_operations.jl:408)."""
    return int(round(c / 2))  # Python's round() is ties-to-even as well


def squeeze(X: Tensor) -> Tensor:
    """Checkerboard squeeze, dimensionality_operations.jl:40-47, 79-107.

    Block i=1..2^d takes input pixels (ix::2, iy::2[, iz::2]) with ix=(i+1)%2, iy=((i-1)÷2)%2,
    iz=((i-1)÷4)%2 and writes them to output channels (i-1)*C .. i*C-1.
    """
    nsp = X.dim() - 2
    if any(n % 2 for n in X.shape[2:]):
        raise ValueError("Input dimensions must be multiple of 2")  # :82-84
    blocks = []
    for i in range(2 ** nsp):
        ix, iy, iz = i % 2, (i // 2) % 2, (i // 4) % 2
        if nsp == 2:
            blocks.append(X[:, :, iy::2, ix::2])
        elif nsp == 3:
            blocks.append(X[:, :, iz::2, iy::2, ix::2])
        else:
            raise ValueError("only 2-D / 3-D spatial tensors")
    return torch.cat(blocks, dim=1).contiguous()


def unsqueeze(Y: Tensor) -> Tensor:
    """Inverse of `squeeze`, dimensionality_operations.jl:137-166."""
    nsp = Y.dim() - 2
    nb = 2 ** nsp
    if Y.shape[1] % nb:
        raise ValueError("number of channels must be divisible by 2^(N-2)")  # :141-143
    C = Y.shape[1] // nb
    out_sp = [2 * n for n in Y.shape[2:]]
    X = Y.new_zeros((Y.shape[0], C, *out_sp))
    for i in range(nb):
        ix, iy, iz = i % 2, (i // 2) % 2, (i // 4) % 2
        blk = Y[:, i * C:(i + 1) * C]
        if nsp == 2:
            X[:, :, iy::2, ix::2] = blk
        else:
            X[:, :, iz::2, iy::2, ix::2] = blk
    return X


def tensor_split(X: Tensor, split_index: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    """dimensionality_operations.jl:405-417 (copies, like Julia slicing)."""
    k = _julia_round_half(X.shape[1]) if split_index is None else split_index
    return X[:, :k].clone(), X[:, k:].clone()


def tensor_cat(X: Tensor, Y: Tensor) -> Tensor:
    """dimensionality_operations.jl:435-444."""
    if X.shape[1] == 0:
        return Y
    if Y.shape[1] == 0:
        return X
    return torch.cat((X, Y), dim=1)


def cat_states(Z_save: Sequence[Tensor], X: Tensor) -> Tensor:
    """dimensionality_operations.jl:474-476 : [vec(Z_1); ...; vec(Z_{L-1}); vec(X)]."""
    return torch.cat([z.reshape(-1) for z in Z_save] + [X.reshape(-1)])


def _last_dims(zd: Sequence[int], squeezed: bool) -> Tuple[int, ...]:
    """xy_dims, dimensionality_operations.jl:464-467 (torch dim order (B,C,spatial...))."""
    if not squeezed:
        return tuple(zd)
    nsp = len(zd) - 2
    return (zd[0], zd[1] * 2 ** nsp, *[n // 2 for n in zd[2:]])


def split_states(Y: Tensor, Z_dims: Sequence[Sequence[int]], L_net: int = 2):
    """dimensionality_operations.jl:480-490."""
    L = len(Z_dims) + 1
    sizes = [int(math.prod(zd)) for zd in Z_dims]
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    Z_save = [Y[offs[j]:offs[j + 1]].reshape(tuple(Z_dims[j])) for j in range(L - 1)]
    X = Y[offs[L - 1]:].reshape(_last_dims(Z_dims[-1], L_net > 1))
    return Z_save, X


# ----------------------------------------------------------------------------------------
# Activations  (src/utils/activation_functions.jl)
# ----------------------------------------------------------------------------------------
def relu(x: Tensor) -> Tensor:
    """activation_functions.jl:63 (NNlib relu = max(0,x))."""
    return torch.clamp_min(x, 0)


def relu_grad(dy: Tensor, x: Tensor) -> Tensor:
    """activation_functions.jl:82-84 : passes Δy where x >= 0 (NOT torch's x > 0)."""
    return torch.where(x < 0, torch.zeros_like(dy), dy)


def sigmoid(x: Tensor, low: float = 0.0, high: float = 1.0) -> Tensor:
    """activation_functions.jl:160-162 : high/(1+exp(-x)) + low/(1+exp(x))."""
    return high / (1 + torch.exp(-x)) + low / (1 + torch.exp(x))


def sigmoid_inv(y: Tensor, low: float = 0.0, high: float = 1.0) -> Tensor:
    """activation_functions.jl:180."""
    return torch.log(y - low) - torch.log(high - y)


def sigmoid_grad(dy: Tensor, y: Tensor, low: float = 0.0, high: float = 1.0) -> Tensor:
    """activation_functions.jl:213-217 : evaluated from the *output* through a logit round trip."""
    x = sigmoid_inv(y, low, high)
    e = torch.exp(-x)
    return (high - low) * dy * e / (1 + e) ** 2


# ----------------------------------------------------------------------------------------
# NNlib convolutions restated  (external, see module docstring)
# ----------------------------------------------------------------------------------------
def _flip(w: Tensor) -> Tensor:
    return w.flip(tuple(range(2, w.dim())))


def nn_conv(x: Tensor, w: Tensor, pad: int, stride: int = 1) -> Tensor:
    f = F.conv2d if w.dim() == 4 else F.conv3d
    return f(x, _flip(w), padding=pad, stride=stride)


def nn_conv_data(y: Tensor, w: Tensor, pad: int, stride: int = 1) -> Tensor:
    f = F.conv_transpose2d if w.dim() == 4 else F.conv_transpose3d
    return f(y, _flip(w), padding=pad, stride=stride)


def nn_conv_filter(x: Tensor, dy: Tensor, wshape, pad: int, stride: int = 1) -> Tensor:
    if len(wshape) == 4:
        g = torch.nn.grad.conv2d_weight(x, tuple(wshape), dy, padding=pad, stride=stride)
    else:
        g = torch.nn.grad.conv3d_weight(x, tuple(wshape), dy, padding=pad, stride=stride)
    return _flip(g)


# ----------------------------------------------------------------------------------------
# ActNorm  (src/layers/invertible_layer_actnorm.jl:42-123, 185-195)
# ----------------------------------------------------------------------------------------
class ActNorm:
    def __init__(self, k: int, logdet: bool = False):
        self.k = k
        self.s = Parameter(None)
        self.b = Parameter(None)
        self.logdet = logdet

    def params(self) -> List[Parameter]:
        return [self.s, self.b]

    @staticmethod
    def _bc(v: Tensor, X: Tensor) -> Tensor:
        return v.reshape(1, -1, *([1] * (X.dim() - 2)))

    def forward(self, X: Tensor, logdet: Optional[bool] = None):
        logdet = self.logdet if logdet is None else logdet
        red = [0] + list(range(2, X.dim()))
        if self.s.data is None:  # :67-72 data-dependent init, Statistics.var is unbiased
            mu = X.mean(dim=red)
            var = X.var(dim=red, unbiased=True)
            self.s.data = 1 / torch.sqrt(var)
            self.b.data = -mu / torch.sqrt(var)
        Y = X * self._bc(self.s.data, X) + self._bc(self.b.data, X)  # :73
        if logdet:  # :185-195  prod(spatial) * sum(log|s|), no batch division
            nsp = math.prod(X.shape[2:])
            return Y, nsp * torch.sum(torch.log(torch.abs(self.s.data)))
        return Y

    def inverse(self, Y: Tensor) -> Tensor:
        return (Y - self._bc(self.b.data, Y)) / self._bc(self.s.data, Y)  # :93

    def backward(self, dY: Tensor, Y: Tensor):
        red = [0] + list(range(2, Y.dim()))
        nsp = math.prod(Y.shape[2:])
        X = self.inverse(Y)  # :105
        dX = dY * self._bc(self.s.data, Y)  # :106
        ds = torch.sum(dY * X, dim=red)  # :107
        if self.logdet:
            ds = ds - nsp / self.s.data  # :108-110, logdet_backward :190
        db = torch.sum(dY, dim=red)  # :111
        self.s.grad = ds  # :113-114 (overwrite)
        self.b.grad = db
        return dX, X


# ----------------------------------------------------------------------------------------
# Conv1x1 Householder channel mix  (src/layers/invertible_layer_conv1x1.jl, compute_utils.jl)
# ----------------------------------------------------------------------------------------
def chain_lr(x: Tensor, *vs: Tensor) -> Tensor:
    """compute_utils.jl:20-30 applied to all batch elements at once.

    x: (B, k, px) is, per batch element, the (px x k) matrix Xi of conv1x1.jl:184 stored with
    channel stride px.  One reflection: tmp = out*v ; tmp *= -2/dot(v,v) ; out += tmp*v'.
    """
    out = 1 * x
    for v in vs:
        n = -2 / torch.dot(v, v)
        tmp = torch.einsum("bkp,k->bp", out, v)
        tmp = tmp * n
        out = out + tmp[:, None, :] * v[None, :, None]
    return out


def partial_derivative_outer(v: Tensor) -> Tensor:
    """conv1x1.jl:69-87 : outer[i,:,:] = d/dv_i ( v v' / (v'v) )."""
    k = v.numel()
    out1 = torch.outer(v, v)
    n = torch.dot(v, v)
    outer = out1[None, :, :].repeat(k, 1, 1)
    outer = v[:, None, None] * outer
    outer = (-2 / n) * outer
    for j in range(k):
        outer[j, :, j] = outer[j, :, j] + v
        outer[j, j, :] = outer[j, j, :] + v
    return (1 / n) * outer


class Conv1x1:
    def __init__(self, v1: Tensor, v2: Tensor, v3: Tensor, freeze: bool = False):
        self.k = v1.numel()
        self.v1, self.v2, self.v3 = Parameter(v1), Parameter(v2), Parameter(v3)
        self.freeze = freeze

    def params(self) -> List[Parameter]:
        return [self.v1, self.v2, self.v3]

    def _flat(self, X: Tensor) -> Tensor:
        return X.reshape(X.shape[0], X.shape[1], -1)

    def forward(self, X: Tensor) -> Tensor:
        """conv1x1.jl:174-189 : Y_i = X_i H1 H2 H3."""
        return chain_lr(self._flat(X), self.v1.data, self.v2.data, self.v3.data).reshape(X.shape)

    def inverse(self, Y: Tensor) -> Tensor:
        """conv1x1.jl:209-224 : X_i = Y_i H3 H2 H1."""
        return chain_lr(self._flat(Y), self.v3.data, self.v2.data, self.v1.data).reshape(Y.shape)

    def grad_v(self, X: Tensor, dY: Tensor, faithful_batch_loop: bool = True, adjoint: bool = False):
        """conv1x1.jl:118-170; adjoint=false is what `inverse(::Tuple)` uses (:234), adjoint=true what
        `forward(::Tuple)` uses (:196): every dV_j[i] is transposed (:150,154,157)."""
        v1, v2, v3 = self.v1.data, self.v2.data, self.v3.data
        k = self.k
        if self.freeze:  # :132-134
            z = torch.zeros_like(v1)
            return z, z.clone(), z.clone()
        eye = torch.eye(k, dtype=v1.dtype)
        V1 = torch.outer(v1, v1) / torch.dot(v1, v1)
        V2 = torch.outer(v2, v2) / torch.dot(v2, v2)
        V3 = torch.outer(v3, v3) / torch.dot(v3, v3)
        dV1, dV2, dV3 = (partial_derivative_outer(v) for v in (v1, v2, v3))
        M1 = eye - 2 * (V2 + V3) + 4 * V2 @ V3  # :144
        M3 = eye - 2 * (V1 + V2) + 4 * V1 @ V2  # :145
        for i in range(k):  # :147-158
            dV1[i] = dV1[i] @ M1
            t = dV2[i]
            dV2[i] = t + (4 * V1 @ t @ V3 - 2 * (V1 @ t + t @ V3))
            dV3[i] = M3 @ dV3[i]
            if adjoint:
                dV1[i], dV2[i], dV3[i] = dV1[i].T.clone(), dV2[i].T.clone(), dV3[i].T.clone()
        Xf, dYf = self._flat(X), self._flat(dY)
        dv = [torch.zeros_like(v1) for _ in range(3)]
        if faithful_batch_loop:
            # :160-168 with mat_tens_i :110-116 : per batch element, per i, dot(Xi*dV[i], dYi)
            for b in range(Xf.shape[0]):
                Xi = -2 * Xf[b].transpose(0, 1)  # (px, k)
                dYi = dYf[b].transpose(0, 1)
                for j, dV in enumerate((dV1, dV2, dV3)):
                    prod = torch.stack([torch.sum((Xi @ dV[i]) * dYi) for i in range(k)])
                    dv[j] = dv[j] + prod
        else:
            # same quantity through the Gram matrix  sum_b Xi' * dYi  
            G = -2 * torch.einsum("bkp,blp->kl", Xf, dYf)
            for j, dV in enumerate((dV1, dV2, dV3)):
                dv[j] = torch.einsum("ikl,kl->i", dV, G)
        return dv[0], dv[1], dv[2]

    def forward_tuple(self, dX: Tensor, X: Tensor, faithful_batch_loop: bool = True):
        """conv1x1.jl:191-205 : (ΔY, Y) = (forward(ΔX), forward(X)), Δv from (Y, ΔX) with adjoint=true, accumulated."""
        dY = self.forward(dX)
        Y = self.forward(X)
        d1, d2, d3 = self.grad_v(Y, dX, faithful_batch_loop, adjoint=True)
        for p, d in ((self.v1, d1), (self.v2, d2), (self.v3, d3)):
            p.grad = d if p.grad is None else p.grad + d
        return dY, Y

    def inverse_tuple(self, dY: Tensor, Y: Tensor, faithful_batch_loop: bool = True):
        """conv1x1.jl:227-245 : (ΔX, X) and accumulate Δv (+= when already set, :237-239)."""
        dX = self.inverse(dY)
        X = self.inverse(Y)
        d1, d2, d3 = self.grad_v(X, dY, faithful_batch_loop)
        for p, d in ((self.v1, d1), (self.v2, d2), (self.v3, d3)):
            p.grad = d if p.grad is None else p.grad + d
        return dX, X


# ----------------------------------------------------------------------------------------
# ResidualBlock  (src/layers/layer_residual_block.jl:67-178, InvertibleNetworks.jl:31-36)
# ----------------------------------------------------------------------------------------
class ResidualBlock:
    """fan=True variant used by Glow (invertible_layer_glow.jl:95).  Weights in torch layout:
    W1 (nh, Cin, k1..), W2 (nh, nh, k2..), W3 (nh, Cout, k1..)."""

    def __init__(self, W1, W2, W3, b1, b2, p1: int = 1, p2: int = 0):
        self.W1, self.W2, self.W3 = Parameter(W1), Parameter(W2), Parameter(W3)
        self.b1, self.b2 = Parameter(b1), Parameter(b2)
        self.p1, self.p2 = p1, p2

    def params(self) -> List[Parameter]:
        return [self.W1, self.W2, self.W3, self.b1, self.b2]  # struct field order :67-73

    @staticmethod
    def _bc(v: Tensor, X: Tensor) -> Tensor:
        return v.reshape(1, -1, *([1] * (X.dim() - 2)))

    def forward(self, X1: Tensor, save: bool = False):
        Y1 = nn_conv(X1, self.W1.data, self.p1) + self._bc(self.b1.data, X1)  # :122
        X2 = relu(Y1)
        Y2 = X2 + nn_conv(X2, self.W2.data, self.p2) + self._bc(self.b2.data, X1)  # :125
        X3 = relu(Y2)
        Y3 = nn_conv_data(X3, self.W3.data, self.p1)  # :128-129 (DCDims :31-36)
        if save:
            return Y1, Y2, Y3
        return relu(Y3)  # :133 fan == true

    def backward(self, dX4: Tensor, X1: Tensor) -> Tensor:
        red = [0] + list(range(2, X1.dim()))
        Y1, Y2, Y3 = self.forward(X1, save=True)  # :143 recompute
        dY3 = relu_grad(dX4, Y3)  # :150
        dX3 = nn_conv(dY3, self.W3.data, self.p1)  # :151
        dW3 = nn_conv_filter(dY3, relu(Y2), self.W3.data.shape, self.p1)  # :152
        dY2 = relu_grad(dX3, Y2)  # :154
        dX2 = nn_conv_data(dY2, self.W2.data, self.p2) + dY2  # :155
        dW2 = nn_conv_filter(relu(Y1), dY2, self.W2.data.shape, self.p2)  # :156
        db2 = torch.sum(dY2, dim=red)  # :157
        dY1 = relu_grad(dX2, Y1)  # :161
        dX1 = nn_conv_data(dY1, self.W1.data, self.p1)  # :162
        dW1 = nn_conv_filter(X1, dY1, self.W1.data.shape, self.p1)  # :163
        db1 = torch.sum(dY1, dim=red)  # :164
        self.W1.grad, self.W2.grad, self.W3.grad = dW1, dW2, dW3  # :168-172 overwrite
        self.b1.grad, self.b2.grad = db1, db2
        return dX1


# ----------------------------------------------------------------------------------------
# CouplingLayerGlow / ConditionalLayerGlow
# (src/layers/invertible_layer_glow.jl:63-170,210-211; conditional_layers/conditional_layer_glow.jl)
# ----------------------------------------------------------------------------------------
def glow_logdet_forward(S: Tensor) -> Tensor:
    return torch.sum(torch.log(torch.abs(S))) / S.shape[0]  # :210 (batch is dim 0 here)


def glow_logdet_backward(S: Tensor) -> Tensor:
    return 1 / S / S.shape[0]  # :211


class CouplingLayerGlow:
    def __init__(self, C: Conv1x1, RB: ResidualBlock, logdet: bool = False,
                 low: float = 0.0, high: float = 1.0, faithful_batch_loop: bool = True):
        self.C, self.RB, self.logdet = C, RB, logdet
        self.low, self.high = low, high
        self.faithful = faithful_batch_loop

    def params(self) -> List[Parameter]:
        return self.C.params() + self.RB.params()

    def _rb_in(self, X2: Tensor, cond: Optional[Tensor]) -> Tensor:
        return X2 if cond is None else tensor_cat(X2, cond)  # conditional_layer_glow.jl:102

    def forward(self, X: Tensor, cond: Optional[Tensor] = None):
        X_ = self.C.forward(X)  # :105
        X1, X2 = tensor_split(X_)
        Y2 = X2.clone()
        logS_T = self.RB.forward(self._rb_in(X2, cond))  # :109
        logS, T = tensor_split(logS_T)
        S = sigmoid(logS, self.low, self.high)
        Y1 = S * X1 + T  # :112
        Y = tensor_cat(Y1, Y2)
        if self.logdet:
            return Y, glow_logdet_forward(S)
        return Y

    def inverse(self, Y: Tensor, cond: Optional[Tensor] = None, save: bool = False):
        Y1, Y2 = tensor_split(Y)
        X2 = Y2.clone()
        logS_T = self.RB.forward(self._rb_in(X2, cond))  # :124
        logS, T = tensor_split(logS_T)
        S = sigmoid(logS, self.low, self.high)
        eps = torch.finfo(Y.dtype).eps
        X1 = (Y1 - T) / (S + eps)  # :127  eps(T) of the array eltype
        X_ = tensor_cat(X1, X2)
        X = self.C.inverse(X_)
        if save:
            return X, X1, X2, S
        return X

    def backward(self, dY: Tensor, Y: Tensor, cond: Optional[Tensor] = None):
        X, X1, X2, S = self.inverse(Y, cond, save=True)  # :139
        dY1, dY2 = tensor_split(dY)
        dT = dY1.clone()
        dS = dY1 * X1  # :144
        if self.logdet:
            dS = dS - glow_logdet_backward(S)  # :145-147
        dX1 = dY1 * S  # :149
        dRB = self.RB.backward(tensor_cat(sigmoid_grad(dS, S, self.low, self.high), dT),
                               self._rb_in(X2, cond))  # :151
        if cond is None:
            dX2 = dRB + dY2
            dC = None
        else:  # conditional_layer_glow.jl:150-152
            dX2, dC = tensor_split(dRB, split_index=dY2.shape[1])
            dX2 = dX2 + dY2
        dX_ = tensor_cat(dX1, dX2)
        dX, _ = self.C.inverse_tuple(dX_, tensor_cat(X1, X2), self.faithful)  # :159
        if cond is None:
            return dX, X
        return dX, X, dC


# ----------------------------------------------------------------------------------------
# Weight initialisation helpers (Flux.glorot_uniform semantics; parity tests inject weights)
# ----------------------------------------------------------------------------------------
def glorot_uniform(gen: torch.Generator, *shape_torch: int, dtype=torch.float32) -> Tensor:
    """Flux.glorot_uniform(dims...) = (rand - 0.5) * sqrt(24 / (fan_in + fan_out)), i.e.
    U(-a, a) with a = sqrt(6 / (fan_in + fan_out)); restated for the torch weight layout
    (Cout, Cin, k...).  A length-n vector has nfan = (1, n)."""
    if len(shape_torch) == 1:
        fan_in, fan_out = 1, shape_torch[0]
    else:
        rf = int(math.prod(shape_torch[2:]))
        fan_in, fan_out = shape_torch[1] * rf, shape_torch[0] * rf
    a = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(*shape_torch, generator=gen, dtype=torch.float64) * 2 - 1) * a).to(dtype)


def make_coupling(gen, n_in: int, n_hidden: int, n_cond: int = 0, ndims: int = 2, logdet=True,
                  k1=3, k2=1, p1=1, p2=0, low=0.0, high=1.0, freeze=False,
                  dtype=torch.float32, faithful=True) -> CouplingLayerGlow:
    """invertible_layer_glow.jl:82-99 / conditional_layer_glow.jl:77-89 (channel bookkeeping)."""
    split_num = _julia_round_half(n_in)
    in_chan = n_in - split_num + n_cond
    out_chan = 2 * split_num
    C = Conv1x1(*(glorot_uniform(gen, n_in, dtype=dtype) for _ in range(3)), freeze=freeze)
    kk1, kk2 = (k1,) * ndims, (k2,) * ndims
    RB = ResidualBlock(glorot_uniform(gen, n_hidden, in_chan, *kk1, dtype=dtype),
                       glorot_uniform(gen, n_hidden, n_hidden, *kk2, dtype=dtype),
                       glorot_uniform(gen, n_hidden, out_chan, *kk1, dtype=dtype),
                       torch.zeros(n_hidden, dtype=dtype), torch.zeros(n_hidden, dtype=dtype),
                       p1=p1, p2=p2)
    return CouplingLayerGlow(C, RB, logdet=logdet, low=low, high=high, faithful_batch_loop=faithful)


# ----------------------------------------------------------------------------------------
# NetworkGlow  (src/networks/invertible_network_glow.jl:64-191)
# ----------------------------------------------------------------------------------------
class NetworkGlow:
    def __init__(self, n_in: int, n_hidden: int, L: int, K: int, *, logdet=True,
                 split_scales=False, ndims=2, k1=3, k2=1, p1=1, p2=0, low=0.0, high=1.0,
                 freeze_conv=False, seed: int = 0, dtype=torch.float32, faithful=True):
        if n_in == 1:
            split_scales = True  # :79
        gen = torch.Generator().manual_seed(seed)
        self.L, self.K, self.logdet, self.split_scales, self.ndims = L, K, logdet, split_scales, ndims
        self.Z_dims: Optional[List[Tuple[int, ...]]] = [(1, 1)] * max(L - 1, 1) if split_scales else None
        cf = 2 ** ndims if split_scales else 1
        self.AN = [[None] * K for _ in range(L)]
        self.CL = [[None] * K for _ in range(L)]
        for i in range(L):
            n_in *= cf  # :94
            for j in range(K):
                self.AN[i][j] = ActNorm(n_in, logdet=logdet)
                self.CL[i][j] = make_coupling(gen, n_in, n_hidden, ndims=ndims, logdet=logdet,
                                              k1=k1, k2=k2, p1=p1, p2=p2, low=low, high=high,
                                              freeze=freeze_conv, dtype=dtype, faithful=faithful)
            if i < L - 1 and split_scales:
                n_in //= 2  # :100

    def get_params(self) -> List[Parameter]:
        """neuralnet.jl:72-88 : field order AN (row-major [i,j]; s,b) then CL (v1,v2,v3,W1,W2,W3,b1,b2)."""
        out: List[Parameter] = []
        for row in self.AN:
            for an in row:
                out += an.params()
        for row in self.CL:
            for cl in row:
                out += cl.params()
        return out

    def forward(self, X: Tensor):
        Z_save: List[Optional[Tensor]] = [None] * max(self.L - 1, 1)
        logdet_ = 0
        for i in range(self.L):
            if self.split_scales:
                X = squeeze(X)  # :114
            for j in range(self.K):
                if self.logdet:
                    X, ld1 = self.AN[i][j].forward(X)
                    X, ld2 = self.CL[i][j].forward(X)
                    logdet_ = logdet_ + (ld1 + ld2)
                else:
                    X = self.AN[i][j].forward(X)
                    X = self.CL[i][j].forward(X)
            if self.split_scales and (i < self.L - 1 or i == 0):  # :120
                X, Z = tensor_split(X)
                Z_save[i] = Z
                self.Z_dims[i] = tuple(Z.shape)
        if self.split_scales:
            X = cat_states(Z_save, X)  # :126
        return (X, logdet_) if self.logdet else X

    def inverse(self, Z: Tensor) -> Tensor:
        X = Z
        if self.split_scales:
            Z_save, X = split_states(X, self.Z_dims, L_net=self.L)
        for i in reversed(range(self.L)):
            if self.split_scales and (i < self.L - 1 or self.L == 1):  # :136
                X = tensor_cat(X, Z_save[i])
            for j in reversed(range(self.K)):
                X = self.CL[i][j].inverse(X)
                X = self.AN[i][j].inverse(X)
            if self.split_scales:
                X = unsqueeze(X)
        return X

    def backward(self, dZ: Tensor, Z: Tensor):
        dX, X = dZ, Z
        if self.split_scales:  # :154-157
            dX_save, dX = split_states(dX, self.Z_dims, L_net=self.L)
            X_save, X = split_states(X, self.Z_dims, L_net=self.L)
        for i in reversed(range(self.L)):
            if self.split_scales and (i < self.L - 1 or self.L == 1):
                X = tensor_cat(X, X_save[i])
                dX = tensor_cat(dX, dX_save[i])
            for j in reversed(range(self.K)):
                dX, X = self.CL[i][j].backward(dX, X)
                dX, X = self.AN[i][j].backward(dX, X)
            if self.split_scales:
                X = unsqueeze(X)
                dX = unsqueeze(dX)
        return dX, X


# ----------------------------------------------------------------------------------------
# NetworkConditionalGlow  (src/networks/invertible_network_conditional_glow.jl:64-181)
# ----------------------------------------------------------------------------------------
class NetworkConditionalGlow:
    def __init__(self, n_in: int, n_cond: int, n_hidden: int, L: int, K: int, *,
                 split_scales=False, ndims=2, k1=3, k2=1, p1=1, p2=0, low=0.0, high=1.0,
                 freeze_conv=False, seed: int = 0, dtype=torch.float32, faithful=True):
        gen = torch.Generator().manual_seed(seed)
        self.L, self.K, self.split_scales, self.ndims = L, K, split_scales, ndims
        self.Z_dims = [(1, 1)] * (L - 1) if split_scales else None
        cf = 2 ** ndims if split_scales else 1
        self.AN = [[None] * K for _ in range(L)]
        self.CL = [[None] * K for _ in range(L)]
        self.AN_C = ActNorm(n_cond, logdet=False)  # :80
        for i in range(L):
            n_in *= cf
            n_cond *= cf  # :93 the condition is squeezed but never split
            for j in range(K):
                self.AN[i][j] = ActNorm(n_in, logdet=True)
                self.CL[i][j] = make_coupling(gen, n_in, n_hidden, n_cond=n_cond, ndims=ndims,
                                              logdet=True, k1=k1, k2=k2, p1=p1, p2=p2, low=low,
                                              high=high, freeze=freeze_conv, dtype=dtype,
                                              faithful=faithful)
            if i < L - 1 and split_scales:
                n_in //= 2

    def get_params(self) -> List[Parameter]:
        """struct field order :64-73 : AN, AN_C, CL."""
        out: List[Parameter] = []
        for row in self.AN:
            for an in row:
                out += an.params()
        out += self.AN_C.params()
        for row in self.CL:
            for cl in row:
                out += cl.params()
        return out

    def forward(self, X: Tensor, C: Tensor):
        Z_save: List[Optional[Tensor]] = [None] * (self.L - 1)
        orig_shape = X.shape
        C = self.AN_C.forward(C)  # :111
        logdet = 0
        for i in range(self.L):
            if self.split_scales:
                X = squeeze(X)
                C = squeeze(C)
            for j in range(self.K):
                X, ld1 = self.AN[i][j].forward(X)
                X, ld2 = self.CL[i][j].forward(X, C)
                logdet = logdet + (ld1 + ld2)
            if self.split_scales and i < self.L - 1:  # :122
                X, Z = tensor_split(X)
                Z_save[i] = Z
                self.Z_dims[i] = tuple(Z.shape)
        if self.split_scales:
            X = cat_states(Z_save, X).reshape(orig_shape)  # :128
        return X, C, logdet

    def inverse(self, X: Tensor, C: Tensor) -> Tensor:
        if self.split_scales:
            Z_save, X = split_states(X.reshape(-1), self.Z_dims)
        for i in reversed(range(self.L)):
            if self.split_scales and i < self.L - 1:
                X = tensor_cat(X, Z_save[i])
            for j in reversed(range(self.K)):
                X = self.CL[i][j].inverse(X, C)
                X = self.AN[i][j].inverse(X)
            if self.split_scales:
                X = unsqueeze(X)
                C = unsqueeze(C)
        return X

    def backward(self, dX: Tensor, X: Tensor, C: Tensor):
        if self.split_scales:
            dZ_save, dX = split_states(dX.reshape(-1), self.Z_dims)
            Z_save, X = split_states(X.reshape(-1), self.Z_dims)
        dC = 0 * C  # :160
        for i in reversed(range(self.L)):
            if self.split_scales and i < self.L - 1:
                X = tensor_cat(X, Z_save[i])
                dX = tensor_cat(dX, dZ_save[i])
            for j in reversed(range(self.K)):
                dX, X, dC_ = self.CL[i][j].backward(dX, X, C)
                dX, X = self.AN[i][j].backward(dX, X)
                dC = dC + dC_
            if self.split_scales:
                C = unsqueeze(C)
                dC = unsqueeze(dC)
                X = unsqueeze(X)
                dX = unsqueeze(dX)
        dC, C = self.AN_C.backward(dC, C)  # :179
        return dX, X, dC


# ----------------------------------------------------------------------------------------
# Objective  (src/utils/objective_functions.jl:54,65) and the training step of
# examples/networks/network_glow.jl:26-43 / glow_seismic.jl:100-123
# ----------------------------------------------------------------------------------------
def log_likelihood(X: Tensor, batch: int) -> Tensor:
    """objective_functions.jl:54 with mu=0, sigma=1; `batch` is size(X, N) of the reference."""
    return (1 / batch) * torch.sum(-0.5 * X ** 2)


def glow_train_step(G: NetworkGlow, X: Tensor):
    """forward -> f = ||Z||^2/(2B) - logdet -> ΔZ = Z/B -> backward.  Returns (f, ΔX)."""
    B = X.shape[0]
    if G.logdet:
        Z, ld = G.forward(X)
    else:
        Z, ld = G.forward(X), 0.0
    f = 0.5 * torch.sum(Z * Z) / B - ld
    dZ = Z / B
    for p in G.get_params():
        p.grad = None
    dX, _ = G.backward(dZ, Z)
    return f, dX
