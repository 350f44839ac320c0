"""Host-side mirror of the reference's HINT interface, same conventions as glow.py:
  wavelet_squeeze / wavelet_unsqueeze, Haar_squeeze / invHaar_unsqueeze   src/utils/dimensionality_operations.jl:199-371
  CouplingLayerHINT (over CouplingLayerBasic)                             src/layers/invertible_layer_hint.jl:52-297
  NetworkMultiScaleHINT                                                   src/networks/invertible_network_hint_multiscale.jl
Every operation is a call into libinb200.so; there is no CPU fallback."""
from __future__ import annotations

import ctypes
import math
from typing import List, Optional

import torch

from . import lib as _l
from .glow import Parameter, Tensor, _check, _geom, _glorot


def _haar(name: str, X: Tensor, kind: str, up: bool) -> Tensor:
    X = _check(X)
    if X.dim() != 4:
        raise _l.InbError("the Haar / wavelet squeeze is on the B200 path for 4-D tensors only")
    B, Cc, ny, nx = X.shape
    if up:
        if Cc % 4:
            raise _l.InbError("number of channels must be divisible by 4")
        out = torch.empty(B, Cc // 4, ny * 2, nx * 2, device=X.device)
    else:
        if nx % 2 or ny % 2:
            raise _l.InbError("Input dimensions must be multiple of 2")
        out = torch.empty(B, Cc * 4, ny // 2, nx // 2, device=X.device)
    _l.call(name, nx, ny, B, Cc, _l.SQUEEZE_TYPES[kind], _l.ptr(X), _l.ptr(out), _l.stream())
    return out


def wavelet_squeeze(X: Tensor) -> Tensor:
    """dimensionality_operations.jl:199-216 with type = WT.db1"""
    return _haar("inb_haar_squeeze", X, "wavelet", False)


def wavelet_unsqueeze(Y: Tensor) -> Tensor:
    """dimensionality_operations.jl:243-258"""
    return _haar("inb_haar_unsqueeze", Y, "wavelet", True)


def Haar_squeeze(X: Tensor) -> Tensor:
    """dimensionality_operations.jl:318-331"""
    return _haar("inb_haar_squeeze", X, "haar", False)


def invHaar_unsqueeze(Y: Tensor) -> Tensor:
    """dimensionality_operations.jl:354-371"""
    return _haar("inb_haar_unsqueeze", Y, "haar", True)


def get_depth(n_in: int) -> int:
    """invertible_layer_hint.jl:63-71"""
    return _l.load().inb_hint_depth(n_in)


def _hint_shapes(n_in, n_hidden, k1, k2, ndims, with_conv):
    """parameter shapes of one CouplingLayerHINT in get_params order (torch weight layout)"""
    shapes = []
    for j in range(1, get_depth(n_in) + 1):
        c = n_in // 2 ** j
        shapes += [(n_hidden, c) + (k1,) * ndims, (n_hidden, n_hidden) + (k2,) * ndims,
                   (n_hidden, 2 * c) + (k1,) * ndims, (n_hidden,), (n_hidden,)]
    if with_conv:
        shapes += [(with_conv,)] * 3
    return shapes


class CouplingLayerBasic:
    """L = CouplingLayerBasic(n_in, n_hidden; k1, k2, p1, p2, logdet, activation=SigmoidLayer(low, high))
    invertible_layer_basic.jl:79-86: X1 (n_in channels) conditions, X2 (n_in channels) is transformed."""

    def __init__(self, n_in: int, n_hidden: int, *, k1=3, k2=3, p1=1, p2=1, logdet=False, ndims=2, low=0.0, high=1.0,
                 precision="fp32", gen: Optional[torch.Generator] = None, device="cuda"):
        if p1 != (k1 - 1) // 2 or p2 != (k2 - 1) // 2:
            raise _l.InbError("only 'same' padding is supported on the B200 path")
        gen = gen or torch.Generator().manual_seed(0)
        self.n_in, self.n_hidden, self.k1, self.k2, self.ndims = n_in, n_hidden, k1, k2, ndims
        self.logdet, self.low, self.high, self.precision = logdet, low, high, _l.PRECISIONS[precision]
        shapes = [(n_hidden, n_in) + (k1,) * ndims, (n_hidden, n_hidden) + (k2,) * ndims,
                  (n_hidden, 2 * n_in) + (k1,) * ndims, (n_hidden,), (n_hidden,)]
        self._params = [Parameter(_glorot(gen, *s, device=device) if len(s) > 1 else torch.zeros(s, device=device))
                        for s in shapes]

    def get_params(self) -> List[Parameter]:
        return self._params

    def _args(self, X1, X2):
        nd, nx, ny, nz = _geom(X1)
        if X1.shape != X2.shape or X1.shape[1] != self.n_in:
            raise _l.InbError(f"expected two tensors of {self.n_in} channels and equal shape")
        return [nd, nx, ny, nz, X1.shape[0], self.n_in, self.n_hidden, self.k1, self.k2, self.low, self.high]

    def forward(self, X1: Tensor, X2: Tensor):
        """X1, Y2[, logdet] = L.forward(X1, X2)   (:90-105)"""
        X1, X2 = _check(X1, "X1"), _check(X2, "X2")
        Y2 = torch.empty_like(X2)
        ld = torch.empty(1, device=X1.device) if self.logdet else None
        _l.call("inb_basic_coupling_forward", *self._args(X1, X2), self.precision, _l.ptr(X1), _l.ptr(X2),
                _l.ptr_table([p.data for p in self._params]), _l.ptr(Y2), _l.ptr(ld), _l.stream())
        return (X1, Y2, ld[0]) if self.logdet else (X1, Y2)

    def inverse(self, Y1: Tensor, Y2: Tensor):
        """Y1, X2 = L.inverse(Y1, Y2)   (:108-121)"""
        Y1, Y2 = _check(Y1, "Y1"), _check(Y2, "Y2")
        X2 = torch.empty_like(Y2)
        _l.call("inb_basic_coupling_inverse", *self._args(Y1, Y2), self.precision, _l.ptr(Y1), _l.ptr(Y2),
                _l.ptr_table([p.data for p in self._params]), _l.ptr(X2), _l.stream())
        return Y1, X2

    def backward(self, dY1: Tensor, dY2: Tensor, Y1: Tensor, Y2: Tensor):
        """dX1, dX2, X1, X2 = L.backward(dY1, dY2, Y1, Y2)   (:124-149, set_grad=true)"""
        dY1, dY2, Y1, Y2 = _check(dY1, "dY1"), _check(dY2, "dY2"), _check(Y1, "Y1"), _check(Y2, "Y2")
        dX1, dX2, X2 = torch.empty_like(Y1), torch.empty_like(Y2), torch.empty_like(Y2)
        g = [torch.empty_like(p.data) for p in self._params]
        _l.call("inb_basic_coupling_backward", *self._args(Y1, Y2), int(self.logdet), self.precision, _l.ptr(dY1),
                _l.ptr(dY2), _l.ptr(Y1), _l.ptr(Y2), _l.ptr_table([p.data for p in self._params]), _l.ptr_table(g),
                _l.ptr(dX1), _l.ptr(dX2), _l.ptr(X2), _l.stream())
        for p, t in zip(self._params, g):  # layer_residual_block.jl:168-172: overwritten
            p.grad = t
        return dX1, dX2, Y1, X2


class CouplingLayerHINT:
    """H = CouplingLayerHINT(n_in, n_hidden; logdet, permute, k1, k2, p1, p2, activation=SigmoidLayer(low, high))
    invertible_layer_hint.jl:78-101.  permute in {"none", "full", "lower", "both"}."""

    def __init__(self, n_in: int, n_hidden: int, *, logdet=False, permute="none", k1=3, k2=3, p1=1, p2=1, ndims=2,
                 low=0.0, high=1.0, shared_grads="sum", precision="fp32", gen: Optional[torch.Generator] = None,
                 device="cuda"):
        if permute not in _l.PERMUTES:
            raise _l.InbError(f"unknown permute mode {permute!r}")
        if p1 != (k1 - 1) // 2 or p2 != (k2 - 1) // 2:
            raise _l.InbError("only 'same' padding is supported on the B200 path")
        gen = gen or torch.Generator().manual_seed(0)
        self.n_in, self.n_hidden, self.logdet, self.permute = n_in, n_hidden, logdet, permute
        self.k1, self.k2, self.ndims, self.low, self.high = k1, k2, ndims, low, high
        self.shared_grads, self.precision = shared_grads, _l.PRECISIONS[precision]
        nconv = {"none": 0, "full": n_in, "lower": n_in // 2, "both": n_in}[permute]
        shapes = _hint_shapes(n_in, n_hidden, k1, k2, ndims, nconv)
        self._params = []
        for i, s in enumerate(shapes):
            is_bias = len(s) == 1 and i < len(shapes) - (3 if nconv else 0)
            self._params.append(Parameter(torch.zeros(s, device=device) if is_bias else _glorot(gen, *s, device=device)))

    def get_params(self) -> List[Parameter]:
        return self._params

    def _args(self, X):
        nd, nx, ny, nz = _geom(X)
        if X.shape[1] != self.n_in:
            raise _l.InbError(f"expected {self.n_in} channels, got {X.shape[1]}")
        return [nd, nx, ny, nz, X.shape[0], self.n_in, self.n_hidden, self.k1, self.k2, self.low, self.high,
                _l.PERMUTES[self.permute]]

    def forward(self, X: Tensor):
        X = _check(X)
        Y = torch.empty_like(X)
        ld = torch.empty(1, device=X.device) if self.logdet else None
        _l.call("inb_hint_coupling_forward", *self._args(X), self.precision, _l.ptr(X),
                _l.ptr_table([p.data for p in self._params]), _l.ptr(Y), _l.ptr(ld), _l.stream())
        return (Y, ld[0]) if self.logdet else Y

    def inverse(self, Y: Tensor):
        Y = _check(Y)
        X = torch.empty_like(Y)
        _l.call("inb_hint_coupling_inverse", *self._args(Y), self.precision, _l.ptr(Y),
                _l.ptr_table([p.data for p in self._params]), _l.ptr(X), _l.stream())
        return X

    def backward(self, dY: Tensor, Y: Tensor):
        dY, Y = _check(dY), _check(Y)
        dX, X = torch.empty_like(Y), torch.empty_like(Y)
        g = [torch.empty_like(p.data) for p in self._params]
        _l.call("inb_hint_coupling_backward", *self._args(Y), int(self.logdet), _l.SHARED_GRADS[self.shared_grads],
                self.precision, _l.ptr(dY), _l.ptr(Y), _l.ptr_table([p.data for p in self._params]),
                _l.ptr_table(g), _l.ptr(dX), _l.ptr(X), _l.stream())
        nconv = 0 if self.permute == "none" else 3
        for i, (p, t) in enumerate(zip(self._params, g)):
            if i >= len(g) - nconv and p.grad is not None:  # conv1x1.jl:237-239
                p.grad = p.grad + t
            else:
                p.grad = t
        return dX, X


class NetworkMultiScaleHINT:
    """H = NetworkMultiScaleHINT(n_in, n_hidden, L, K; split_scales, k1, k2, p1, p2, activation)
    invertible_network_hint_multiscale.jl:68-94 (2-D).  Parameters live in one flat buffer (get_params order), the
    gradients in another, like the Glow networks."""

    def __init__(self, n_in, n_hidden, L, K, *, split_scales=False, k1=3, k2=3, p1=1, p2=1, low=0.0, high=1.0,
                 squeezer="wavelet", shared_grads="sum", precision="fp32", seed=0, device="cuda"):
        if p1 != (k1 - 1) // 2 or p2 != (k2 - 1) // 2:
            raise _l.InbError("only 'same' padding is supported on the B200 path")
        self.n_in, self.n_hidden, self.L, self.K, self.split_scales = n_in, n_hidden, L, K, split_scales
        self.k1, self.k2, self.p1, self.p2, self.low, self.high = k1, k2, p1, p2, low, high
        self.squeezer, self.shared_grads, self.precision = squeezer, shared_grads, _l.PRECISIONS[precision]
        self.device = torch.device(device)
        self.logdet = True
        self._plan, self._plan_key, self._an_ready, self._tabs = None, None, False, None
        an_shapes, cl_shapes, self._hh_idx = [], [], []
        c = n_in
        for i in range(L):
            C4 = 4 * c
            for _ in range(K):
                an_shapes += [(C4,), (C4,)]
                cl_shapes += _hint_shapes(C4, n_hidden, k1, k2, 2, C4)
            c = C4 // 2 if split_scales else C4
        shapes = an_shapes + cl_shapes
        sizes = [int(math.prod(s)) for s in shapes]
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += (n + 63) // 64 * 64
        self.flat_params = torch.zeros(o, device=self.device)
        self.flat_grads = torch.zeros(o, device=self.device)
        self._params = [Parameter(self.flat_params[a:a + n].view(s)) for a, n, s in zip(offs, sizes, shapes)]
        self._gviews = [self.flat_grads[a:a + n].view(s) for a, n, s in zip(offs, sizes, shapes)]
        gen = torch.Generator().manual_seed(seed)
        idx = len(an_shapes)
        c = n_in
        for i in range(L):
            C4 = 4 * c
            per = _hint_shapes(C4, n_hidden, k1, k2, 2, C4)
            for _ in range(K):
                for w, s in enumerate(per):
                    if len(s) > 1 or w >= len(per) - 3:  # weights and Householder vectors: glorot; biases stay zero
                        self._params[idx].data.copy_(_glorot(gen, *s, device=self.device))
                    if w >= len(per) - 3:
                        self._hh_idx.append(idx)
                    idx += 1
            c = C4 // 2 if split_scales else C4

    def get_params(self) -> List[Parameter]:
        return self._params

    def _mark_initialized(self):
        self._an_ready = True

    def _tables(self):
        if self._tabs is None:
            self._tabs = (_l.ptr_table([p.data for p in self._params]), _l.ptr_table(self._gviews))
        return self._tabs

    def _get_plan(self, X: Tensor):
        if X.dim() != 4 or X.shape[1] != self.n_in:
            raise _l.InbError(f"expected a (B, {self.n_in}, ny, nx) tensor")
        B, _, ny, nx = X.shape
        if self._plan is not None and self._plan_key[:2] == (nx, ny) and B <= self._plan_key[2]:
            return self._plan
        self._free_plan()
        d = _l.HintDesc(nx, ny, self.n_in, self.n_hidden, self.L, self.K, B, int(self.split_scales), self.k1,
                        self.k2, self.p1, self.p2, self.low, self.high, _l.SQUEEZE_TYPES[self.squeezer],
                        _l.SHARED_GRADS[self.shared_grads], self.precision)
        plan = _l.P()
        _l.call("inb_hint_plan_create", ctypes.byref(d), ctypes.byref(plan))
        n = _l.load().inb_hint_num_params(plan)
        if n != len(self._params):
            raise _l.InbError(f"internal: library counts {n} parameters, host {len(self._params)}")
        self._plan, self._plan_key = plan, (nx, ny, B)
        return plan

    def _free_plan(self):
        if getattr(self, "_plan", None) is not None:
            _l.load().inb_hint_plan_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self._free_plan()
        except Exception:
            pass

    def _z_shape(self, X):
        if self.split_scales:
            return (X.numel(),)
        f = 2 ** self.L
        return (X.shape[0], X.shape[1] * 4 ** self.L, X.shape[2] // f, X.shape[3] // f)

    def forward(self, X: Tensor):
        """Z, logdet = H.forward(X)   (:98-117)"""
        X = _check(X)
        plan = self._get_plan(X)
        Z = torch.empty(self._z_shape(X), device=X.device)
        ld = torch.empty(1, device=X.device)
        _l.call("inb_hint_forward", plan, X.shape[0], _l.ptr(X), self._tables()[0], _l.ptr(Z), _l.ptr(ld),
                int(not self._an_ready), _l.stream())
        self._an_ready = True
        self._in_shape = tuple(X.shape)
        return Z, ld[0]

    def _shape_for(self, Z: Tensor):
        if getattr(self, "_in_shape", None) is None or math.prod(self._in_shape) != Z.numel():
            if self.split_scales:
                raise _l.InbError("inverse/backward before forward: X_dims unknown (hint_multiscale.jl:111)")
            f = 2 ** self.L
            return (Z.shape[0], Z.shape[1] // 4 ** self.L, Z.shape[2] * f, Z.shape[3] * f)
        return self._in_shape

    def inverse(self, Z: Tensor):
        """X = H.inverse(Z)   (:120-133)"""
        Z = _check(Z, "Z")
        X = torch.empty(self._shape_for(Z), device=Z.device)
        plan = self._get_plan(X)
        _l.call("inb_hint_inverse", plan, X.shape[0], _l.ptr(Z), self._tables()[0], _l.ptr(X), _l.stream())
        return X

    def backward(self, dZ: Tensor, Z: Tensor):
        """dX, X = H.backward(dZ, Z)   (:136-174, set_grad=true)"""
        dZ, Z = _check(dZ, "dZ"), _check(Z, "Z")
        shape = self._shape_for(Z)
        X, dX = torch.empty(shape, device=Z.device), torch.empty(shape, device=Z.device)
        plan = self._get_plan(X)
        saved = [(i, self._params[i].grad.clone()) for i in self._hh_idx if self._params[i].grad is not None]
        _l.call("inb_hint_backward", plan, shape[0], _l.ptr(dZ), _l.ptr(Z), *self._tables(), _l.ptr(dX), _l.ptr(X),
                _l.stream())
        for p, g in zip(self._params, self._gviews):
            p.grad = g
        for idx, old in saved:  # conv1x1.jl:237-239
            self._params[idx].grad.add_(old)
        return dX, X
