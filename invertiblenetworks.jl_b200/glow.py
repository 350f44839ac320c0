"""Host-side mirror of the reference's operator interface for the Glow path, in Python because
this image has no Julia (the Julia shim with the same structure is julia/InvertibleNetworksB200.jl).

Same names, argument meaning and error behaviour as the reference:
  NetworkGlow / NetworkConditionalGlow   src/networks/invertible_network_glow.jl:64-191,
                                         src/networks/invertible_network_conditional_glow.jl:64-181
  ActNorm, Conv1x1, ResidualBlock, CouplingLayerGlow   src/layers/*.jl
  Parameter, get_params, clear_grad, set_params, get_grads   src/utils/parameter.jl, neuralnet.jl
  squeeze / unsqueeze (checkerboard)     src/utils/dimensionality_operations.jl

Tensors are float32 CUDA torch tensors in (B, C, [nz,] ny, nx) order - byte-identical to the
reference's column-major (nx, ny[, nz], C, B) CuArrays.  torch is used for device memory and
streams only; every operation is a call into libinb200.so (lib.py).  No CPU fallback exists.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch

from . import lib as _l

Tensor = torch.Tensor


class Parameter:
    """src/utils/parameter.jl:7-10"""
    __slots__ = ("data", "grad")

    def __init__(self, data: Optional[Tensor] = None, grad: Optional[Tensor] = None):
        self.data = data
        self.grad = grad


def clear_grad(obj) -> None:
    """clear_grad! (parameter.jl:53-57, neuralnet.jl:110): gradients become `nothing`."""
    for p in (obj if isinstance(obj, (list, tuple)) else obj.get_params()):
        p.grad = None


def get_params(obj) -> List[Parameter]:
    return obj.get_params()


def get_grads(obj) -> List[Parameter]:
    """neuralnet.jl:120"""
    return [Parameter(p.grad) for p in obj.get_params()]


def set_params(obj, new: Sequence) -> None:
    """set_params! (parameter.jl:73-88): copies data of `new` (Parameters or tensors)."""
    ps = obj.get_params()
    if len(ps) != len(new):
        raise ValueError("parameter count mismatch")
    n_an = getattr(obj, "_n_an", 0)
    an_given = n_an > 0
    for i, (p, q) in enumerate(zip(ps, new)):
        src = q.data if isinstance(q, Parameter) else q
        if src is None:
            if i < n_an:
                an_given = False
            continue
        src = src.to(device=p.data.device, dtype=torch.float32)
        if p.data.shape != src.shape:
            raise ValueError(f"parameter shape mismatch {tuple(p.data.shape)} vs {tuple(src.shape)}")
        p.data.copy_(src)
    # the data-dependent ActNorm initialisation is skipped only when every s, b was actually supplied
    if an_given and hasattr(obj, "_mark_initialized"):
        obj._mark_initialized()


def save_params(obj, path: str) -> None:
    """Checkpoint in get_params order (the Julia side keeps `BSON.@save "net.bson" get_params(G) |> cpu`
    of examples/utils/save_load_network.jl:24-26 unchanged, because the shim never touches the Parameter containers).
    Here: one .npz with arrays p000, p001, ... in get_params order, stored in the REFERENCE's axis order - a torch tensor
    (Cout, Cin, ky, kx) is written as the Julia array (kx, ky, Cin, Cout), so `NPZ.npzread` + `set_params!` loads it."""
    import numpy as np
    if hasattr(obj, "_an_ready") and not obj._an_ready:
        raise _l.InbError("save_params: ActNorm is still uninitialised (run one forward first); its s, b are "
                          "data-dependent (invertible_layer_actnorm.jl:67-72) and would be saved as zeros")
    arrs = {}
    for i, p in enumerate(obj.get_params()):
        a = p.data.detach().cpu().numpy()
        arrs[f"p{i:03d}"] = np.ascontiguousarray(a.transpose(*reversed(range(a.ndim))))
    np.savez(path, **arrs)


def load_params(obj, path: str) -> None:
    """Inverse of save_params (set_params! semantics: shapes must match the architecture)."""
    import numpy as np
    import os
    if not os.path.exists(path) and os.path.exists(path + ".npz"):
        path = path + ".npz"  # numpy.savez appends the suffix when it is missing
    z = np.load(path)
    ps = obj.get_params()
    if len(z.files) != len(ps):
        raise ValueError(f"checkpoint holds {len(z.files)} parameters, the network {len(ps)}")
    new = []
    for i in range(len(ps)):
        a = z[f"p{i:03d}"]
        new.append(torch.from_numpy(np.ascontiguousarray(a.transpose(*reversed(range(a.ndim))))))
    set_params(obj, new)


def _check(x: Tensor, name="X") -> Tensor:
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise _l.InbError(f"{name}: the B200 path takes CUDA tensors only (no CPU fallback)")
    if x.dtype != torch.float32:
        raise _l.InbError(f"{name}: only Float32 is supported on this path (got {x.dtype})")
    return x.contiguous()


def _geom(x: Tensor):
    nsp = x.dim() - 2
    if nsp not in (2, 3):
        raise _l.InbError("expected a 4-D or 5-D tensor (nx, ny[, nz], C, B)")
    if nsp == 2:
        return 2, x.shape[3], x.shape[2], 1
    return 3, x.shape[4], x.shape[3], x.shape[2]


def _glorot(gen, *shape, device):
    """Flux.glorot_uniform: U(-a, a), a = sqrt(6 / (fan_in + fan_out)) (torch weight layout)."""
    if len(shape) == 1:
        fan_in, fan_out = 1, shape[0]
    else:
        rf = int(math.prod(shape[2:]))
        fan_in, fan_out = shape[1] * rf, shape[0] * rf
    a = math.sqrt(6.0 / (fan_in + fan_out))
    w = (torch.rand(*shape, generator=gen, dtype=torch.float64) * 2 - 1) * a
    return w.to(torch.float32).to(device)


def _split_k(c: int) -> int:
    return int(round(c / 2))  # ties to even, like Julia's round (dimensionality_operations.jl:408)


# ------------------------------------------------------------------------------------------
# dimensionality operations
# ------------------------------------------------------------------------------------------
def squeeze(X: Tensor, pattern: str = "checkerboard") -> Tensor:
    """dimensionality_operations.jl:79-107 (checkerboard only on this path)."""
    if pattern != "checkerboard":
        raise _l.InbError("only the checkerboard squeeze is on the B200 path")
    X = _check(X)
    nd, nx, ny, nz = _geom(X)
    if nx % 2 or ny % 2 or (nd == 3 and nz % 2):
        raise _l.InbError("Input dimensions must be multiple of 2")
    B, Cc = X.shape[0], X.shape[1]
    shape = (B, Cc * 2 ** nd) + tuple(s // 2 for s in X.shape[2:])
    Y = torch.empty(shape, device=X.device, dtype=torch.float32)
    _l.call("inb_squeeze", nd, nx, ny, nz, B, Cc, _l.ptr(X), _l.ptr(Y), _l.stream())
    return Y


def unsqueeze(Y: Tensor, pattern: str = "checkerboard") -> Tensor:
    """dimensionality_operations.jl:137-166"""
    if pattern != "checkerboard":
        raise _l.InbError("only the checkerboard squeeze is on the B200 path")
    Y = _check(Y)
    nd, nx, ny, nz = _geom(Y)
    B, Cc = Y.shape[0], Y.shape[1]
    if Cc % (2 ** nd):
        raise _l.InbError(f"number of channels must be divisible by {2 ** nd}")
    shape = (B, Cc // 2 ** nd) + tuple(s * 2 for s in Y.shape[2:])
    X = torch.empty(shape, device=Y.device, dtype=torch.float32)
    _l.call("inb_unsqueeze", nd, nx, ny, nz, B, Cc, _l.ptr(Y), _l.ptr(X), _l.stream())
    return X


# ------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------
class ActNorm:
    """src/layers/invertible_layer_actnorm.jl:42-123"""

    def __init__(self, k: int, logdet: bool = False):
        self.k = k
        self.s = Parameter(None)
        self.b = Parameter(None)
        self.logdet = logdet

    def get_params(self):
        return [self.s, self.b]

    def forward(self, X: Tensor):
        X = _check(X)
        B, Cc = X.shape[0], X.shape[1]
        sp = X[0, 0].numel()
        if self.s.data is None:  # :67-72 data-dependent initialisation
            self.s.data = torch.empty(Cc, device=X.device)
            self.b.data = torch.empty(Cc, device=X.device)
            _l.call("inb_actnorm_init", B, Cc, sp, _l.ptr(X), _l.ptr(self.s.data), _l.ptr(self.b.data),
                    _l.stream())
        Y = torch.empty_like(X)
        ld = torch.empty(1, device=X.device) if self.logdet else None
        _l.call("inb_actnorm_forward", B, Cc, sp, _l.ptr(X), _l.ptr(self.s.data), _l.ptr(self.b.data),
                _l.ptr(Y), _l.ptr(ld), _l.stream())
        return (Y, ld[0]) if self.logdet else Y

    def inverse(self, Y: Tensor):
        Y = _check(Y)
        X = torch.empty_like(Y)
        _l.call("inb_actnorm_inverse", Y.shape[0], Y.shape[1], Y[0, 0].numel(), _l.ptr(Y),
                _l.ptr(self.s.data), _l.ptr(self.b.data), _l.ptr(X), _l.stream())
        return X

    def backward(self, dY: Tensor, Y: Tensor):
        dY, Y = _check(dY), _check(Y)
        dX, X = torch.empty_like(Y), torch.empty_like(Y)
        ds, db = torch.empty_like(self.s.data), torch.empty_like(self.b.data)
        _l.call("inb_actnorm_backward", Y.shape[0], Y.shape[1], Y[0, 0].numel(), _l.ptr(dY), _l.ptr(Y),
                _l.ptr(self.s.data), _l.ptr(self.b.data), int(self.logdet), _l.ptr(dX), _l.ptr(X),
                _l.ptr(ds), _l.ptr(db), _l.stream())
        self.s.grad, self.b.grad = ds, db  # :113-114 overwrite
        return dX, X


class Conv1x1:
    """src/layers/invertible_layer_conv1x1.jl:42-245 (three Householder reflections)"""

    def __init__(self, k: int, freeze: bool = False, gen: Optional[torch.Generator] = None,
                 device="cuda"):
        gen = gen or torch.Generator().manual_seed(0)
        self.k = k
        self.v1 = Parameter(_glorot(gen, k, device=device))
        self.v2 = Parameter(_glorot(gen, k, device=device))
        self.v3 = Parameter(_glorot(gen, k, device=device))
        self.freeze = freeze

    def get_params(self):
        return [self.v1, self.v2, self.v3]

    def _call(self, name, X):
        X = _check(X)
        Y = torch.empty_like(X)
        _l.call(name, X.shape[0], X.shape[1], X[0, 0].numel(), _l.ptr(X), _l.ptr(self.v1.data),
                _l.ptr(self.v2.data), _l.ptr(self.v3.data), _l.ptr(Y), _l.stream())
        return Y

    def forward(self, X):
        if isinstance(X, tuple):
            raise _l.InbError("forward((dX, X)) is not on the Glow training path")
        return self._call("inb_conv1x1_forward", X)

    def inverse(self, Y):
        if isinstance(Y, tuple):  # :227-245
            return self._inverse_tuple(*Y)
        return self._call("inb_conv1x1_inverse", Y)

    def _inverse_tuple(self, dY, Y):
        dY, Y = _check(dY), _check(Y)
        dX, X = torch.empty_like(Y), torch.empty_like(Y)
        d = [torch.empty_like(self.v1.data) for _ in range(3)]
        _l.call("inb_conv1x1_backward", Y.shape[0], Y.shape[1], Y[0, 0].numel(), _l.ptr(dY), _l.ptr(Y),
                _l.ptr(self.v1.data), _l.ptr(self.v2.data), _l.ptr(self.v3.data), int(self.freeze),
                _l.ptr(dX), _l.ptr(X), _l.ptr(d[0]), _l.ptr(d[1]), _l.ptr(d[2]), _l.stream())
        for p, g in zip(self.get_params(), d):  # :237-239 accumulate unless cleared
            p.grad = g if p.grad is None else p.grad + g
        return dX, X


class ResidualBlock:
    """src/layers/layer_residual_block.jl:67-178, fan=true; weights in the reference's bytes:
    W1 (k1..,Cin,nh) == torch (nh,Cin,k1..), W2 (nh,nh,k2..), W3 (k1..,Cout,nh) == torch (nh,Cout,k1..)."""

    def __init__(self, n_in: int, n_hidden: int, n_out: Optional[int] = None, k1=3, k2=3, p1=1, p2=1,
                 ndims=2, precision="fp32", gen: Optional[torch.Generator] = None, device="cuda"):
        gen = gen or torch.Generator().manual_seed(0)
        n_out = 2 * n_in if n_out is None else n_out
        self.n_in, self.n_hidden, self.n_out, self.k1, self.k2, self.ndims = n_in, n_hidden, n_out, k1, k2, ndims
        if p1 != (k1 - 1) // 2 or p2 != (k2 - 1) // 2:
            raise _l.InbError("only 'same' padding is supported on the B200 path")
        self.precision = _l.PRECISIONS[precision]
        kk1, kk2 = (k1,) * ndims, (k2,) * ndims
        self.W1 = Parameter(_glorot(gen, n_hidden, n_in, *kk1, device=device))
        self.W2 = Parameter(_glorot(gen, n_hidden, n_hidden, *kk2, device=device))
        self.W3 = Parameter(_glorot(gen, n_hidden, n_out, *kk1, device=device))
        self.b1 = Parameter(torch.zeros(n_hidden, device=device))
        self.b2 = Parameter(torch.zeros(n_hidden, device=device))

    def get_params(self):
        return [self.W1, self.W2, self.W3, self.b1, self.b2]

    def _ints(self, X):
        nd, nx, ny, nz = _geom(X)
        return [nd, nx, ny, nz, X.shape[0], self.n_in, self.n_hidden, self.n_out, self.k1, self.k2,
                self.precision]

    def forward(self, X: Tensor):
        X = _check(X)
        Y = torch.empty((X.shape[0], self.n_out) + tuple(X.shape[2:]), device=X.device)
        _l.call("inb_resblock_forward", *self._ints(X), _l.ptr(X), *[_l.ptr(p.data) for p in self.get_params()],
                _l.ptr(Y), _l.stream())
        return Y

    def backward(self, dY: Tensor, X: Tensor):
        dY, X = _check(dY), _check(X)
        dX = torch.empty_like(X)
        g = [torch.empty_like(p.data) for p in self.get_params()]
        _l.call("inb_resblock_backward", *self._ints(X), _l.ptr(dY), _l.ptr(X),
                *[_l.ptr(p.data) for p in self.get_params()], _l.ptr(dX), *[_l.ptr(t) for t in g], _l.stream())
        for p, t in zip(self.get_params(), g):  # :168-172 overwrite
            p.grad = t
        return dX


class CouplingLayerGlow:
    """src/layers/invertible_layer_glow.jl:63-170 and, with n_cond > 0,
    src/conditional_layers/conditional_layer_glow.jl:61-158 (ConditionalLayerGlow)."""

    def __init__(self, n_in: int, n_hidden: int, n_cond: int = 0, k1=3, k2=1, p1=1, p2=0, logdet=False,
                 freeze_conv=False, ndims=2, low=0.0, high=1.0, precision="fp32",
                 gen: Optional[torch.Generator] = None, device="cuda"):
        gen = gen or torch.Generator().manual_seed(0)
        k = _split_k(n_in)
        self.n_in, self.n_cond, self.n_hidden = n_in, n_cond, n_hidden
        self.logdet, self.low, self.high = logdet, low, high
        self.C = Conv1x1(n_in, freeze=freeze_conv, gen=gen, device=device)
        self.RB = ResidualBlock(n_in - k + n_cond, n_hidden, n_out=2 * k, k1=k1, k2=k2, p1=p1, p2=p2,
                                ndims=ndims, precision=precision, gen=gen, device=device)

    def get_params(self):
        return self.C.get_params() + self.RB.get_params()

    def _ints(self, X):
        nd, nx, ny, nz = _geom(X)
        return [nd, nx, ny, nz, X.shape[0], self.n_in, self.n_cond, self.n_hidden, self.RB.k1, self.RB.k2]

    def forward(self, X: Tensor, cond: Optional[Tensor] = None):
        X = _check(X)
        cond = _check(cond, "C") if self.n_cond else None
        Y = torch.empty_like(X)
        ld = torch.empty(1, device=X.device) if self.logdet else None
        _l.call("inb_coupling_forward", *self._ints(X), self.low, self.high, self.RB.precision, _l.ptr(X),
                _l.ptr(cond), _l.ptr_table([p.data for p in self.get_params()]), _l.ptr(Y), _l.ptr(ld),
                _l.stream())
        return (Y, ld[0]) if self.logdet else Y

    def inverse(self, Y: Tensor, cond: Optional[Tensor] = None):
        Y = _check(Y)
        cond = _check(cond, "C") if self.n_cond else None
        X = torch.empty_like(Y)
        _l.call("inb_coupling_inverse", *self._ints(Y), self.low, self.high, self.RB.precision, _l.ptr(Y),
                _l.ptr(cond), _l.ptr_table([p.data for p in self.get_params()]), _l.ptr(X), _l.stream())
        return X

    def backward(self, dY: Tensor, Y: Tensor, cond: Optional[Tensor] = None):
        dY, Y = _check(dY), _check(Y)
        cond = _check(cond, "C") if self.n_cond else None
        dX, X = torch.empty_like(Y), torch.empty_like(Y)
        dC = torch.empty_like(cond) if self.n_cond else None
        ps = self.get_params()
        g = [torch.empty_like(p.data) for p in ps]
        _l.call("inb_coupling_backward", *self._ints(Y), self.low, self.high, int(self.logdet),
                int(self.C.freeze), self.RB.precision, _l.ptr(dY), _l.ptr(Y), _l.ptr(cond),
                _l.ptr_table([p.data for p in ps]), _l.ptr_table(g), _l.ptr(dX), _l.ptr(X), _l.ptr(dC),
                _l.stream())
        for i, (p, t) in enumerate(zip(ps, g)):
            if i < 3 and p.grad is not None:  # conv1x1.jl:237-239
                p.grad = p.grad + t
            else:
                p.grad = t
        return (dX, X, dC) if self.n_cond else (dX, X)


# ------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------
class _GlowBase:
    """Parameter storage shared by both networks: ONE flat fp32 buffer for all parameters and one
    for all gradients (get_params order), so data-parallel training all-reduces a single tensor."""

    def _setup(self, n_in, n_cond, n_hidden, L, K, *, logdet, split_scales, ndims, k1, k2, p1, p2, low, high,
               freeze_conv, precision, seed, device):
        if p1 != (k1 - 1) // 2 or p2 != (k2 - 1) // 2:
            raise _l.InbError("only 'same' padding is supported on the B200 path")
        self.n_in, self.n_cond, self.n_hidden, self.L, self.K = n_in, n_cond, n_hidden, L, K
        self.logdet, self.split_scales, self.ndims = logdet, split_scales, ndims
        self.k1, self.k2, self.p1, self.p2, self.low, self.high = k1, k2, p1, p2, low, high
        self.freeze_conv, self.precision = freeze_conv, _l.PRECISIONS[precision]
        self.device = torch.device(device)
        self.Z_dims = None
        self._plan = None
        self._plan_key = None
        self._an_ready = False
        # shapes in get_params order
        shapes = []
        cf = 2 ** ndims if split_scales else 1
        c, cc = n_in, n_cond
        an_shapes, cl_shapes = [], []
        for i in range(L):
            c *= cf
            cc *= cf
            k = _split_k(c)
            for _ in range(K):
                an_shapes += [(c,), (c,)]
                cl_shapes += [(c,), (c,), (c,),
                              (n_hidden, c - k + cc) + (k1,) * ndims,
                              (n_hidden, n_hidden) + (k2,) * ndims,
                              (n_hidden, 2 * k) + (k1,) * ndims,
                              (n_hidden,), (n_hidden,)]
            if i < L - 1 and split_scales:
                c //= 2
        shapes = an_shapes + ([(n_cond,), (n_cond,)] if n_cond else []) + cl_shapes
        sizes = [int(math.prod(s)) for s in shapes]
        # 64-element alignment so every tensor starts on a 256-byte boundary
        offs, o = [], 0
        for n in sizes:
            offs.append(o)
            o += (n + 63) // 64 * 64
        self.flat_params = torch.zeros(o, device=self.device)
        self.flat_grads = torch.zeros(o, device=self.device)
        self._slots = list(zip(offs, sizes, shapes))
        self._params = [Parameter(self.flat_params[a:a + n].view(s)) for a, n, s in zip(offs, sizes, shapes)]
        self._gviews = [self.flat_grads[a:a + n].view(s) for a, n, s in zip(offs, sizes, shapes)]
        self._tabs = None  # (params, grads) device-pointer tables, built on first use
        # initialisation (invertible_layer_conv1x1.jl:54-59, layer_residual_block.jl:92-96): glorot
        # uniform for v and W, zero biases; ActNorm s,b stay unset until the first forward.
        gen = torch.Generator().manual_seed(seed)
        n_an = len(an_shapes) + (2 if n_cond else 0)
        for idx in range(n_an, len(shapes)):
            which = (idx - n_an) % 8
            if which < 6:
                self._params[idx].data.copy_(_glorot(gen, *shapes[idx], device=self.device))
        self._n_an = n_an
        self._hh_idx = [n_an + 8 * l + w for l in range(L * K) for w in range(3)]

    # ---- reference API
    def get_params(self) -> List[Parameter]:
        return self._params

    def _mark_initialized(self):
        self._an_ready = True

    @property
    def _ptab(self):
        # The library, the optimiser and the data-parallel all-reduce work on the flat buffers.  The reference idiom
        # `p.data = new_array` (parameter.jl:73-76 does it in set_params!) is honoured by copying a rebound array back
        # into the parameter's slot of the flat buffer (same shape required), so views and pointer tables stay valid.
        for p, (a, n, shp) in zip(self._params, self._slots):
            d = p.data
            if d is None or d.data_ptr() != self.flat_params.data_ptr() + 4 * a:
                if d is None or tuple(d.shape) != tuple(shp):
                    raise _l.InbError("a network parameter was replaced by an array of another shape (or None)")
                view = self.flat_params[a:a + n].view(shp)
                view.copy_(d.to(device=self.device, dtype=torch.float32))
                p.data = view
        if self._tabs is None:
            self._tabs = (_l.ptr_table([p.data for p in self._params]), _l.ptr_table(self._gviews))
        return self._tabs[0]

    @property
    def _gtab(self):
        self._ptab
        return self._tabs[1]

    def _get_plan(self, X: Tensor):
        nd, nx, ny, nz = _geom(X)
        if nd != self.ndims:
            raise _l.InbError(f"network was built for ndims={self.ndims}")
        if X.shape[1] != self.n_in:
            raise _l.InbError(f"expected {self.n_in} input channels, got {X.shape[1]}")
        B = X.shape[0]
        key = (nx, ny, nz)
        if self._plan is not None and self._plan_key[:3] == key and B <= self._plan_key[3]:
            return self._plan
        self._free_plan()
        d = _l.GlowDesc(nd, nx, ny, nz, self.n_in, self.n_cond, self.n_hidden, self.L, self.K, B,
                        int(self.split_scales), int(self.logdet), self.k1, self.k2, self.p1, self.p2,
                        self.low, self.high, int(self.freeze_conv), self.precision)
        plan = _l.P()
        import ctypes
        _l.call("inb_glow_plan_create", ctypes.byref(d), ctypes.byref(plan))
        self._plan, self._plan_key = plan, key + (B,)
        if getattr(self, "_comm", None) is not None:  # dp.attach: the communicator follows the plan
            _l.call("inb_glow_plan_set_comm", plan, self._comm._h)
        return plan

    def graph_stats(self):
        """{captures, replays, direct} of the plan's CUDA-graph replay (diagnostics for bench.py)."""
        import ctypes
        if getattr(self, "_plan", None) is None:
            return {"captures": 0, "replays": 0, "direct": 0}
        a, b, c = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_longlong()
        _l.call("inb_glow_graph_stats", self._plan, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return {"captures": a.value, "replays": b.value, "direct": c.value}

    def _free_plan(self):
        if getattr(self, "_plan", None) is not None:
            _l.load().inb_glow_plan_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self._free_plan()
        except Exception:
            pass

    def _fill_zdims(self, B):
        import ctypes
        if not self.split_scales:
            return
        n_blocks = max(self.L - 1, 1) if not self.n_cond else self.L - 1
        dims = (ctypes.c_int * 5)()
        out = []
        for i in range(n_blocks):
            n = _l.load().inb_glow_zdims(self._plan, B, i, dims)
            out.append(tuple(dims[j] for j in range(n)))
        self.Z_dims = out  # invertible_network_glow.jl:123

    def _assign_grads(self, saved_hh):
        for p, g in zip(self._params, self._gviews):
            p.grad = g
        for idx, old in saved_hh:  # conv1x1.jl:237-239: Householder grads accumulate unless cleared
            self._params[idx].grad.add_(old)

    def _save_hh(self):
        return [(i, self._params[i].grad.clone()) for i in self._hh_idx if self._params[i].grad is not None]


class NetworkGlow(_GlowBase):
    """G = NetworkGlow(n_in, n_hidden, L, K; logdet, split_scales, k1, k2, p1, p2, ndims, freeze_conv)
    invertible_network_glow.jl:78-104.  `activation=SigmoidLayer(low, high)` is given as low/high."""

    def __init__(self, n_in, n_hidden, L, K, *, logdet=True, split_scales=False, ndims=2, k1=3, k2=1, p1=1,
                 p2=0, low=0.0, high=1.0, freeze_conv=False, precision="fp32", seed=0, device="cuda"):
        if n_in == 1:
            split_scales = True  # :79
        self._setup(n_in, 0, n_hidden, L, K, logdet=logdet, split_scales=split_scales, ndims=ndims, k1=k1,
                    k2=k2, p1=p1, p2=p2, low=low, high=high, freeze_conv=freeze_conv, precision=precision,
                    seed=seed, device=device)

    def forward(self, X: Tensor):
        """Z, logdet = G.forward(X)   (:109-129); Z is the flat latent vector when split_scales."""
        X = _check(X)
        plan = self._get_plan(X)
        B = X.shape[0]
        Z = torch.empty(X.numel() if self.split_scales else X.shape, device=X.device)
        ld = torch.empty(1, device=X.device) if self.logdet else None
        _l.call("inb_glow_forward", plan, B, _l.ptr(X), self._ptab, _l.ptr(Z), _l.ptr(ld),
                int(not self._an_ready), _l.stream())
        self._an_ready = True
        self._fill_zdims(B)
        self._in_shape = tuple(X.shape)
        return (Z, ld[0]) if self.logdet else Z

    def _shape_for(self, Z: Tensor):
        if not self.split_scales:
            return tuple(Z.shape)
        if getattr(self, "_in_shape", None) is None or math.prod(self._in_shape) != Z.numel():
            raise _l.InbError("inverse/backward before forward: Z_dims unknown (invertible_network_glow.jl:123)")
        return self._in_shape

    def inverse(self, Z: Tensor):
        """X = G.inverse(Z)   (:132-147)"""
        Z = _check(Z, "Z")
        shape = self._shape_for(Z)
        X = torch.empty(shape, device=Z.device)
        plan = self._get_plan(X)
        _l.call("inb_glow_inverse", plan, shape[0], _l.ptr(Z), self._ptab, _l.ptr(X), _l.stream())
        return X

    def backward(self, dZ: Tensor, Z: Tensor):
        """dX, X = G.backward(dZ, Z)   (:150-191, set_grad=true); fills p.grad of every parameter."""
        dZ, Z = _check(dZ, "dZ"), _check(Z, "Z")
        shape = self._shape_for(Z)
        X, dX = torch.empty(shape, device=Z.device), torch.empty(shape, device=Z.device)
        plan = self._get_plan(X)
        saved = self._save_hh()
        _l.call("inb_glow_backward", plan, shape[0], _l.ptr(dZ), _l.ptr(Z), self._ptab, self._gtab, _l.ptr(dX),
                _l.ptr(X), _l.stream())
        self._assign_grads(saved)
        return dX, X


def NetworkGlow3D(*args, **kw):
    """invertible_network_glow.jl:106"""
    return NetworkGlow(*args, ndims=3, **kw)


class NetworkConditionalGlow(_GlowBase):
    """G = NetworkConditionalGlow(n_in, n_cond, n_hidden, L, K; ...)  conditional_glow.jl:78-102"""

    def __init__(self, n_in, n_cond, n_hidden, L, K, *, split_scales=False, ndims=2, k1=3, k2=1, p1=1, p2=0,
                 low=0.0, high=1.0, freeze_conv=False, precision="fp32", seed=0, device="cuda"):
        if n_cond < 1:
            raise _l.InbError("n_cond must be >= 1")
        self._setup(n_in, n_cond, n_hidden, L, K, logdet=True, split_scales=split_scales, ndims=ndims, k1=k1,
                    k2=k2, p1=p1, p2=p2, low=low, high=high, freeze_conv=freeze_conv, precision=precision,
                    seed=seed, device=device)

    def _zc_shape(self, Cnd):
        if not self.split_scales:
            return tuple(Cnd.shape)
        f = 2 ** self.L
        return (Cnd.shape[0], Cnd.shape[1] * (2 ** self.ndims) ** self.L) + tuple(s // f for s in Cnd.shape[2:])

    def forward(self, X: Tensor, Cnd: Tensor):
        """ZX, ZC, logdet = G.forward(X, C)   (:107-130)"""
        X, Cnd = _check(X), _check(Cnd, "C")
        if Cnd.shape[1] != self.n_cond or Cnd.shape[0] != X.shape[0] or Cnd.shape[2:] != X.shape[2:]:
            raise _l.InbError("condition must have n_cond channels and X's batch / spatial size")
        plan = self._get_plan(X)
        B = X.shape[0]
        ZX = torch.empty_like(X)  # :128 reshaped to the input's shape
        ZC = torch.empty(self._zc_shape(Cnd), device=X.device)
        ld = torch.empty(1, device=X.device)
        _l.call("inb_cglow_forward", plan, B, _l.ptr(X), _l.ptr(Cnd), self._ptab, _l.ptr(ZX), _l.ptr(ZC),
                _l.ptr(ld), int(not self._an_ready), _l.stream())
        self._an_ready = True
        self._fill_zdims(B)
        return ZX, ZC, ld[0]

    def inverse(self, ZX: Tensor, ZC: Tensor):
        """X = G.inverse(ZX, ZC)   (:133-148)"""
        ZX, ZC = _check(ZX, "ZX"), _check(ZC, "ZC")
        X = torch.empty_like(ZX)
        plan = self._get_plan(X)
        _l.call("inb_cglow_inverse", plan, X.shape[0], _l.ptr(ZX), _l.ptr(ZC), self._ptab, _l.ptr(X), _l.stream())
        return X

    def backward(self, dZX: Tensor, ZX: Tensor, ZC: Tensor):
        """dX, X, dC = G.backward(dZX, ZX, ZC)   (:151-181)"""
        dZX, ZX, ZC = _check(dZX, "dZX"), _check(ZX, "ZX"), _check(ZC, "ZC")
        X, dX = torch.empty_like(ZX), torch.empty_like(ZX)
        plan = self._get_plan(X)
        f = 2 ** self.L if self.split_scales else 1
        cshape = (ZX.shape[0], self.n_cond) + tuple(ZX.shape[2:])
        dC = torch.empty(cshape, device=ZX.device)
        saved = self._save_hh()
        _l.call("inb_cglow_backward", plan, X.shape[0], _l.ptr(dZX), _l.ptr(ZX), _l.ptr(ZC), self._ptab,
                self._gtab, _l.ptr(dX), _l.ptr(X), _l.ptr(dC), _l.stream())
        self._assign_grads(saved)
        return dX, X, dC


def nll_grad(Z: Tensor, batch: int):
    """f = ||Z||^2 / (2B) and dZ = Z / B (objective_functions.jl:54,65) in one pass."""
    Z = _check(Z, "Z")
    dZ = torch.empty_like(Z)
    loss = torch.empty(1, device=Z.device)
    _l.call("inb_nll_grad", Z.numel(), batch, _l.ptr(Z), _l.ptr(dZ), _l.ptr(loss), _l.stream())
    return loss[0], dZ


class ADAM:
    """Flux.Optimise.ADAM(eta, (beta1, beta2)) applied to the flat parameter buffer of a network in one launch:
    `opt = ADAM(G); ...; G.backward(dZ, Z); opt.step()` replaces the loop `for p in get_params(G):
    update!(opt, p.data, p.grad)` of examples/networks/network_glow.jl:38-42 (480 launches for cfg2)."""

    def __init__(self, G, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.G, self.lr, self.betas, self.eps = G, float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.m = torch.zeros_like(G.flat_params)
        self.v = torch.zeros_like(G.flat_params)
        self.t = 0

    def step(self):
        G = self.G
        self.t += 1
        b1, b2 = self.betas
        _l.call("inb_adam_update", G.flat_params.numel(), _l.ptr(G.flat_params), _l.ptr(G.flat_grads), _l.ptr(self.m),
                _l.ptr(self.v), self.lr, b1, b2, self.eps, b1 ** self.t, b2 ** self.t, _l.stream())
