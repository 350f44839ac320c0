"""ctypes binding of libinb200.so (include/inb200.h) - the stand-in, in this Julia-less image, for
the `ccall` shim of julia/InvertibleNetworksB200.jl.  Every compute call goes through the C ABI;
there is NO fallback: if the shared library is missing the import of a symbol raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libinb200.so")

PREC_FP32, PREC_BF16X3, PREC_BF16, PREC_FP16X3 = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16, "fp16x3": PREC_FP16X3}


class GlowDesc(C.Structure):
    _fields_ = [
        ("ndims", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
        ("n_in", C.c_int), ("n_cond", C.c_int), ("n_hidden", C.c_int),
        ("L", C.c_int), ("K", C.c_int), ("batch", C.c_int),
        ("split_scales", C.c_int), ("logdet", C.c_int),
        ("k1", C.c_int), ("k2", C.c_int), ("p1", C.c_int), ("p2", C.c_int),
        ("sig_low", C.c_float), ("sig_high", C.c_float),
        ("freeze_conv", C.c_int), ("precision", C.c_int),
    ]


class HintDesc(C.Structure):
    _fields_ = [
        ("nx", C.c_int), ("ny", C.c_int), ("n_in", C.c_int), ("n_hidden", C.c_int),
        ("L", C.c_int), ("K", C.c_int), ("batch", C.c_int), ("split_scales", C.c_int),
        ("k1", C.c_int), ("k2", C.c_int), ("p1", C.c_int), ("p2", C.c_int),
        ("sig_low", C.c_float), ("sig_high", C.c_float),
        ("squeeze_type", C.c_int), ("shared_grads", C.c_int), ("precision", C.c_int),
    ]


SQUEEZE_TYPES = {"wavelet": 0, "haar": 1}
PERMUTES = {"none": 0, "full": 1, "lower": 2, "both": 3}
SHARED_GRADS = {"sum": 0, "last": 1}

P = C.c_void_p      # device pointer / stream / plan
PP = C.POINTER(C.c_void_p)
I = C.c_int
LL = C.c_longlong
F = C.c_float

# name -> (restype, argtypes); mirrors include/inb200.h one to one (checked by tests/test_abi.py)
SIGNATURES = {
    "inb_last_error": (C.c_char_p, []),
    "inb_version": (I, []),
    "inb_device_ok": (I, []),
    "inb_glow_plan_create": (I, [C.POINTER(GlowDesc), C.POINTER(P)]),
    "inb_glow_plan_destroy": (I, [P]),
    "inb_glow_num_params": (I, [P]),
    "inb_glow_param_numel": (I, [P, I, C.POINTER(LL)]),
    "inb_glow_workspace_bytes": (LL, [P]),
    "inb_glow_zdims": (I, [P, I, I, C.POINTER(I)]),
    "inb_glow_forward": (I, [P, I, P, PP, P, P, I, P]),
    "inb_glow_inverse": (I, [P, I, P, PP, P, P]),
    "inb_glow_backward": (I, [P, I, P, P, PP, PP, P, P, P]),
    "inb_glow_flat_layout": (I, [P, C.POINTER(LL), C.POINTER(LL)]),
    "inb_comm_unique_id": (I, [C.c_char_p]),
    "inb_comm_create": (I, [I, I, C.c_char_p, C.POINTER(P)]),
    "inb_comm_wrap": (I, [P, C.POINTER(P)]),
    "inb_comm_destroy": (I, [P]),
    "inb_comm_info": (I, [P, C.POINTER(I), C.POINTER(I), C.POINTER(LL), C.POINTER(LL)]),
    "inb_glow_plan_set_comm": (I, [P, P]),
    "inb_allreduce_grads": (I, [P, PP, P, P]),
    "inb_broadcast_params": (I, [P, PP, P, I, P]),
    "inb_cglow_forward": (I, [P, I, P, P, PP, P, P, P, I, P]),
    "inb_cglow_inverse": (I, [P, I, P, P, PP, P, P]),
    "inb_cglow_backward": (I, [P, I, P, P, P, PP, PP, P, P, P, P]),
    "inb_actnorm_init": (I, [I, I, LL, P, P, P, P]),
    "inb_actnorm_forward": (I, [I, I, LL, P, P, P, P, P, P]),
    "inb_actnorm_inverse": (I, [I, I, LL, P, P, P, P, P]),
    "inb_actnorm_backward": (I, [I, I, LL, P, P, P, P, I, P, P, P, P, P]),
    "inb_conv1x1_forward": (I, [I, I, LL, P, P, P, P, P, P]),
    "inb_conv1x1_inverse": (I, [I, I, LL, P, P, P, P, P, P]),
    "inb_conv1x1_backward": (I, [I, I, LL, P, P, P, P, P, I, P, P, P, P, P, P]),
    "inb_resblock_forward": (I, [I] * 11 + [P] * 8),
    "inb_resblock_backward": (I, [I] * 11 + [P] * 14),
    "inb_coupling_forward": (I, [I] * 10 + [F, F, I, P, P, PP, P, P, P]),
    "inb_coupling_inverse": (I, [I] * 10 + [F, F, I, P, P, PP, P, P]),
    "inb_coupling_backward": (I, [I] * 10 + [F, F, I, I, I, P, P, P, PP, PP, P, P, P, P]),
    "inb_squeeze": (I, [I, I, I, I, I, I, P, P, P]),
    "inb_unsqueeze": (I, [I, I, I, I, I, I, P, P, P]),
    "inb_haar_squeeze": (I, [I, I, I, I, I, P, P, P]),
    "inb_haar_unsqueeze": (I, [I, I, I, I, I, P, P, P]),
    "inb_basic_coupling_forward": (I, [I] * 9 + [F, F, I, P, P, PP, P, P, P]),
    "inb_basic_coupling_inverse": (I, [I] * 9 + [F, F, I, P, P, PP, P, P]),
    "inb_basic_coupling_backward": (I, [I] * 9 + [F, F, I, I, P, P, P, P, PP, PP, P, P, P, P]),
    "inb_hint_depth": (I, [I]),
    "inb_hint_coupling_forward": (I, [I] * 9 + [F, F, I, I, P, PP, P, P, P]),
    "inb_hint_coupling_inverse": (I, [I] * 9 + [F, F, I, I, P, PP, P, P]),
    "inb_hint_coupling_backward": (I, [I] * 9 + [F, F, I, I, I, I, P, P, PP, PP, P, P, P]),
    "inb_hint_plan_create": (I, [C.POINTER(HintDesc), C.POINTER(P)]),
    "inb_hint_plan_destroy": (I, [P]),
    "inb_hint_num_params": (I, [P]),
    "inb_hint_param_numel": (I, [P, I, C.POINTER(LL)]),
    "inb_hint_workspace_bytes": (LL, [P]),
    "inb_hint_forward": (I, [P, I, P, PP, P, P, I, P]),
    "inb_hint_inverse": (I, [P, I, P, PP, P, P]),
    "inb_hint_backward": (I, [P, I, P, P, PP, PP, P, P, P]),
    "inb_nll_grad": (I, [LL, I, P, P, P, P]),
    "inb_adam_update": (I, [LL, P, P, P, P, F, F, F, F, F, F, P]),
    "inb_launch_count": (LL, []),
    "inb_prof_enable": (I, [I]),
    "inb_debug_chain_trace": (I, [P]),
    "inb_glow_graph_stats": (I, [P, C.POINTER(LL), C.POINTER(LL), C.POINTER(LL)]),
    "inb_prof_reset": (I, []),
    "inb_prof_num": (I, []),
    "inb_prof_get": (I, [I, C.c_char_p, I, C.POINTER(LL), C.POINTER(LL), C.POINTER(C.c_double),
                         C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


class InbError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise InbError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                " - there is no CPU or PyTorch fallback for this path")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def call(name: str, *args) -> None:
    """Call an int-status entry point and raise InbError with inb_last_error() on failure
    (the Julia shim rethrows the same text as an ErrorException)."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise InbError(lib.inb_last_error().decode())


def ptr(t) -> int:
    """Device pointer of a contiguous float32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    import torch
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise InbError("expected a contiguous float32 CUDA tensor (CuArray{Float32} in the reference)")
    return t.data_ptr()


def ptr_table(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = ptr(t)
    return arr


def launch_count() -> int:
    return load().inb_launch_count()


def prof_table():
    """[{name, launches, scopes, ms, flops, bytes}] for every kernel family with launches."""
    lib = load()
    out = []
    name = C.create_string_buffer(64)
    la, sc = LL(), LL()
    ms, fl, by = C.c_double(), C.c_double(), C.c_double()
    for i in range(lib.inb_prof_num()):
        lib.inb_prof_get(i, name, 64, C.byref(la), C.byref(sc), C.byref(ms), C.byref(fl), C.byref(by))
        if la.value:
            out.append(dict(name=name.value.decode(), launches=la.value, scopes=sc.value, ms=ms.value,
                            flops=fl.value, bytes=by.value))
    return out


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
