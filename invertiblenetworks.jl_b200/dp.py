"""Data-parallel host logic of the Glow training path: one process per GPU, the batch sharded
along its outermost dimension, parameters replicated, ONE all-reduce (average) over the flat gradient buffer
per step.  Averaging the per-shard gradients reproduces the single-process gradient of the global batch:
every term of the objective f = ||Z||^2/(2B) - logdet is a mean over samples except ActNorm's
prod(spatial)*sum(log|s|) (invertible_layer_actnorm.jl:185-195), which is batch-independent and therefore
invariant under averaging.  ActNorm's data-dependent initialisation (:67-72) runs on rank 0's shard and the
flat parameter buffer is broadcast, so every rank starts from identical parameters.

The data plane is the library's own (include/inb200.h, "data-parallel training"): `Communicator` wraps an `inb_comm`
(an ncclComm_t created through the C ABI from a unique id, or torch's own communicator handed over as a raw pointer),
`attach(G, comm)` makes `G.backward` average every scale's gradients over the ranks on the communicator's stream while
the remaining scales still run, and makes the first `G.forward` initialise ActNorm from the statistics of the GLOBAL
batch.  torch.distributed is used for the rendezvous only (shipping the 128-byte id, barriers); the helpers at the top
of this file work on any backend and carry the CPU (gloo) tests of the host logic."""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from . import lib as _l


def shard_bounds(global_batch: int, rank: int, world: int):
    """[lo, hi) of the contiguous slab of samples owned by `rank` (the batch is the outermost memory
    dimension, so a shard is one contiguous block: no halos, no data-path collective)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} does not divide over {world} ranks")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def flatten(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([t.reshape(-1) for t in tensors])


def unflatten_into(flat: torch.Tensor, tensors: Sequence[torch.Tensor]) -> None:
    o = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[o:o + n].view_as(t))
        o += n


def broadcast_params(flat_params: torch.Tensor, src: int = 0) -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat_params, src=src)


def allreduce_grads(flat_grads: torch.Tensor) -> None:
    """In-place average over the ranks (ncclAvg on NCCL; sum / world elsewhere)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat_grads, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
        flat_grads.div_(dist.get_world_size())


def allreduce_mean_scalar(x: torch.Tensor) -> torch.Tensor:
    """Mean over the ranks of a per-shard mean (the reported loss / coupling logdet)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    y = x.clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM)
    return y / dist.get_world_size()


# ------------------------------------------------------------------------------------------
# the C-ABI data plane (csrc/dp.cu): NCCL through libinb200, no torch collective on the data path
# ------------------------------------------------------------------------------------------
class Communicator:
    """An `inb_comm` (include/inb200.h).  Build one per process with `Communicator.from_dist()` (own ncclComm_t, id
    shipped through torch.distributed), `Communicator.create(...)` (any other rendezvous) or
    `Communicator.wrap_torch()` (torch's communicator, shared)."""

    def __init__(self, handle: ctypes.c_void_p):
        self._h = handle
        self._attached = []  # weak references to the networks this communicator is attached to

    @staticmethod
    def unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        _l.call("inb_comm_unique_id", buf)
        return buf.raw

    @classmethod
    def create(cls, nranks: int, rank: int, uid: bytes) -> "Communicator":
        if len(uid) != 128:
            raise _l.InbError("the NCCL unique id is 128 bytes")
        h = _l.P()
        _l.call("inb_comm_create", nranks, rank, ctypes.create_string_buffer(uid, 128), ctypes.byref(h))
        return cls(h)

    @classmethod
    def from_dist(cls, group=None) -> "Communicator":
        """Rank 0 draws the id, torch.distributed ships it (works on gloo and nccl groups), every rank joins."""
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls.create(world, rank, box[0])

    @classmethod
    def wrap_torch(cls, group=None, device=None) -> "Communicator":
        """torch's own ncclComm_t (ProcessGroupNCCL._comm_ptr()): the library's collectives and torch's share it."""
        pg = group if group is not None else dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        ptr = pg._get_backend(dev)._comm_ptr()
        h = _l.P()
        _l.call("inb_comm_wrap", ctypes.c_void_p(ptr), ctypes.byref(h))
        return cls(h)

    def info(self):
        n, r, calls, nbytes = ctypes.c_int(), ctypes.c_int(), ctypes.c_longlong(), ctypes.c_longlong()
        _l.call("inb_comm_info", self._h, ctypes.byref(n), ctypes.byref(r), ctypes.byref(calls), ctypes.byref(nbytes))
        return {"nranks": n.value, "rank": r.value, "allreduce_calls": calls.value, "allreduce_bytes": nbytes.value}

    def destroy(self):
        """Detaches from every network first: a plan's captured graphs hold the communicator's collectives and NCCL
        wants them gone before ncclCommDestroy."""
        if self._h:
            for ref in self._attached:
                G = ref()
                if G is not None and getattr(G, "_comm", None) is self:
                    attach(G, None)
            self._attached = []
            _l.load().inb_comm_destroy(self._h)
            self._h = None


def attach(G, comm: Optional[Communicator]) -> None:
    """inb_glow_plan_set_comm for the network's plan (now and whenever the plan is rebuilt): global-batch ActNorm
    initialisation in the first forward, per-scale overlapped gradient averaging in backward."""
    import weakref
    G._comm = comm
    if comm is not None:
        comm._attached.append(weakref.ref(G))
    if getattr(G, "_plan", None) is not None:
        _l.call("inb_glow_plan_set_comm", G._plan, comm._h if comm is not None else None)


def allreduce_grads_abi(G, comm: Communicator) -> None:
    """The same average as ONE explicit call on the current stream (inb_allreduce_grads): for callers that do not
    attach the communicator (no overlap with backward)."""
    _l.call("inb_allreduce_grads", G._plan, G._gtab, comm._h, _l.stream())


def broadcast_params_abi(G, comm: Communicator, root: int = 0) -> None:
    _l.call("inb_broadcast_params", G._plan, G._ptab, comm._h, root, _l.stream())
