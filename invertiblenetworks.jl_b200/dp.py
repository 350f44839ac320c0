"""Data-parallel host logic of the Glow training path: one process per GPU, the batch sharded
along its outermost dimension, parameters replicated, ONE all-reduce (average) over the flat gradient buffer
per step.  Averaging the per-shard gradients reproduces the single-process gradient of the global batch:
every term of the objective f = ||Z||^2/(2B) - logdet is a mean over samples except ActNorm's
prod(spatial)*sum(log|s|) (invertible_layer_actnorm.jl:185-195), which is batch-independent and therefore
invariant under averaging.  ActNorm's data-dependent initialisation (:67-72) runs on rank 0's shard and the
flat parameter buffer is broadcast, so every rank starts from identical parameters.

Works on any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist


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
