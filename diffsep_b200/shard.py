"""Utterance sharding across the GPUs of one box: contiguous blocks of ceil(B/G) utterances per
rank, no data-path collective, ONE all-gather of the estimates at the end.  The
reference's own multi-GPU scheme is process-per-shard over dataset indices with no collective
(``evaluate_mp.py:465-513``); this keeps its partitioning and adds the gather."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int):
    """[lo, hi) of the contiguous block owned by ``rank`` (blocks of ceil(n/world))."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = -(-n_items // world)
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def gather_estimates(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gathers per-rank estimates [b_r, C, T] into [n_items, C, T] in global utterance order.
    Ranks may own fewer items than ceil(n/world) (ragged tail, even zero): blocks are padded for the
    collective and trimmed afterwards."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = -(-n_items // world)
    C, T = local.shape[1], local.shape[2]
    pad = torch.zeros(per, C, T, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty(world * per, C, T, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:n_items]


def separate_sharded(model, mix: torch.Tensor, group=None, **sampler_kwargs):
    """mix: the FULL batch [B,1,T] (host or device) on every rank -> estimates [B,2,T] on every rank."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_bounds(mix.shape[0], rank, world)
    dev = model.dev
    if hi > lo:
        (mine, _), _, _ = model.normalize_batch((mix[lo:hi].to(dev), None))
        est, nfe = model.get_pc_sampler("reverse_diffusion", "ald2", mine, **sampler_kwargs)()
    else:
        est, nfe = torch.zeros(0, 2, mix.shape[2], device=dev), 0
    return gather_estimates(est, mix.shape[0], group), nfe
