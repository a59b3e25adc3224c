"""Multi-GPU execution of the denoising path: prompts are independent, so they are sharded across
ranks (one process per GPU, weights replicated, each prompt's CFG cond/uncond pair on one device) with
NO collective inside the loop and one all-gather of the results at the end.
``torch.distributed`` is the plumbing: backend "nccl" over NVLink/NVSwitch on GPUs, "gloo" in the CPU
tests.  Noise is derived per *prompt index*, so results do not depend on the world size.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) chunk of ``n_items`` for ``rank``; the first ``n_items % world`` ranks get one more."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def prompt_seed(base_seed: int, prompt_index: int) -> int:
    """Seed of the noise stream of one prompt: a function of the global prompt index only."""
    return (base_seed * 0x9E3779B97F4A7C15 + prompt_index * 0xBF58476D1CE4E5B9 + 0x94D049BB133111EB) & 0x7FFFFFFFFFFFFFFF


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather row shards produced with ``shard_range`` into the full [n_total, ...] tensor on every rank.
    One collective; shards may differ by one row (padded for the fixed-size all_gather, then trimmed)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(out[r][:hi - lo])
    return torch.cat(parts, 0)


def sharded_generate(fn: Callable[[int, int], torch.Tensor], n_prompts: int, group=None) -> torch.Tensor:
    """Run ``fn(lo, hi)`` (which generates prompts [lo, hi) on this rank's device and returns one row per
    prompt) on this rank's shard and gather every rank's rows."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(n_prompts, rank, world)
    return gather_rows(fn(lo, hi), n_prompts, group)
