"""Seed sharding across the GPUs of one box and the final gather.

Generation is embarrassingly parallel: image i depends only on (prompt, seed i) (data_generation.py:56-59), so rank r
of W takes {i : i mod W == r} and there is NO data-path collective.  The only exchange is one all_gather of fixed-size
records at the end (boxes padded to max_boxes, counts, heat maps or u8 stacks) — NCCL over NVLink on the GPU box,
gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def shard_seeds(num_images: int, rank: int, world_size: int) -> List[int]:
    """Seeds (== image indices, data_generation.py:56) owned by `rank`: i mod world_size == rank."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    return list(range(rank, num_images, world_size))


def padded_count(num_images: int, world_size: int) -> int:
    """Every rank gathers the same number of records: ceil(num_images / world_size)."""
    return (num_images + world_size - 1) // world_size


def gather_records(local: Dict[str, torch.Tensor], seeds: Sequence[int], num_images: int,
                   group=None) -> Dict[str, torch.Tensor]:
    """all_gather every tensor of `local` (leading dim = this rank's images, in `seeds` order) and return them
    re-ordered by seed, shape [num_images, ...].  Ranks with fewer images are padded (records beyond num_images are
    dropped after the gather).  Works on CUDA tensors with NCCL and on CPU tensors with gloo."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    per_rank = padded_count(num_images, world)
    out: Dict[str, torch.Tensor] = {}
    for name, t in local.items():
        if t.shape[0] != len(seeds):
            raise ValueError(f"{name}: leading dim {t.shape[0]} != number of local seeds {len(seeds)}")
        pad = per_rank - t.shape[0]
        if pad:
            t = torch.cat([t, t.new_zeros((pad,) + tuple(t.shape[1:]))], 0)
        t = t.contiguous()
        if world == 1:
            gathered = t[None]
        else:
            flat = t.new_empty((world * per_rank,) + tuple(t.shape[1:]))  # concatenated form (gloo and nccl)
            dist.all_gather_into_tensor(flat, t, group=group)
            gathered = flat.view((world, per_rank) + tuple(t.shape[1:]))
        # record (r, k) holds seed r + k*world  ->  seed-major order is the transpose
        merged = gathered.transpose(0, 1).reshape((per_rank * world,) + tuple(t.shape[1:]))
        out[name] = merged[:num_images].contiguous()
    del rank
    return out
