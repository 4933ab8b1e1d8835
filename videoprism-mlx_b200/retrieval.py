"""Multi-GPU plumbing for the data-parallel path: contiguous batch sharding (no collective inside the encoder)
and the one real exchange step of the video-text retrieval workload, an all-gather of the pooled embeddings
(BASELINE.json configs 4-5; the consumer is README.md:81 / the colab's `compute_similarity_matrix`).

Uses torch.distributed only as plumbing: NCCL over NVLink for CUDA tensors, gloo in the CPU tests."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n units for `rank` of `world`; ragged tails go to the first ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_embeddings(local: torch.Tensor, total: Optional[int] = None, group=None) -> torch.Tensor:
    """All-gathers row shards `[n_local, D]` (in rank order) into `[n_total, D]` on every rank.  Shards may be
    ragged (they follow `shard_range`); a single process returns its input."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    n_max = max(sizes)
    padded = local
    if local.shape[0] < n_max:
        padded = torch.cat([local, local.new_zeros((n_max - local.shape[0],) + tuple(local.shape[1:]))], dim=0)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    out = torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
    if total is not None and out.shape[0] != total:
        raise RuntimeError(f"gathered {out.shape[0]} rows, expected {total}")
    return out


def retrieval_similarity(model, video_shard, ids_shard, paddings_shard, group=None) -> torch.Tensor:
    """One retrieval step on this rank's shard: video-text forward (device buffers), all-gather of the pooled
    embeddings, similarity matrix `[num_clips, num_queries]` (identical on every rank)."""
    from .models import compute_similarity_matrix
    v, t, _ = model(video_shard, ids_shard, paddings_shard)
    v_all = gather_embeddings(v, group=group)
    t_all = gather_embeddings(t, group=group)
    return compute_similarity_matrix(v_all, t_all)
