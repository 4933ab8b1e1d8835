"""Multi-GPU plumbing for the data-parallel path: contiguous batch sharding (no collective inside the encoder)
and the one real exchange step of the video-text retrieval workload, an all-gather of the pooled embeddings
(BASELINE.json configs 4-5; the consumer is README.md:81 / the colab's `compute_similarity_matrix`).

Uses torch.distributed only as plumbing: NCCL over NVLink for CUDA tensors, gloo in the CPU tests."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n units for `rank` of `world`; ragged tails go to the first ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group) -> int:
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


def _all_gather_rows(padded: torch.Tensor, world: int, group) -> torch.Tensor:
    """[n, D] on every rank -> [world, n, D]; ONE collective, no host synchronisation."""
    out = padded.new_empty((world,) + tuple(padded.shape))
    try:
        dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):   # a backend without the flat form: the list form moves the same bytes
        dist.all_gather(list(out.unbind(0)), padded.contiguous(), group=group)
    return out


def gather_embeddings(local: torch.Tensor, total: Optional[int] = None, group=None) -> torch.Tensor:
    """All-gathers row shards `[n_local, D]` (in rank order) into `[n_total, D]` on every rank; a single process returns
    its input.  With `total` given the shard sizes follow from `shard_range` on every rank, so the gather is ONE
    collective with no size exchange and no device-to-host synchronisation (ragged shards are padded to the largest);
    without it the sizes are exchanged first (one more collective and a host read)."""
    world = _world(group)
    if world == 1:
        if total is not None and local.shape[0] != total:
            raise RuntimeError(f"gathered {local.shape[0]} rows, expected {total}")
        return local
    if total is not None:
        sizes = [hi - lo for lo, hi in (shard_range(total, r, world) for r in range(world))]
        rank = dist.get_rank(group)
        if local.shape[0] != sizes[rank]:
            raise RuntimeError(f"rank {rank} holds {local.shape[0]} rows, shard_range({total}, {rank}, {world}) says {sizes[rank]}")
    else:
        n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        got = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(got, n_local, group=group)
        sizes = [int(s.item()) for s in got]
    n_max = max(sizes)
    padded = local
    if local.shape[0] < n_max:
        padded = torch.cat([local, local.new_zeros((n_max - local.shape[0],) + tuple(local.shape[1:]))], dim=0)
    parts = _all_gather_rows(padded, world, group)
    if min(sizes) == n_max:
        return parts.reshape((world * n_max,) + tuple(local.shape[1:]))
    return torch.cat([parts[r, :s] for r, s in enumerate(sizes)], dim=0)


def gather_embedding_pair(v_local: torch.Tensor, t_local: torch.Tensor, total_v: int, total_t: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The retrieval step's exchange as ONE collective: every rank contributes one `[nv_max + nt_max, D]` buffer holding
    its video rows followed by its text rows (shards as `shard_range` cuts them; ragged ones zero-padded), one
    `all_gather_into_tensor`, then both matrices are sliced out.  No size exchange, no host synchronisation."""
    world = _world(group)
    if world == 1:
        return v_local, t_local
    if v_local.shape[1:] != t_local.shape[1:] or v_local.dtype != t_local.dtype:
        return gather_embeddings(v_local, total_v, group), gather_embeddings(t_local, total_t, group)
    vs = [hi - lo for lo, hi in (shard_range(total_v, r, world) for r in range(world))]
    ts = [hi - lo for lo, hi in (shard_range(total_t, r, world) for r in range(world))]
    rank = dist.get_rank(group)
    if v_local.shape[0] != vs[rank] or t_local.shape[0] != ts[rank]:
        raise RuntimeError(f"rank {rank} holds {v_local.shape[0]} video / {t_local.shape[0]} text rows, expected {vs[rank]} / {ts[rank]}")
    nv, nt = max(vs), max(ts)
    buf = v_local.new_zeros((nv + nt,) + tuple(v_local.shape[1:]))
    buf[: v_local.shape[0]] = v_local
    buf[nv: nv + t_local.shape[0]] = t_local
    parts = _all_gather_rows(buf, world, group)
    if min(vs) == nv and min(ts) == nt:
        return (parts[:, :nv].reshape((world * nv,) + tuple(v_local.shape[1:])),
                parts[:, nv:].reshape((world * nt,) + tuple(t_local.shape[1:])))
    v_all = torch.cat([parts[r, :s] for r, s in enumerate(vs)], dim=0)
    t_all = torch.cat([parts[r, nv: nv + s] for r, s in enumerate(ts)], dim=0)
    return v_all, t_all


def retrieval_similarity(model, video_shard, ids_shard, paddings_shard, group=None, total_clips: Optional[int] = None,
                         total_queries: Optional[int] = None) -> torch.Tensor:
    """One retrieval step on this rank's shard: video-text forward (device buffers), all-gather of the pooled
    embeddings, similarity matrix `[num_clips, num_queries]` (identical on every rank).  With the global counts given
    (shards cut by `shard_range`) the exchange is a single collective without any host synchronisation."""
    from .models import compute_similarity_matrix
    v, t, _ = model(video_shard, ids_shard, paddings_shard)
    if total_clips is not None and total_queries is not None:
        v_all, t_all = gather_embedding_pair(v, t, total_clips, total_queries, group=group)
    else:
        v_all = gather_embeddings(v, group=group)
        t_all = gather_embeddings(t, group=group)
    return compute_similarity_matrix(v_all, t_all)
