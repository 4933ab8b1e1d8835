"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous sharding and the embedding all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from videoprism_b200.retrieval import gather_embedding_pair, gather_embeddings, shard_range


def test_shard_range_is_a_contiguous_partition():
    for n in (0, 1, 7, 32, 256, 1024):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_range(32, r, 8) for r in (0, 7)] == [(0, 4), (28, 32)]   # BASELINE config 2: 4 clips per GPU
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_video, n_text, dim, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full_v = torch.arange(n_video * dim, dtype=torch.float32).reshape(n_video, dim) / 7.0
        full_t = torch.arange(n_text * dim, dtype=torch.float32).reshape(n_text, dim) / 3.0 - 5.0
        lo, hi = shard_range(n_video, rank, world)
        v_all = gather_embeddings(full_v[lo:hi].clone(), total=n_video)
        lo, hi = shard_range(n_text, rank, world)
        t_all = gather_embeddings(full_t[lo:hi].clone(), total=n_text)
        ok = torch.equal(v_all, full_v) and torch.equal(t_all, full_t)
        # the same gather without the global count (sizes exchanged first) and as ONE collective for both matrices
        lo_v, hi_v = shard_range(n_video, rank, world)
        ok = ok and torch.equal(gather_embeddings(full_v[lo_v:hi_v].clone()), full_v)
        v2, t2 = gather_embedding_pair(full_v[lo_v:hi_v].clone(), full_t[lo:hi].clone(), n_video, n_text)
        ok = ok and torch.equal(v2, full_v) and torch.equal(t2, full_t)
        try:   # a shard that does not follow shard_range is refused, not silently misplaced
            gather_embeddings(full_v[: hi_v - lo_v + 1].clone(), total=n_video)
            ok = False
        except RuntimeError:
            pass
        sim = v_all @ t_all.T
        np.save(os.path.join(out_dir, f"sim_{rank}.npy"), sim.numpy())
        with open(os.path.join(out_dir, f"ok_{rank}"), "w") as f:
            f.write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_video,n_text", [(2, 8, 16), (2, 7, 5), (3, 7, 5)])   # even and ragged shards; 3 ranks
def test_all_gather_of_sharded_embeddings_world2(tmp_path, world, n_video, n_text):
    dim = 24
    mp.spawn(_worker, args=(world, _free_port(), n_video, n_text, dim, str(tmp_path)), nprocs=world, join=True)
    assert all(open(tmp_path / f"ok_{r}").read() == "1" for r in range(world))
    sims = [np.load(tmp_path / f"sim_{r}.npy") for r in range(world)]
    assert sims[0].shape == (n_video, n_text) and np.array_equal(sims[0], sims[1])   # identical on every rank


def test_single_process_gather_is_identity():
    x = torch.randn(5, 8)
    assert gather_embeddings(x) is x
