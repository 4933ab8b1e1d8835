"""The retrieval step (BASELINE.json configs[3]/[4]) on TWO GPUs over NCCL against the oracle's similarity matrix V.T^T:
clips and text queries sharded over the ranks (shard_range), video-text forward on each rank's shard, ONE all-gather of the
pooled embeddings, similarity matrix on every rank (README.md:81; encoders.py:784-910).  Skipped on a single-GPU box;
the host logic of the same code runs on the CPU with gloo in tests/test_distributed.py."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_CLIPS, N_QUERIES = 6, 10   # even video shards, even text shards at world 2
N_CLIPS_RAGGED, N_QUERIES_RAGGED = 5, 7


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(n_clips, n_queries):
    import videoprism_oracle as O
    cfg = O.tiny_config("clip")
    W = O.make_synthetic_weights(cfg)
    video = O.make_video(n_clips, 4, 16, seed=21, kind="normal")
    ids, pad = O.make_text(n_queries, vocab=cfg["vocabulary_size"], max_len=8)
    return cfg, W, video, ids, pad


def _worker(rank, world, port, n_clips, n_queries, out_dir):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import videoprism_b200 as vp
        from videoprism_b200.retrieval import retrieval_similarity, shard_range
        cfg, W, video, ids, pad = _inputs(n_clips, n_queries)
        model = vp.FactorizedVideoCLIP(**{k: v for k, v in cfg.items() if k != "kind"})
        model.load_state(W)
        lo, hi = shard_range(n_clips, rank, world)
        qlo, qhi = shard_range(n_queries, rank, world)
        v = torch.from_numpy(video[lo:hi]).cuda()
        i = torch.from_numpy(ids[qlo:qhi]).cuda()
        p = torch.from_numpy(pad[qlo:qhi]).cuda()
        sim = retrieval_similarity(model, v, i, p, total_clips=n_clips, total_queries=n_queries)     # one collective
        sim2 = retrieval_similarity(model, v, i, p)                                                   # sizes exchanged first
        torch.cuda.synchronize()
        assert torch.equal(sim, sim2)
        np.save(os.path.join(out_dir, f"sim_{rank}.npy"), sim.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips,n_queries", [(N_CLIPS, N_QUERIES), (N_CLIPS_RAGGED, N_QUERIES_RAGGED)])
def test_retrieval_similarity_two_gpus_nccl(tmp_path, n_clips, n_queries):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the single-GPU form of the same step is test_parity_gpu.py / test_golden_gpu.py)")
    import videoprism_oracle as O
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_clips, n_queries, str(tmp_path)), nprocs=world, join=True)
    sims = [np.load(tmp_path / f"sim_{r}.npy") for r in range(world)]
    assert sims[0].shape == (n_clips, n_queries)
    assert np.array_equal(sims[0], sims[1])                      # identical on every rank
    cfg, W, video, ids, pad = _inputs(n_clips, n_queries)
    want_v, want_t, _ = O.run_clip(cfg, W, video, ids, pad)
    want = want_v @ want_t.T
    err = float(np.abs(sims[0] - want).max())
    print(f"[nccl] retrieval similarity {n_clips} x {n_queries} on 2 GPUs vs oracle V.T^T: max-abs {err:.4g} (|sim| max {np.abs(want).max():.3g})")
    assert err < 2e-2                                            # bf16 path; embeddings are unit vectors


def test_retrieval_similarity_single_gpu_matches_oracle():
    """The same step on one GPU (world 1: the gather is the identity), against the oracle."""
    import torch
    import videoprism_b200 as vp
    import videoprism_oracle as O
    from videoprism_b200.retrieval import retrieval_similarity
    cfg, W, video, ids, pad = _inputs(N_CLIPS, N_QUERIES)
    model = vp.FactorizedVideoCLIP(**{k: v for k, v in cfg.items() if k != "kind"})
    model.load_state(W)
    sim = retrieval_similarity(model, torch.from_numpy(video).cuda(), torch.from_numpy(ids).cuda(), torch.from_numpy(pad).cuda(),
                               total_clips=N_CLIPS, total_queries=N_QUERIES)
    want_v, want_t, _ = O.run_clip(cfg, W, video, ids, pad)
    err = float(np.abs(sim.cpu().numpy() - want_v @ want_t.T).max())
    print(f"[nccl] retrieval similarity on 1 GPU vs oracle: max-abs {err:.4g}")
    assert err < 2e-2
