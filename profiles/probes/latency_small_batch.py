"""Single-clip latency of the encoder forward (BASELINE.json configs[0] is the reference's B = 1 case):
device-resident (CUDA events) and through the host entry point (numpy in, numpy out, wall clock).

    python profiles/probes/latency_small_batch.py [model]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import videoprism_b200 as vp

name = {"base": "videoprism_public_v1_base", "large": "videoprism_public_v1_large"}[sys.argv[1] if len(sys.argv) > 1 else "base"]
model = vp.get_model(name)
model.load_state(vp.synthetic_state(model, seed=1234))
for B in (1, 2, 4):
    v = np.random.default_rng(0).random((B, 16, 288, 288, 3), dtype=np.float32)
    vd = torch.from_numpy(v).cuda()
    for _ in range(5):
        model(vd)
    torch.cuda.synchronize()
    n0 = model.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        model(vd)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / reps
    launches = (model.kernel_launches - n0) // reps
    # one forward at a time with a sync in between: what a latency-sensitive caller sees
    t0 = time.perf_counter()
    for _ in range(reps):
        model(vd)
        torch.cuda.synchronize()
    sync_ms = (time.perf_counter() - t0) / reps * 1e3
    out = vp.pinned_empty((B, 4096, model.config["model_dim"]))
    vp_in = vp.pinned_empty(v.shape); vp_in[...] = v
    for _ in range(3):
        model(vp_in, out=out)
    t0 = time.perf_counter()
    for _ in range(20):
        model(vp_in, out=out)
    host_ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f"[graphs {'off' if os.environ.get('VP_GRAPHS') == '0' else 'on'}] {name} B={B}: device back-to-back {dev_ms:.3f} ms ({B / dev_ms * 1e3:.0f} clips/s, {launches} launches), "
          f"one at a time {sync_ms:.3f} ms, host numpy->numpy (pinned) {host_ms:.3f} ms")

# the reference's own benchmark shape (scripts/benchmark_performance.py:70-94): video-text model, 1 clip + 3 text queries
if len(sys.argv) <= 1 or sys.argv[1] == "base":
    lvt = vp.get_model("videoprism_lvt_public_v1_base")
    lvt.load_state(vp.synthetic_state(lvt, seed=1234))
    v1 = torch.from_numpy(np.random.default_rng(0).random((1, 16, 288, 288, 3), dtype=np.float32)).cuda()
    ids = torch.from_numpy(np.random.default_rng(2).integers(1, 32000, (3, 64), dtype=np.int32)).cuda()
    pad = torch.zeros((3, 64), device="cuda"); pad[:, 12:] = 1.0
    ve = torch.empty((1, 768), device="cuda"); te = torch.empty((3, 768), device="cuda")
    import videoprism_b200._lib as _L
    lib = _L.lib()
    h = lvt._ensure_handle()
    st = int(torch.cuda.current_stream().cuda_stream)

    def lvt_pass():   # fixed output buffers: the call sequence a serving loop makes (and the one the graph cache keys on)
        assert lib.vp_clip_video_forward(h, v1.data_ptr(), 1, 16, 288, 288, None, 1, ve.data_ptr(), None, None, None, st) == 0
        assert lib.vp_clip_text_forward(h, ids.data_ptr(), pad.data_ptr(), 3, 64, 1, te.data_ptr(), st) == 0
    for _ in range(5):
        lvt_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        lvt_pass()
    e1.record()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        lvt_pass()
        torch.cuda.synchronize()
    sync_ms = (time.perf_counter() - t0) / 50 * 1e3
    print(f"videoprism_lvt_public_v1_base 1 clip + 3 queries: device back-to-back {e0.elapsed_time(e1) / 50:.3f} ms, one at a time {sync_ms:.3f} ms "
          f"(graphs {'off' if os.environ.get('VP_GRAPHS') == '0' else 'on'})")
