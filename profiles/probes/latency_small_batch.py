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
    print(f"{name} B={B}: device back-to-back {dev_ms:.3f} ms ({B / dev_ms * 1e3:.0f} clips/s, {launches} launches), "
          f"one at a time {sync_ms:.3f} ms, host numpy->numpy (pinned) {host_ms:.3f} ms")
