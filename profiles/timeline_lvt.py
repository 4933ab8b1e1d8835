"""In-situ kernel timeline of one video-text retrieval step (LvT): video side (encoder, auxiliary encoder, pooler) and
text side, CUDA events after every launch (vp_trace), no profiler attached.

    python profiles/timeline_lvt.py [clips] [queries] [model]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import videoprism_b200 as vp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 128
name = {"base": "videoprism_lvt_public_v1_base", "large": "videoprism_lvt_public_v1_large"}[sys.argv[3] if len(sys.argv) > 3 else "base"]
model = vp.get_model(name)
model.load_state(vp.synthetic_state(model, seed=1234))
video = torch.from_numpy(np.random.default_rng(0).random((B, 16, 288, 288, 3), dtype=np.float32)).cuda()
ids = np.random.default_rng(2).integers(1, 32000, (Q, 64), dtype=np.int32)
lens = np.random.default_rng(3).integers(4, 33, (Q,))
pad = (np.arange(64)[None, :] >= lens[:, None]).astype(np.float32)
ids_t = torch.from_numpy(np.where(pad > 0, 0, ids).astype(np.int32)).cuda()
pad_t = torch.from_numpy(pad).cuda()
for _ in range(3):
    model(video, ids_t, pad_t)
torch.cuda.synchronize()
steps = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    model(video, ids_t, pad_t)
e1.record()
torch.cuda.synchronize()
plain = e0.elapsed_time(e1) / steps
model.trace(True)
for _ in range(steps):
    model(video, ids_t, pad_t)
rows = model.trace_report()
model.trace(False)
print(f"{name}: {B} clips + {Q} queries: step without tracing {plain:.3f} ms = {B / plain * 1e3:.1f} clips/s")
print(f"{'kernel':24s} {'n/step':>7s} {'us each':>9s} {'ms/step':>8s} {'share':>6s}")
total = sum(ms for label, n, ms in rows if label != "TOTAL")
for label, n, ms in rows:
    if label == "TOTAL":
        continue
    print(f"{label:24s} {n / steps:7.1f} {ms / n * 1e3:9.1f} {ms / steps:8.3f} {100 * ms / total:5.1f}%")
