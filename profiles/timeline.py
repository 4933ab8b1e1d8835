"""In-situ kernel timeline of the encoder forward (CUDA events after every launch, no profiler attached).

    python profiles/timeline.py [B] [model] [steps]

Unlike the ncu launch list (cold cache, serialised, unthrottled clocks) these are the durations the kernels have
INSIDE a sustained run: warm L2, power-capped clocks, back-to-back launches.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import videoprism_b200 as vp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model_name = {"base": "videoprism_public_v1_base", "large": "videoprism_public_v1_large"}[sys.argv[2] if len(sys.argv) > 2 else "base"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
model = vp.get_model(model_name)
model.load_state(vp.synthetic_state(model, seed=1234))
cfg = model.config
D, F, H = cfg["model_dim"], cfg["mlp_dim"], cfg["num_heads"]
T, N = 16, 256
M = B * T * N
video = torch.from_numpy(np.random.default_rng(0).random((B, T, 288, 288, 3), dtype=np.float32)).cuda()
bufs = [video, video.clone()]
for i in range(5):
    model(bufs[i % 2])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    model(bufs[i % 2])
e1.record()
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / steps
model.trace(True)
for i in range(steps):
    model(bufs[i % 2])
rows = model.trace_report()
model.trace(False)

flops = {  # algorithmic flop per launch
    "patch_proj": 2.0 * M * 972 * D,
    "qkv": 2.0 * M * D * 3 * D, "outproj": 2.0 * M * D * D, "ffn1": 2.0 * M * D * F, "ffn2": 2.0 * M * D * F,
    "spatial.attn": 4.0 * N * N * 64 * H * B * T, "temporal.attn": 4.0 * T * T * 64 * H * B * N,
}
total = [r for r in rows if r[0] == "TOTAL"][0]
print(f"{model_name} B={B}: forward without tracing {plain_ms:.3f} ms = {B / plain_ms * 1e3:.1f} clips/s; "
      f"traced sum {total[2] / steps:.3f} ms over {total[1] // steps} launches")
print(f"{'kernel':<20}{'n/fwd':>6}{'us each':>10}{'ms/fwd':>9}{'share':>8}{'TFLOP/s':>10}")
for label, n, ms in rows:
    if label == "TOTAL":
        continue
    key = label if label in flops else label.split(".")[-1]
    f = flops.get(key)
    each = ms / n
    tf = f"{f / each / 1e9:10.1f}" if f else f"{'':>10}"
    print(f"{label:<20}{n // steps:>6}{each * 1e3:>10.1f}{ms / steps:>9.3f}{ms / total[2] * 100:>7.1f}%{tf}")
