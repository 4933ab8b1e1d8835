"""Full-size giant encoder (models.py:105-115: D = 1408, 16 heads of 88, F = 6144, 40 + 4 blocks, ~1.0 B parameters; no
released checkpoint): one forward on random-init weights, finite outputs and throughput.  dim_per_head = 88 runs on the
generic mma.sync attention kernels (zero-padded 128-wide head tiles), N = 1408 on 128-column single-SM GEMM tiles.

    python profiles/probes/giant_forward.py [B] [--parity]     (--parity: clip 0 against the fp32 CPU oracle, ~1 min of CPU)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import videoprism_b200 as vp

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
m = vp.models.videoprism_v1_giant()
m.load_state(vp.synthetic_state(m, seed=1))
v = torch.from_numpy(np.random.default_rng(0).random((B, 16, 288, 288, 3), dtype=np.float32)).cuda()
out, _ = m(v)
torch.cuda.synchronize()
assert out.shape == (B, 4096, 1408) and bool(torch.isfinite(out).all())
one, _ = m(v[:1])
assert torch.equal(one[0], out[0])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    m(v)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
D, F, Ls, Lt = 1408, 6144, 40, 4
gf = (2 * 972 * D + (Ls + Lt) * (8 * D * D + 4 * D * F) + Ls * 4 * 256 * D + Lt * 4 * 16 * D) * 4096 / 1e9
print(f"giant encoder B={B}: {ms:.1f} ms/forward = {B / ms * 1e3:.1f} clips/s = {B / ms * gf:.0f} TFLOP/s ({gf:.0f} GF/clip), "
      f"per-token variance {float(out.float().var(-1).mean()):.3f}")

if "--parity" in sys.argv:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import videoprism_oracle as O
    cfg = dict(kind="encoder", **{k: x for k, x in m.config.items() if k != "scan"})
    W = vp.synthetic_state(m, seed=1)
    want, _ = O.run_encoder(cfg, W, v[:1].cpu().numpy())
    got = out[:1].cpu().numpy()
    a, b = got.reshape(-1, 1408).astype(np.float64), want.reshape(-1, 1408).astype(np.float64)
    cos = (a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1))
    print(f"giant encoder, clip 0 vs fp32 oracle: min per-token cosine {cos.min():.6f}, max-abs {np.abs(got - want).max():.4g} "
          f"(ref max {np.abs(want).max():.4g})")
    assert cos.min() >= 0.999
