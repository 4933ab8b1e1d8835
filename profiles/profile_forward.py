"""Profiling harness: one base-encoder forward inside cudaProfilerStart/Stop (use ncu --profile-from-start off).

    python profiles/profile_forward.py [B] [model]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import videoprism_b200 as vp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
name = {"base": "videoprism_public_v1_base", "large": "videoprism_public_v1_large"}[sys.argv[2] if len(sys.argv) > 2 else "base"]
model = vp.get_model(name)
model.load_state(vp.synthetic_state(model))
x = torch.from_numpy(np.random.default_rng(0).random((B, 16, 288, 288, 3), dtype=np.float32)).cuda()
for _ in range(2):
    model(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
model(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", model.kernel_launches)
