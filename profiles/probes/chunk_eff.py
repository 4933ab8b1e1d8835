"""Device-resident forward time per clip as a function of the batch (what a chunk of the host pipeline costs)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import videoprism_b200 as vp
m = vp.get_model("videoprism_public_v1_base"); m.load_state(vp.synthetic_state(m, seed=1234))
x = torch.from_numpy(np.random.default_rng(0).random((32, 16, 288, 288, 3), dtype=np.float32)).cuda()
for b in (1, 2, 4, 8, 16, 32):
    n = 32 // b
    for _ in range(2):
        for i in range(n): m(x[i*b:(i+1)*b])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        for i in range(n): m(x[i*b:(i+1)*b])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"32 clips as {n:2d} forwards of {b:2d}: {dt*1e3:7.2f} ms  ({32/dt:6.0f} clips/s)")
