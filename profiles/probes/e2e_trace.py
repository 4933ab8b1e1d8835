"""Where the host-buffer (e2e) call spends its time: wall clock per call vs the traced per-kernel intervals."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import videoprism_b200 as vp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = vp.get_model("videoprism_public_v1_base"); m.load_state(vp.synthetic_state(m, seed=1234))
x = torch.from_numpy(np.random.default_rng(0).random((B, 16, 288, 288, 3), dtype=np.float32)).pin_memory().numpy()
out = vp.pinned_empty((B, 4096, 768))
for _ in range(3): m(x, out=out)
t0 = time.perf_counter()
for _ in range(5): m(x, out=out)
wall = (time.perf_counter() - t0) / 5
m.trace(True)
t0 = time.perf_counter(); m(x, out=out); w1 = time.perf_counter() - t0
rows = m.trace_report(); m.trace(False)
print(f"B={B}: wall per call {wall*1e3:.2f} ms ({B/wall:.0f} clips/s); traced call {w1*1e3:.2f} ms")
for r in rows: print(f"  {r[0]:<18}{r[1]:>5}{r[2]:>10.3f} ms")
xd = torch.from_numpy(x).cuda()
for _ in range(3): m(xd)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): m(xd)
torch.cuda.synchronize(); print(f"device-resident: {(time.perf_counter()-t0)/5*1e3:.2f} ms per call")
