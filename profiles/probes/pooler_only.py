"""One video-text forward (video side) for `ncu -k regex:pool_` : per-kernel durations of the pooling head."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import videoprism_b200 as vp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = vp.get_model("videoprism_lvt_public_v1_base")
m.load_state(vp.synthetic_state(m, seed=1234))
v = torch.from_numpy(np.random.default_rng(0).random((B, 16, 288, 288, 3), dtype=np.float32)).cuda()
m(v, None, None)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m(v, None, None)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
