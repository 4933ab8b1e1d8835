"""Runs only the spatial attention kernel (B clips worth of frames), for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import videoprism_b200._lib as L
lib = L.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
qs = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
D, H = 768, 12
num_seq = B * 16
qkv = torch.randn((num_seq * 256, 3 * D), device="cuda"); qkv[:, :D] *= qs; qkv = qkv.bfloat16()
out = torch.zeros((num_seq * 256, D), dtype=torch.bfloat16, device="cuda")
st = int(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    assert lib.vp_attention(qkv.data_ptr(), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, 3 * D, out.data_ptr(), D, num_seq, 256, 1, H, 64, 50.0, None, 0, st) == 0
torch.cuda.synchronize()
print("ok")
