"""VP_ATTN_TRACE=1 python profiles/attn_trace.py: one launch of the key-loop attention kernel per shape with the in-kernel
timeline of CTA 0 printed to stderr (clock64 at the hand-over points between softmax warps, MMA issuer and drain warps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import videoprism_b200._lib as L
lib = L.lib()
D, H = 768, 12
st = int(torch.cuda.current_stream().cuda_stream)
for num_seq, S in ((128, 256), (1, 4096)):
    qkv = torch.randn((num_seq * S, 3 * D), device="cuda"); qkv[:, :D] *= 0.2; qkv = qkv.bfloat16()
    out = torch.zeros((num_seq * S, D), dtype=torch.bfloat16, device="cuda")
    for _ in range(2):
        assert lib.vp_attention(qkv.data_ptr(), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, 3 * D, out.data_ptr(), D, num_seq, S, 1, H, 64, 50.0, None, 0, st) == 0
    torch.cuda.synchronize()
print("ok")
