"""Runs one GEMM shape through the C ABI a few times (for ncu captures).  python profiles/gemm_only.py M N K act resid"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import videoprism_b200._lib as L
lib = L.lib()
M, N, K, act, resid = (int(x) for x in sys.argv[1:6])
A = (torch.randn((M, K), device="cuda") * 0.5).bfloat16()
Wt = (torch.randn((N, K), device="cuda") * 0.02).bfloat16()
bias = torch.zeros((N,), device="cuda")
C = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
st = int(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    assert lib.vp_gemm_bf16(A.data_ptr(), K, Wt.data_ptr(), K, C.data_ptr(), N, M, N, K, bias.data_ptr(), act, C.data_ptr() if resid else None,
                            N if resid else 0, None, None, 0, 0, st) == 0
torch.cuda.synchronize()
print("ok")
