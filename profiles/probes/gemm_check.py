"""Correctness probe of the bf16 GEMM on shapes big enough for the SM-pair / cluster paths (vs torch.matmul)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import videoprism_b200._lib as L
lib = L.lib()
st = int(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)
for (M, N, K) in ((16384, 768, 768), (4096, 3072, 768), (16384 + 256, 2304, 768), (131072, 768, 3072)):
    A = (torch.randn((M, K), device="cuda") * 0.5).bfloat16()
    Wt = (torch.randn((N, K), device="cuda") * 0.05).bfloat16()
    C = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    assert lib.vp_gemm_bf16(A.data_ptr(), K, Wt.data_ptr(), K, C.data_ptr(), N, M, N, K, None, 0, None, 0, None, None, 0, 0, st) == 0
    torch.cuda.synchronize()
    ref = (A.float() @ Wt.float().t())
    err = (C.float() - ref).abs()
    bad = err > 0.05 + 0.02 * ref.abs()
    print(f"{M}x{N}x{K}: max err {err.max().item():.4f}, bad {int(bad.sum())} of {M*N}")
    if bad.any():
        rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
        print("   bad rows (first/last/count):", int(rows[0]), int(rows[-1]), len(rows), " row tiles of 128:", sorted(set((rows // 128).tolist()))[:24])
        print("   bad cols (first/last/count):", int(cols[0]), int(cols[-1]), len(cols), " col groups of 64:", sorted(set((cols // 64).tolist()))[:48])
