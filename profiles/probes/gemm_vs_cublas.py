"""Our tcgen05 GEMM against torch.matmul (cuBLAS) on the shape MEASURED_PEAKS.json is quoted on (8192^3 bf16), both
back to back for ~2 s (power-capped steady state), with the SM clock and power sampled while they run."""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import videoprism_b200._lib as L
lib = L.lib()
st = int(torch.cuda.current_stream().cuda_stream)
def smi():
    o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip().split(",")
    return float(o[0]), float(o[1])
def sustained(fn, flops, secs=2.0):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    n = max(20, int(secs * 1e3 / e0.elapsed_time(e1)))
    for _ in range(n // 2): fn()
    e0.record()
    for _ in range(n // 2): fn()
    e1.record()
    time.sleep(secs * 0.3)
    mhz, w = smi()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (n // 2)
    return flops / ms / 1e9, mhz, w
for (M, N, K) in ((8192, 8192, 8192), (131072, 768, 3072), (131072, 3072, 768)):
    A = (torch.randn((M, K), device="cuda") * 0.5).bfloat16()
    Wt = (torch.randn((N, K), device="cuda") * 0.02).bfloat16()
    C = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    def ours():
        assert lib.vp_gemm_bf16(A.data_ptr(), K, Wt.data_ptr(), K, C.data_ptr(), N, M, N, K, None, 0, None, 0, None, None, 0, 0, st) == 0
    def cublas():
        torch.matmul(A, Wt.t(), out=C)
    fl = 2.0 * M * N * K
    a = sustained(ours, fl); b = sustained(cublas, fl)
    print(f"{M}x{N}x{K}: ours {a[0]:7.1f} TFLOP/s ({a[1]:.0f} MHz, {a[2]:.0f} W)   cuBLAS {b[0]:7.1f} TFLOP/s ({b[1]:.0f} MHz, {b[2]:.0f} W)")
