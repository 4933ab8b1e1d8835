"""Sustained (seconds-long, power-capped) throughput of the four block GEMMs, the same way MEASURED_PEAKS.json's
bf16_tflops_sustained was taken for cuBLAS (back-to-back launches for ~3 s), plus torch.matmul for reference."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import videoprism_b200._lib as L
lib = L.lib()
st = int(torch.cuda.current_stream().cuda_stream)
M, D, F = 131072, 768, 3072
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0

def run(name, fn, flops):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(20): fn()
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:34s} {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s sustained over {secs:.0f} s")

def mk(Mm, Nn, Kk, act, resid):
    A = (torch.randn((Mm, Kk), device="cuda") * 0.5).bfloat16()
    Wt = (torch.randn((Nn, Kk), device="cuda") * 0.02).bfloat16()
    bias = torch.zeros((Nn,), device="cuda")
    C = torch.zeros((Mm, Nn), dtype=torch.bfloat16, device="cuda")
    def f():
        lib.vp_gemm_bf16(A.data_ptr(), Kk, Wt.data_ptr(), Kk, C.data_ptr(), Nn, Mm, Nn, Kk, bias.data_ptr(), act, C.data_ptr() if resid else None,
                         Nn if resid else 0, None, None, 0, 0, st)
    return f, 2.0 * Mm * Nn * Kk, (A, Wt)

for name, args in [("QKV", (M, 3 * D, D, 0, False)), ("out-proj+resid", (M, D, D, 0, True)), ("FFN1+GELU", (M, F, D, 1, False)), ("FFN2+resid", (M, D, F, 0, True))]:
    f, fl, _ = mk(*args)
    run("ours " + name, f, fl)
A = torch.randn((M, D), device="cuda").bfloat16(); W = torch.randn((D, F), device="cuda").bfloat16()
run("torch.matmul (cuBLAS) FFN1 shape", lambda: torch.matmul(A, W), 2.0 * M * D * F)
A = torch.randn((8192, 8192), device="cuda").bfloat16(); W = torch.randn((8192, 8192), device="cuda").bfloat16()
run("torch.matmul (cuBLAS) 8192^3", lambda: torch.matmul(A, W), 2.0 * 8192**3)
