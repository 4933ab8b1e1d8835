"""Per-kernel micro-benchmarks through the C ABI (CUDA events, isolated kernels, base-model shapes at B clips).

    python profiles/kernel_bench.py [B] [model]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import videoprism_b200._lib as L

lib = L.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model = sys.argv[2] if len(sys.argv) > 2 else "base"
D, H, F = (768, 12, 3072) if model == "base" else (1024, 16, 4096)
T, N = 16, 256
M = B * T * N
st = int(torch.cuda.current_stream().cuda_stream)


SUSTAINED = float(os.environ.get("VP_KB_SUSTAINED_S", "0"))   # > 0: run every kernel back to back for this many seconds
                                                                # (power-capped steady state) and sample clocks / power


def _smi():
    import subprocess
    out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                         capture_output=True, text=True).stdout.strip().split(",")
    return float(out[0]), float(out[1])


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if SUSTAINED > 0:
        import time
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        n = max(20, int(SUSTAINED * 1e3 / max(e0.elapsed_time(e1), 1e-3)))
        for _ in range(n // 2):   # reach the steady state first
            fn()
        e0.record()
        for _ in range(n // 2):
            fn()
        e1.record()
        time.sleep(SUSTAINED * 0.25)
        mhz, watts = _smi()      # sampled while the second half is still running
        torch.cuda.synchronize()
        print(f"      [sustained {n // 2} launches: {mhz:.0f} MHz, {watts:.0f} W]", end=" ")
        return e0.elapsed_time(e1) / (n // 2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gemm(Mm, Nn, Kk, act=0, resid=False):
    A = (torch.randn((Mm, Kk), device="cuda") * 0.5).bfloat16()
    Wt = (torch.randn((Nn, Kk), device="cuda") * 0.02).bfloat16()
    bias = torch.zeros((Nn,), device="cuda")
    C = torch.zeros((Mm, Nn), dtype=torch.bfloat16, device="cuda")
    def f():
        rc = lib.vp_gemm_bf16(A.data_ptr(), Kk, Wt.data_ptr(), Kk, C.data_ptr(), Nn, Mm, Nn, Kk, bias.data_ptr(), act,
                              C.data_ptr() if resid else None, Nn if resid else 0, None, None, 0, 0, st)
        assert rc == 0
    ms = timeit(f)
    return ms, 2.0 * Mm * Nn * Kk / ms / 1e9


def gemm_ln(Mm, Nn, Kk, act=0, resid=False, stats_out=False):
    """The engine's LayerNorm-folded form: raw A rows + per-row (sum, sumsq) slots + column sums (vp_gemm_bf16_ln)."""
    A = (torch.randn((Mm, Kk), device="cuda") * 0.5).bfloat16()
    Wt = (torch.randn((Nn, Kk), device="cuda") * 0.02).bfloat16()
    bias = torch.zeros((Nn,), device="cuda")
    colsum = Wt.float().sum(1).contiguous()
    slots = lib.vp_gemm_stats_slots(Kk)
    stats = torch.zeros((Mm, slots, 2), device="cuda")
    stats[:, 0, 0] = A.float().sum(1)
    stats[:, 0, 1] = (A.float() ** 2).sum(1)
    C = torch.zeros((Mm, Nn), dtype=torch.bfloat16, device="cuda")
    so = torch.zeros((Mm, lib.vp_gemm_stats_slots(Nn), 2), device="cuda") if stats_out else None
    def f():
        rc = lib.vp_gemm_bf16_ln(A.data_ptr(), Kk, Wt.data_ptr(), Kk, C.data_ptr(), Nn, Mm, Nn, Kk, bias.data_ptr(), act,
                                 C.data_ptr() if resid else None, Nn if resid else 0, stats.data_ptr(), slots, colsum.data_ptr(), Kk,
                                 so.data_ptr() if stats_out else None, st)
        assert rc == 0
    ms = timeit(f)
    return ms, 2.0 * Mm * Nn * Kk / ms / 1e9


def attention(num_seq, S, group, force=0, qscale=0.2):
    qkv = torch.randn((num_seq * S, 3 * D), device="cuda")
    qkv[:, :D] *= qscale      # |logits| of a few units (fast path of the cap); qscale=1 gives |s| > cap/2 rows (MUFU.TANH path)
    qkv = qkv.bfloat16()
    out = torch.zeros((num_seq * S, D), dtype=torch.bfloat16, device="cuda")
    def f():
        rc = lib.vp_attention(qkv.data_ptr(), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, 3 * D, out.data_ptr(), D, num_seq, S, group, H, 64,
                              50.0, None, force * 2, st)
        assert rc == 0
    ms = timeit(f)
    return ms, 4.0 * S * S * 64 * H * num_seq / ms / 1e9


def layernorm():
    x = torch.randn((M, D), device="cuda").bfloat16()
    y = torch.empty_like(x)
    g = torch.ones((D,), device="cuda"); b = torch.zeros((D,), device="cuda")
    def f():
        assert lib.vp_layernorm(x.data_ptr(), D, g.data_ptr(), b.data_ptr(), y.data_ptr(), None, None, 1, 1, M, D, st) == 0
    ms = timeit(f)
    return ms, 2.0 * M * D * 2 / ms / 1e6


print(f"B={B} model={model} M={M}")
for name, args in [("QKV   ", (M, 3 * D, D, 0, False)), ("outprj", (M, D, D, 0, True)), ("FFN1  ", (M, F, D, 1, False)), ("FFN2  ", (M, D, F, 0, True))]:
    ms, tf = gemm(*args)
    print(f"gemm {name} {args[0]}x{args[1]}x{args[2]}: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s")
ms, tf = gemm_ln(M, 3 * D, D)
print(f"gemm QKV    LN-folded            : {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s")
ms, tf = gemm_ln(M, F, D, act=1)
print(f"gemm FFN1   LN-folded + GELU     : {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s")
ms, tf = gemm_ln(M, D, F, resid=True, stats_out=True)
print(f"gemm FFN2   resid + stats_out    : {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s")
ms, tf = gemm_ln(M, D, D, resid=True, stats_out=True)
print(f"gemm outprj resid + stats_out    : {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s")
ms, tf = attention(B * T, N, 1)
print(f"attention spatial (tcgen05): {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s")
ms, tf = attention(B * T, N, 1, qscale=1.0)
print(f"attention spatial (tcgen05, large logits -> tanh path): {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s")
ms, tf = attention(B * T, N, 1, force=1)
print(f"attention spatial (mma.sync): {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s")
ms, tf = attention(B * N, T, N)
print(f"attention temporal: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s   ({(4.0*M*D*2)/ms/1e6:.0f} GB/s algorithmic)")
ms, gbs = layernorm()
print(f"layernorm [{M},{D}]: {ms*1e3:8.1f} us  {gbs:7.0f} GB/s")
if model == "base":
    ms, tf = attention(B // 8 if B >= 8 else 1, T * N, 1)
    print(f"attention auxiliary S=4096 (tcgen05 key loop), {max(B // 8, 1)} clips: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s")
    ms, tf = attention(B // 8 if B >= 8 else 1, T * N, 1, force=1)
    print(f"attention auxiliary S=4096 (mma.sync flash), {max(B // 8, 1)} clips: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s")
# frame ingest: 8 clips of 16 frames, 640x360 -> 288x288 centre crop (HBM-bound: source window read once + output)
fr = torch.randint(0, 256, (8 * 16, 360, 640, 3), dtype=torch.uint8, device="cuda")
dst = torch.empty((8 * 16, 288, 288, 3), dtype=torch.uint8, device="cuda")
def f_ingest():
    assert lib.vp_resize_frames_u8(fr.data_ptr(), 8 * 16, 360, 640, dst.data_ptr(), 288, 0, st) == 0
ms = timeit(f_ingest)
algo = 8 * 16 * (360 * 360 * 3 + 288 * 288 * 3)   # the cropped source window (360 rows x 360 columns: x in [140, 500)) + the output
print(f"frame ingest 8x16x360x640 -> 288^2: {ms*1e3:8.1f} us  {algo / ms / 1e6:7.0f} GB/s algorithmic")
