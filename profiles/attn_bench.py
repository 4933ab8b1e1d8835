"""Spatial (S = 256) and auxiliary (S = 4096) attention kernels through the C ABI: error against an fp32 evaluation of
layers.py:601-661 on the same bf16 operands, and isolated timings (CUDA events).  The kernel variant is chosen by the
environment (VP_ATTN_KERNEL, VP_ATTN_POLY), read once per process, so run one process per variant:

    VP_ATTN_KERNEL=2 VP_ATTN_POLY=4 python profiles/attn_bench.py [clips] [sustained_seconds]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import videoprism_b200._lib as L

lib = L.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
SUST = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
D, H = 768, 12
st = int(torch.cuda.current_stream().cuda_stream)
tag = f"kernel={os.environ.get('VP_ATTN_KERNEL', 'default')} poly={os.environ.get('VP_ATTN_POLY', 'default')}"


def run(qkv, out, num_seq, S):
    rc = lib.vp_attention(qkv.data_ptr(), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, 3 * D, out.data_ptr(), D, num_seq, S, 1, H, 64,
                          50.0, None, 0, st)
    assert rc == 0


def ref(qkv, num_seq, S):
    x = qkv.float().reshape(num_seq, S, 3, H, 64)
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]
    logits = torch.einsum("btnh,bsnh->bnts", q, k)
    logits = 50.0 * torch.tanh(logits / 50.0)
    return torch.einsum("bnts,bsnh->btnh", torch.softmax(logits, -1), v).reshape(num_seq * S, D)


def errors(num_seq, S, qscale):
    g = torch.Generator(device="cuda").manual_seed(S + int(qscale * 10))
    qkv = torch.randn((num_seq * S, 3 * D), device="cuda", generator=g)
    qkv[:, :D] *= qscale
    qkv = qkv.bfloat16()
    out = torch.zeros((num_seq * S, D), dtype=torch.bfloat16, device="cuda")
    run(qkv, out, num_seq, S)
    torch.cuda.synchronize()
    r = ref(qkv, num_seq, S)
    d = (out.float() - r)
    vmax = float(qkv[:, 2 * D:].float().abs().max())
    return float(d.abs().max()), float(d.pow(2).mean().sqrt()), float(r.pow(2).mean().sqrt()), vmax


def timeit(num_seq, S, qscale=0.2, reps=20):
    qkv = torch.randn((num_seq * S, 3 * D), device="cuda")
    qkv[:, :D] *= qscale
    qkv = qkv.bfloat16()
    out = torch.zeros((num_seq * S, D), dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        run(qkv, out, num_seq, S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if SUST > 0:
        e0.record(); run(qkv, out, num_seq, S); e1.record(); torch.cuda.synchronize()
        reps = max(20, int(SUST * 1e3 / max(e0.elapsed_time(e1), 1e-3)))
        for _ in range(reps):
            run(qkv, out, num_seq, S)
    e0.record()
    for _ in range(reps):
        run(qkv, out, num_seq, S)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 4.0 * S * S * 64 * H * num_seq / ms / 1e9


for (ns, S, qs) in [(8, 256, 0.2), (8, 256, 1.0), (8, 256, 3.0), (1, 4096, 0.2), (1, 4096, 1.0)]:
    mx, rms, rr, vmax = errors(ns, S, qs)
    print(f"[{tag}] error S={S} qscale={qs}: max-abs {mx:.5f} rms {rms:.6f} (ref rms {rr:.4f}, max|v| {vmax:.2f}, bound 2^-7 max|v| = {vmax / 128:.4f})")
for qs in (0.2, 1.0):
    ms, tf = timeit(B * 16, 256, qs)
    print(f"[{tag}] spatial S=256, {B} clips, qscale={qs}: {ms * 1e3:8.1f} us  {tf:7.1f} TFLOP/s")
ms, tf = timeit(max(B // 8, 1), 4096, 0.2)
print(f"[{tag}] auxiliary S=4096, {max(B // 8, 1)} clips: {ms * 1e3:8.1f} us  {tf:7.1f} TFLOP/s")
ms, tf = timeit(max(B // 8, 1), 4096, 1.0)
print(f"[{tag}] auxiliary S=4096, {max(B // 8, 1)} clips, qscale=1.0: {ms * 1e3:8.1f} us  {tf:7.1f} TFLOP/s")
if os.environ.get("VP_AB_SWEEP"):
    for clips in (4, 8, 16, 32, 64):
        ms, tf = timeit(clips * 16, 256, 0.2)
        print(f"[{tag}] sweep spatial S=256, {clips} clips: {ms * 1e3:8.1f} us  {tf:7.1f} TFLOP/s")
    for clips in (1, 2, 4, 8, 16):
        ms, tf = timeit(clips, 4096, 0.2)
        print(f"[{tag}] sweep auxiliary S=4096, {clips} clips: {ms * 1e3:8.1f} us  {tf:7.1f} TFLOP/s")
