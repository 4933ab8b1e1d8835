"""A CPU model of the tcgen05 attention kernel's ARITHMETIC (not its schedule): fp32 scores of bf16 operands, the tiered
polynomial logit cap, exp2, the softmax weights TRUNCATED to bfloat16 before the P.V product, the normaliser summed over
exactly those truncated weights, a bfloat16 result.  It shows that the tolerance the GPU tests apply
(`tests/test_kernels_gpu.py:_attention_close`) follows from those two roundings for flat, moderate and saturated logits, i.e.
that the bound is a property of the algorithm and not a number tuned until the kernel passed."""
import re
import os

import numpy as np
import pytest
import torch

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "videoprism-mlx_b200", "csrc")


def _constants():
    with open(os.path.join(CSRC, "attention_kloop_tcgen05.cu")) as f:
        src = f.read()
    t1, t2 = (float(x) for x in re.search(r"const double t1 = (-?[0-9.eE+-]+), t2 = (-?[0-9.eE+-]+);", src).groups())
    u1 = float(re.search(r"const double u1 = (-?[0-9.eE+-]+);", src).group(1))
    return t1, t2, u1


def kernel_model(q, k, v, cap=50.0):
    """q, k, v: [S, dh] bfloat16 tensors (q pre-scaled).  Returns the kernel's result as float32 [S, dh]."""
    t1, t2, u1 = _constants()
    log2e = 1.4426950408889634
    s = q.float() @ k.float().T                                         # tensor core: bf16 operands, fp32 accumulation
    S = s.shape[1]
    x = torch.empty_like(s)
    for c0 in range(0, S, 16):                                          # the cap tier is chosen per row and 16-column chunk
        blk = s[:, c0:c0 + 16]
        amax = blk.abs().amax(dim=1, keepdim=True)
        y = blk / cap
        cubic = blk * (1.0 + u1 * y * y)
        quintic = blk * (1.0 + t1 * y * y + t2 * y ** 4)
        exact = cap * torch.tanh(y)
        x[:, c0:c0 + 16] = torch.where(amax <= cap / 5, cubic, torch.where(amax <= cap / 2, quintic, exact)) * log2e
    p = torch.exp2(x)                                                   # no row maximum: the cap bounds the exponent
    p_bf16 = (p.view(torch.int32) & -65536).view(torch.float32)         # truncation to bfloat16 (PRMT keeps the high halves)
    o = p_bf16 @ v.float()                                              # tensor core again
    l = p_bf16.sum(dim=1, keepdim=True)                                 # P x ones: the same truncated weights
    return (o / l).bfloat16().float()


def reference(q, k, v, cap=50.0):
    s = q.double() @ k.double().T
    p = torch.softmax(cap * torch.tanh(s / cap), dim=-1)
    return (p @ v.double()).float()


@pytest.mark.parametrize("S,qscale", [(256, 0.2), (256, 1.5), (256, 4.0), (256, 12.0), (1024, 1.5), (1024, 12.0)])
def test_model_of_the_kernel_meets_the_derived_bound(S, qscale):
    g = torch.Generator().manual_seed(int(S + qscale * 10))
    max_err, sq_err, sq_ref, vmax = 0.0, 0.0, 0.0, 0.0
    for _ in range(6):                                                  # a few (sequence, head) problems
        q = (torch.randn((S, 64), generator=g) * qscale).bfloat16()
        k = torch.randn((S, 64), generator=g).bfloat16()
        v = torch.randn((S, 64), generator=g).bfloat16()
        got, want = kernel_model(q, k, v), reference(q, k, v)
        max_err = max(max_err, float((got - want).abs().max()))
        sq_err += float((got - want).pow(2).sum()); sq_ref += float(want.pow(2).sum())
        vmax = max(vmax, float(v.float().abs().max()))
    bound = (2.0 ** -7 + 1.0e-3 + 2.0 ** -9) * vmax
    rms_ratio = (sq_err / sq_ref) ** 0.5
    print(f"[model] S={S} qscale={qscale}: max err {max_err:.4g} (bound {bound:.4g}), rms err / rms ref {rms_ratio:.3g} (bound 4e-3)")
    assert max_err <= bound
    assert rms_ratio <= 4e-3


def test_truncated_weights_and_their_normaliser_cancel_for_a_dominant_key():
    """One key carries the row (saturated cap): out = v_key whatever the rounding of its weight did, because the normaliser
    is the sum of the SAME truncated weights."""
    S = 256
    q = torch.zeros((S, 64)); k = torch.zeros((S, 64)); v = torch.randn((S, 64), generator=torch.Generator().manual_seed(1))
    q[:, 0] = 40.0; k[7, 0] = 40.0                                      # logit 1600 on key 7 -> capped to ~50; 0 elsewhere
    got = kernel_model(q.bfloat16(), k.bfloat16(), v.bfloat16())
    assert torch.allclose(got, v.bfloat16().float()[7].expand_as(got), atol=2.0 ** -8 * float(v.abs().max()))
