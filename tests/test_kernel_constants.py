"""CPU checks of the numerical constants the CUDA kernels are built on, read from the kernel sources themselves:
the odd polynomials that replace cap*tanh(s/cap) in the tcgen05 attention kernel (layers.py:586-594), its FMA-pipe exp2,
and the tanh form of the exact-erf GELU in the GEMM epilogue (layers.py:31).  The GPU tests hold the kernels to their
end-to-end tolerances; these pin the error budget those tolerances are derived from."""
import math
import os
import re

import numpy as np

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "videoprism-mlx_b200", "csrc")


def _src(name):
    with open(os.path.join(CSRC, name)) as f:
        return f.read()


def _num(pattern, text):
    m = re.search(pattern, text)
    assert m, pattern
    return [float(g) for g in m.groups()]


def test_logit_cap_polynomials():
    src = _src("attention_kloop_tcgen05.cu")
    t1, t2 = _num(r"const double t1 = (-?[0-9.eE+-]+), t2 = (-?[0-9.eE+-]+);", src)
    (u1,) = _num(r"const double u1 = (-?[0-9.eE+-]+);", src)
    assert "p.range = 0.5f * a.cap;" in src and "p.range_lo = a.cap / 5.0f;" in src
    cap = 50.0
    # quintic tier: |s| <= cap / 2
    s = np.linspace(-cap / 2, cap / 2, 200001)
    x = s / cap
    quintic = s * (1.0 + t1 * x ** 2 + t2 * x ** 4)
    err5 = np.abs(quintic - cap * np.tanh(x)).max()
    assert err5 <= 7.0e-4, err5                     # 6.7e-4 in the capped logit = 0.07 % in a softmax weight
    # cubic tier: |s| <= cap / 5
    s = np.linspace(-cap / 5, cap / 5, 200001)
    x = s / cap
    err3 = np.abs(s * (1.0 + u1 * x ** 2) - cap * np.tanh(x)).max()
    assert err3 <= 3.7e-4, err3
    # the tiers are exact in the limit s -> 0 and odd
    assert abs(1e-3 * (1.0 + u1 * (1e-3 / cap) ** 2) - cap * math.tanh(1e-3 / cap)) < 1e-12


def test_fma_pipe_exp2():
    """x = n + f by the 1.5 * 2^23 rounding trick, 2^f by a cubic, n added into the exponent field: emulated in float32."""
    src = _src("attention_kloop_tcgen05.cu")
    c3, c2 = _num(r"fma2\(fr, pk2\(([0-9.]+)f, [0-9.]+f\), pk2\(([0-9.]+)f, [0-9.]+f\)\);", src)
    (c1,) = _num(r"pp = fma2\(pp, fr, pk2\(([0-9.]+)f, [0-9.]+f\)\);\n  pp = fma2\(pp, fr, pk2\(1\.f", src)
    x = np.linspace(-72.2, 72.2, 400001).astype(np.float32)
    magic = np.float32(12582912.0)
    tt = (x + magic).astype(np.float32)
    nn = (tt - magic).astype(np.float32)
    fr = (x - nn).astype(np.float32)
    assert np.abs(fr).max() <= 0.5
    pp = ((np.float32(c3) * fr + np.float32(c2)) * fr + np.float32(c1)) * fr + np.float32(1.0)
    bits = pp.astype(np.float32).view(np.uint32).astype(np.int64) + (tt.view(np.uint32).astype(np.int64) << 23)
    got = (bits & 0xFFFFFFFF).astype(np.uint32).view(np.float32).astype(np.float64)
    want = np.exp2(x.astype(np.float64))
    rel = np.abs(got / want - 1.0).max()
    assert rel <= 1.2e-4, rel                       # a 30th of the bf16 step of P (2^-7)


def test_gelu_tanh_form_of_the_exact_erf_gelu():
    src = _src("ptx.cuh")
    c0, = _num(r"const f32x2 c0 = pk2\(([0-9.]+)f,", src)
    c1, = _num(r"const f32x2 c1 = pk2\(([0-9.]+)f,", src)
    c2, = _num(r"const f32x2 c2 = pk2\((-[0-9.]+)f,", src)
    x = np.linspace(-12.0, 12.0, 480001)
    x2 = np.minimum(x * x, 50.0)
    got = 0.5 * x * (1.0 + np.tanh(x * (c0 + c1 * x2 + c2 * x2 * x2)))
    want = 0.5 * x * (1.0 + np.vectorize(math.erf)(x / math.sqrt(2.0)))
    err = np.abs(got - want).max()
    assert err <= 3.0e-5, err
