"""The numpy stand-ins of jax / flax primitives (oracle/refshim) that let the reference's modules run here are this
repo's code, so they are held to independent implementations of the same public definitions: PyTorch's softmax, exact
(erf) GELU, softplus, rsqrt, one-hot, and its antialiased bilinear resize (the same triangle-filter, half-pixel-centre
definition as jax.image.resize(method='bilinear'), which the reference uses for position tables, encoders.py:124,:157).
The oracle's own resize restatement is held to the same."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(os.path.dirname(HERE), "oracle", "refshim")


@pytest.fixture(scope="module")
def shim():
    """Imports the stand-in `jax` without leaving it importable as `jax` for other test modules."""
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "jax" or k.startswith("jax.")}
    sys.path.insert(0, SHIM)
    try:
        import jax as shim_jax
        yield shim_jax
    finally:
        sys.path.remove(SHIM)
        for k in [k for k in sys.modules if k == "jax" or k.startswith("jax.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_elementwise_and_softmax_against_torch(shim):
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((7, 5, 33)) * 4).astype(np.float32)
    t = torch.from_numpy(x)
    np.testing.assert_allclose(np.asarray(shim.nn.softmax(x, axis=-1)), torch.softmax(t, -1).numpy(), rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(np.asarray(shim.nn.softmax(x, axis=1)), torch.softmax(t, 1).numpy(), rtol=2e-6, atol=1e-8)
    # atol: for x < -4 the fp32 sum 1 + erf(x / sqrt 2) cancels to a few ulps of 1 in either library (|gelu| ~ 1e-6 there)
    np.testing.assert_allclose(np.asarray(shim.nn.gelu(x, approximate=False)), F.gelu(t, approximate="none").numpy(), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(np.asarray(shim.nn.gelu(x, approximate=True)), F.gelu(t, approximate="tanh").numpy(), rtol=2e-6, atol=2e-6)
    x64 = x.astype(np.float64)
    np.testing.assert_allclose(np.asarray(shim.nn.gelu(x64, approximate=False)), F.gelu(torch.from_numpy(x64), approximate="none").numpy(), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(np.asarray(shim.nn.softplus(x)), F.softplus(t, threshold=1e9).numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_array_equal(np.asarray(shim.nn.relu(x)), torch.relu(t).numpy())
    pos = np.abs(x) + 0.1
    np.testing.assert_allclose(np.asarray(shim.lax.rsqrt(pos)), torch.rsqrt(torch.from_numpy(pos)).numpy(), rtol=2e-6)
    ids = rng.integers(0, 11, (4, 6))
    np.testing.assert_array_equal(np.asarray(shim.nn.one_hot(ids, 11)), F.one_hot(torch.from_numpy(ids), 11).float().numpy())
    # a fully masked row (all logits equal to the mask constant) is uniform, as layers.py:155-179 relies on
    masked = np.full((2, 9), -0.7 * np.finfo(np.float32).max, np.float32)
    np.testing.assert_allclose(np.asarray(shim.nn.softmax(masked)), 1.0 / 9, rtol=1e-6)


def _torch_resize_1d(emb: torch.Tensor, n_out: int) -> torch.Tensor:
    """emb [n_in, D] -> [n_out, D] with torch's antialiased bilinear; the un-resized axis is given width 3 because
    torch special-cases a degenerate width of 1."""
    n_in, d = emb.shape
    x = emb.T.reshape(1, d, n_in, 1).expand(1, d, n_in, 3).contiguous()
    return F.interpolate(x, size=(n_out, 3), mode="bilinear", align_corners=False, antialias=True)[0, :, :, 1].T


@pytest.mark.parametrize("n_in,n_out", [(16, 4), (16, 5), (4, 8), (8, 16), (16, 7), (5, 13), (16, 8), (16, 12)])
def test_resize_1d_against_torch_antialiased_bilinear(shim, n_in, n_out):
    import videoprism_oracle as O
    emb = torch.randn(n_in, 6, dtype=torch.float64, generator=torch.Generator().manual_seed(n_in * 31 + n_out))
    want = _torch_resize_1d(emb, n_out)
    got_shim = np.asarray(shim.image.resize(emb.numpy(), (n_out, 6), method="bilinear"))
    np.testing.assert_allclose(got_shim, want.numpy(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(O.interpolate_emb_1d(emb, n_out).numpy(), want.numpy(), rtol=0, atol=1e-12)


@pytest.mark.parametrize("src,dst", [((16, 16), (4, 4)), ((4, 4), (8, 8)), ((16, 16), (12, 12)), ((16, 16), (8, 12)), ((16, 16), (18, 18))])
def test_resize_2d_against_torch_antialiased_bilinear(shim, src, dst):
    import videoprism_oracle as O
    emb = torch.randn(src[0] * src[1], 5, dtype=torch.float64, generator=torch.Generator().manual_seed(src[0] + dst[1]))
    want = F.interpolate(emb.T.reshape(1, 5, *src), size=dst, mode="bilinear", align_corners=False, antialias=True)[0].reshape(5, -1).T
    got_shim = np.asarray(shim.image.resize(emb.numpy().reshape(1, src[0], src[1], 5), (1, dst[0], dst[1], 5), method="bilinear"))
    np.testing.assert_allclose(got_shim.reshape(-1, 5), want.numpy(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(O.interpolate_emb_2d(emb, src, dst).numpy(), want.numpy(), rtol=0, atol=1e-12)
