"""CPU checks of the oracle against the reference's structural known-answers."""
import numpy as np
import pytest
import torch

import videoprism_oracle as O


def test_leaf_counts_match_reference_tests():
    # encoders_test.py:112 (16 per scanned stack), :170 (40), layers_test.py:282 (12), encoders_test.py:273 (20), :340 (88)
    enc = O.param_specs(O.CONFIGS["videoprism_public_v1_base"])
    clip = O.param_specs(O.CONFIGS["videoprism_lvt_public_v1_base"])
    assert len(enc) == 40
    assert len(clip) == 88
    assert sum(1 for k, _, _ in enc if "/spatial_encoder/" in k) == 16
    assert sum(1 for k, _, _ in clip if "/contrastive_vision_pooler/" in k) == 12
    assert sum(1 for k, _, _ in clip if "/text_encoder/" in k) == 20


@pytest.mark.parametrize("name,millions", [
    ("videoprism_public_v1_base", 114), ("videoprism_public_v1_large", 354),
    ("videoprism_lvt_public_v1_base", 248), ("videoprism_lvt_public_v1_large", 580)])
def test_param_counts_match_readme(name, millions):
    # README.md:34-37, :159-162
    assert round(O.count_params(O.CONFIGS[name]) / 1e6) == millions


def test_resize_bilinear_upsample_known_answer():
    # jax.image.resize 8 -> 16: out[0]=in[0], out[1]=.75 in[0]+.25 in[1], ..., out[15]=in[7]
    w = O._resize_weights(8, 16, torch.float64).T
    np.testing.assert_allclose(w[0, :2], [1.0, 0.0])
    np.testing.assert_allclose(w[1, :2], [0.75, 0.25])
    np.testing.assert_allclose(w[2, :2], [0.25, 0.75])
    np.testing.assert_allclose(w[15, 6:], [0.0, 1.0])
    np.testing.assert_allclose(w.sum(1), np.ones(16))
    # down-sampling uses the widened (antialiased) triangle and still sums to one
    wd = O._resize_weights(16, 4, torch.float64).T
    np.testing.assert_allclose(wd.sum(1), np.ones(4))
    assert (wd[0] > 0).sum() == 6


def test_image_to_patch_order():
    # encoders.py:95-103: '(m p)(n q) c -> (m n)(p q c)'
    x = torch.arange(2 * 8 * 8 * 3, dtype=torch.float32).reshape(2, 8, 8, 3)
    p = O.image_to_patch(x, 4)
    assert p.shape == (2, 4, 48)
    assert p[1, 3, (2 * 4 + 1) * 3 + 2] == x[1, 4 + 2, 4 + 1, 2]


def test_masks_and_uniform_rows():
    # layers.py:92-152: merged causal + padding mask; a padded query row is fully masked -> uniform softmax
    pad = torch.tensor([[0.0, 0.0, 1.0, 1.0]])
    x = torch.zeros(1, 4, 8)
    m = O.attention_masks_for_fprop(x, pad, causal=True)
    assert m.shape == (1, 1, 4, 4)
    neg = O._neg(torch.float32)
    ok = (m >= neg * 0.5)[0, 0]
    assert ok.tolist() == [[True, False, False, False], [True, True, False, False], [False] * 4, [False] * 4]
    q = torch.randn(1, 4, 2, 4)
    _, probs = O.dot_attention(q, q, q, m, 50.0, None, 8)
    np.testing.assert_allclose(probs[0, :, 2].numpy(), 0.25, rtol=1e-6)


def test_shapes_follow_reference_tests():
    # encoders_test.py:115-181 (tiny FactorizedEncoder incl. pos-emb interpolation, frame paddings, spatial_features)
    cfg = O.tiny_config("encoder", pos_emb_shape=(16, 16, 16), model_dim=32, mlp_dim=16)
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(1, 4, 16, kind="normal")
    fp = np.zeros((1, 4), np.float32); fp[:, 2:] = 1
    out, outs = O.run_encoder(cfg, W, v, return_intermediate=True, frame_paddings=torch.from_numpy(fp))
    assert out.shape == (1, 4 * 16, 32) and outs["spatial_features"].shape == (1, 64, 32)
    assert np.isfinite(out).all()
    out2, outs2 = O.run_encoder(cfg, W, v)
    assert outs2 == {} and not np.allclose(out, out2)


def test_clip_shapes_and_fp64_agreement():
    # encoders_test.py:339-371; models_test.py:55-91
    cfg = O.tiny_config("clip")
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(2, 4, 16, kind="normal")
    ids, pad = O.make_text(3, vocab=128, max_len=8)
    ve, te, outs = O.run_clip(cfg, W, v, ids, pad, return_intermediate=True)
    assert ve.shape == (2, 64) and te.shape == (3, 64)
    assert set(outs) == {"spatial_features", "spatiotemporal_features", "frame_embeddings"}
    assert outs["frame_embeddings"].shape == (2, 4, 64)
    np.testing.assert_allclose(np.linalg.norm(ve, axis=-1), 1.0, rtol=1e-5)
    ve64, te64, _ = O.run_clip(cfg, W, v, ids, pad, dtype=torch.float64)
    assert np.abs(ve - ve64).max() < 1e-5 and np.abs(te - te64).max() < 1e-5
    v_only, t_none, _ = O.run_clip(cfg, W, v)
    assert t_none is None and np.allclose(v_only, ve)


def test_synthetic_inputs_are_deterministic():
    a = O.make_synthetic_weights(O.tiny_config("encoder"))
    b = O.make_synthetic_weights(O.tiny_config("encoder"))
    assert all(np.array_equal(a[k], b[k]) for k in a)
    ids, pad = O.make_text(5)
    assert ids.shape == (5, 64) and ((pad == 1) == (ids == 0)).all()
