"""Oracle vs the committed golden vectors (tests/golden/*.npz), which were produced by the REFERENCE'S OWN
module code (videoprism/{models,encoders,layers}.py imported unmodified over the numpy jax/flax stand-ins
of oracle/refshim; generator: tests/golden/make_golden.py).  This is what pins the oracle.

Tolerance: the reference's own fp32 envelope (Flax vs MLX max-abs 2.24e-4 on features,
FLAX_TO_MLX_CONVERSION_GUIDE.md:321-342): max-abs <= 2e-4 on features, <= 1e-5 on normalised embeddings."""
import os

import numpy as np
import pytest
import torch

import videoprism_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FEAT_TOL, EMB_TOL = 2e-4, 1e-5


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def test_tiny_encoder_interpolated_tables_and_frame_paddings():
    g = load("enc_tiny_interp")
    cfg = O.tiny_config("encoder", pos_emb_shape=(16, 16, 16))
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(2, 4, 16, seed=11, kind="normal")
    out, outs = O.run_encoder(cfg, W, v, return_intermediate=True)
    assert np.abs(out - g["features"]).max() <= FEAT_TOL
    assert np.abs(outs["spatial_features"] - g["spatial_features"]).max() <= FEAT_TOL
    outp, _ = O.run_encoder(cfg, W, v, frame_paddings=torch.from_numpy(g["frame_paddings"]))
    assert np.abs(outp - g["features_frame_paddings"]).max() <= FEAT_TOL
    assert np.abs(outp - out).max() > 1e-2   # the paddings really change the result


def test_tiny_encoder_upsampled_tables():
    g = load("enc_tiny_upsample")
    cfg = O.tiny_config("encoder")
    out, _ = O.run_encoder(cfg, O.make_synthetic_weights(cfg), O.make_video(2, 8, 32, seed=12, kind="normal"))
    assert np.abs(out - g["features"]).max() <= FEAT_TOL


def test_tiny_clip():
    g = load("clip_tiny")
    cfg = O.tiny_config("clip")
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(3, 4, 16, seed=13, kind="normal")
    ve, te, outs = O.run_clip(cfg, W, v, g["ids"], g["paddings"], return_intermediate=True)
    assert np.abs(ve - g["video_emb_norm"]).max() <= EMB_TOL and np.abs(te - g["text_emb_norm"]).max() <= EMB_TOL
    for k in ("spatial_features", "spatiotemporal_features", "frame_embeddings"):
        assert np.abs(outs[k] - g[k]).max() <= FEAT_TOL
    ve, te, _ = O.run_clip(cfg, W, v, g["ids"], g["paddings"], normalize=False)
    assert np.abs(ve - g["video_emb_raw"]).max() <= FEAT_TOL and np.abs(te - g["text_emb_raw"]).max() <= FEAT_TOL


def test_tiny_encoder_giant_head_width():
    """dim_per_head = 88 as in the giant configurations (models.py:105-115), 5 frames (temporal table 4 -> 5), paddings."""
    g = load("enc_tiny_dh88")
    cfg = O.tiny_config("encoder", model_dim=176, num_heads=2, mlp_dim=352)
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(2, 5, 16, seed=15, kind="normal")
    out, outs = O.run_encoder(cfg, W, v, return_intermediate=True)
    assert np.abs(out - g["features"]).max() <= FEAT_TOL
    assert np.abs(outs["spatial_features"] - g["spatial_features"]).max() <= FEAT_TOL
    outp, _ = O.run_encoder(cfg, W, v, frame_paddings=torch.from_numpy(g["frame_paddings"]))
    assert np.abs(outp - g["features_frame_paddings"]).max() <= FEAT_TOL


def test_tiny_clip_primer_hybrid_text_tower():
    """norm_policy 'primer_hybrid' (the giant video-text configuration, models.py:146-161): only the text tower changes
    (encoders.py:899); 8 more leaves than the 'pre' model (88, encoders_test.py:340)."""
    g = load("clip_tiny_primer")
    cfg = O.tiny_config("clip", norm_policy="primer_hybrid")
    W = O.make_synthetic_weights(cfg)
    assert len(W) == 88 + 4
    v = O.make_video(2, 4, 16, seed=16, kind="normal")
    ve, te, _ = O.run_clip(cfg, W, v, g["ids"], g["paddings"])
    assert np.abs(ve - g["video_emb"]).max() <= EMB_TOL and np.abs(te - g["text_emb"]).max() <= EMB_TOL
    ve, te, _ = O.run_clip(cfg, W, v, g["ids"], g["paddings"], normalize=False)
    assert np.abs(ve - g["video_emb_raw"]).max() <= FEAT_TOL and np.abs(te - g["text_emb_raw"]).max() <= FEAT_TOL
    # and it is not the 'pre' model under another name
    cfg0 = O.tiny_config("clip")
    _, te0, _ = O.run_clip(cfg0, O.make_synthetic_weights(cfg0), None, g["ids"], g["paddings"])
    assert np.abs(te0 - g["text_emb"]).max() > 1e-2


def test_tiny_classifier():
    """encoders_test.py:183-231 shapes: FactorizedVideoClassifier through the reference's own code."""
    g = load("classifier_tiny")
    cfg = O.tiny_config("classifier")
    W = O.make_synthetic_weights(cfg)
    assert len(W) == 54                                   # encoders_test.py:224
    v = O.make_video(3, 4, 16, seed=14, kind="normal")
    logits, outs = O.run_classifier(cfg, W, v, return_intermediate=True)
    assert logits.shape == (3, 10)
    assert np.abs(logits - g["logits"]).max() <= FEAT_TOL
    for k in ("spatial_features", "spatiotemporal_features", "global_embeddings"):
        assert np.abs(outs[k] - g[k]).max() <= FEAT_TOL
    lp, _ = O.run_classifier(cfg, W, v, frame_paddings=torch.from_numpy(g["frame_paddings"]))
    assert np.abs(lp - g["logits_frame_paddings"]).max() <= FEAT_TOL
    assert np.abs(lp - logits).max() > 1e-3


@pytest.mark.parametrize("case,T,seed,kind", [("base_config1", 16, 0, "uniform"), ("base_T8", 8, 1, "normal")])
def test_base_encoder_full_size(case, T, seed, kind):
    g = load(case)
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    out, _ = O.run_encoder(cfg, O.make_synthetic_weights(cfg), O.make_video(1, T, 288, seed=seed, kind=kind))
    stride = int(g["token_stride"])
    assert np.abs(out[:, ::stride] - g["features_sample"]).max() <= FEAT_TOL
    x = out.astype(np.float64)
    got = np.array([x.sum(), np.abs(x).sum(), (x * x).sum()])
    np.testing.assert_allclose(got[1:], g["checksum"][1:], rtol=1e-5)


def test_lvt_base_one_clip_three_queries():
    g = load("lvt_base_1clip_3text")
    cfg = O.CONFIGS["videoprism_lvt_public_v1_base"]
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(1, 16, 288, seed=0)
    ve, te, _ = O.run_clip(cfg, W, v, g["ids"], g["paddings"])
    assert np.abs(ve - g["video_emb"]).max() <= EMB_TOL and np.abs(te - g["text_emb"]).max() <= EMB_TOL
    # the verify script's own criterion (verify_clip_models.py:92-95): cosine-similarity matrix within 1e-3
    assert np.abs(ve @ te.T - g["video_emb"] @ g["text_emb"].T).max() < 1e-3


def test_large_models_full_size():
    """BASELINE.json configs[2] / [4] models (large: 24 + 4 blocks, D = 1024, temporal table 8 -> 16) against goldens made by
    the reference's own code through models.get_model."""
    g = load("large_config3")
    cfg = O.CONFIGS["videoprism_public_v1_large"]
    out, _ = O.run_encoder(cfg, O.make_synthetic_weights(cfg), O.make_video(1, 16, 288, seed=5))
    stride = int(g["token_stride"])
    assert np.abs(out[:, ::stride] - g["features_sample"]).max() <= FEAT_TOL
    x = out.astype(np.float64)
    np.testing.assert_allclose(np.array([np.abs(x).sum(), (x * x).sum()]), g["checksum"][1:], rtol=1e-5)
    g = load("lvt_large_1clip_3text")
    cfg = O.CONFIGS["videoprism_lvt_public_v1_large"]
    ve, te, _ = O.run_clip(cfg, O.make_synthetic_weights(cfg), O.make_video(1, 16, 288, seed=6), g["ids"], g["paddings"])
    assert np.abs(ve - g["video_emb"]).max() <= EMB_TOL and np.abs(te - g["text_emb"]).max() <= EMB_TOL
