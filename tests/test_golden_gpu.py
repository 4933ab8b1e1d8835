"""CUDA path (through the reference-facing surface / C ABI) vs the committed golden vectors produced by the
reference's own module code (tests/golden/make_golden.py).  bf16 tolerance of BASELINE.json's north star:
per-token cosine >= 0.999, max-abs reported."""
import os

import numpy as np
import pytest

import videoprism_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COS_MIN = 0.999


def cos_min(a, b):
    a = a.reshape(-1, a.shape[-1]).astype(np.float64); b = b.reshape(-1, b.shape[-1]).astype(np.float64)
    return float(((a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1) + 1e-30)).min())


def check(tag, got, want):
    c = cos_min(got, want)
    print(f"[golden] {tag}: min cosine {c:.6f}, max-abs {np.abs(got - want).max():.4g} (ref max {np.abs(want).max():.4g})")
    assert c >= COS_MIN


def model_for(cfg):
    import videoprism_b200 as vp
    cls = vp.FactorizedEncoder if cfg["kind"] == "encoder" else vp.FactorizedVideoCLIP
    return cls(**{k: v for k, v in cfg.items() if k != "kind"})


def test_tiny_encoder_cases():
    g = np.load(os.path.join(G, "enc_tiny_interp.npz"))
    cfg = O.tiny_config("encoder", pos_emb_shape=(16, 16, 16))
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(2, 4, 16, seed=11, kind="normal")
    m = model_for(cfg)
    out, outs = m.apply(W, v, train=False, return_intermediate=True)
    check("tiny encoder (tables down-sampled)", out, g["features"])
    check("  spatial_features", outs["spatial_features"], g["spatial_features"])
    outp, _ = m.apply(W, v, train=False, frame_paddings=g["frame_paddings"])
    check("  with frame_paddings", outp, g["features_frame_paddings"])
    g = np.load(os.path.join(G, "enc_tiny_upsample.npz"))
    cfg = O.tiny_config("encoder")
    out, _ = model_for(cfg).apply(O.make_synthetic_weights(cfg), O.make_video(2, 8, 32, seed=12, kind="normal"), train=False)
    check("tiny encoder (tables up-sampled)", out, g["features"])


def test_tiny_clip_case():
    g = np.load(os.path.join(G, "clip_tiny.npz"))
    cfg = O.tiny_config("clip")
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(3, 4, 16, seed=13, kind="normal")
    m = model_for(cfg)
    ve, te, outs = m.apply(W, v, g["ids"], g["paddings"], train=False, return_intermediate=True)
    check("tiny clip video_emb", ve, g["video_emb_norm"])
    check("tiny clip text_emb", te, g["text_emb_norm"])
    for k in ("spatial_features", "spatiotemporal_features", "frame_embeddings"):
        check("  " + k, outs[k], g[k])
    ve, te, _ = m.apply(W, v, g["ids"], g["paddings"], train=False, normalize=False)
    check("tiny clip video_emb (raw)", ve, g["video_emb_raw"])
    check("tiny clip text_emb (raw)", te, g["text_emb_raw"])


@pytest.mark.parametrize("case,T,seed,kind", [("base_config1", 16, 0, "uniform"), ("base_T8", 8, 1, "normal")])
def test_base_encoder_full_size(case, T, seed, kind):
    import videoprism_b200 as vp
    g = np.load(os.path.join(G, case + ".npz"))
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    m = vp.get_model("videoprism_public_v1_base")
    out, _ = m.apply(O.make_synthetic_weights(cfg), O.make_video(1, T, 288, seed=seed, kind=kind), train=False)
    check(case, out[:, :: int(g["token_stride"])], g["features_sample"])


def test_lvt_base_one_clip_three_queries():
    import videoprism_b200 as vp
    g = np.load(os.path.join(G, "lvt_base_1clip_3text.npz"))
    cfg = O.CONFIGS["videoprism_lvt_public_v1_base"]
    m = vp.get_model("videoprism_lvt_public_v1_base")
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(1, 16, 288, seed=0)
    ve, te, _ = m.apply(W, v, g["ids"], g["paddings"], train=False)
    check("lvt base video_emb", ve, g["video_emb"])
    check("lvt base text_emb", te, g["text_emb"])
    sim_err = np.abs(ve @ te.T - g["video_emb"] @ g["text_emb"].T).max()
    print(f"[golden] lvt base similarity-matrix max-abs diff {sim_err:.4g}")
    assert sim_err < 2e-2
    ve, te, _ = m.apply(W, v, g["ids"], g["paddings"], train=False, normalize=False)
    check("lvt base video_emb (raw)", ve, g["video_emb_raw"])
    check("lvt base text_emb (raw)", te, g["text_emb_raw"])


def test_tiny_classifier_case():
    """FactorizedVideoClassifier (encoders.py:583-653) against the reference-generated golden: logits within 2e-2 absolute
    (bf16 path; logits are O(0.1), so cosine over 10 classes is also asserted), intermediates by cosine."""
    import videoprism_b200 as vp
    g = np.load(os.path.join(G, "classifier_tiny.npz"))
    cfg = O.tiny_config("classifier")
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(3, 4, 16, seed=14, kind="normal")
    enc = {k: x for k, x in cfg.items() if k not in ("kind", "num_classes")}
    m = vp.FactorizedVideoClassifier(encoder_params=enc, num_classes=cfg["num_classes"])
    assert len(m.param_shapes()) == 54                     # encoders_test.py:224
    logits, outs = m.apply(W, v, train=False, return_intermediate=True)
    assert logits.shape == (3, 10)
    for k in ("spatial_features", "spatiotemporal_features", "global_embeddings"):
        check("classifier " + k, outs[k], g[k])
    print(f"[golden] classifier logits max-abs {np.abs(logits - g['logits']).max():.4g} (ref max {np.abs(g['logits']).max():.4g})")
    assert np.abs(logits - g["logits"]).max() <= 2e-2
    lp, _ = m.apply(W, v, train=False, frame_paddings=g["frame_paddings"])
    assert np.abs(lp - g["logits_frame_paddings"]).max() <= 2e-2
    # torch CUDA tensors in -> torch tensors out, same numbers
    import torch
    lt, _ = m(torch.from_numpy(v).cuda())
    assert np.array_equal(lt.cpu().numpy(), logits)


def test_tiny_encoder_giant_head_width():
    """dim_per_head = 88 (the giant configurations, models.py:105-115): generic attention kernels with zero-padded head
    tiles, GEMMs with K = 176 (not a multiple of the 64-wide K block) and N = 176 (partial 128-column tiles)."""
    g = np.load(os.path.join(G, "enc_tiny_dh88.npz"))
    cfg = O.tiny_config("encoder", model_dim=176, num_heads=2, mlp_dim=352)
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(2, 5, 16, seed=15, kind="normal")
    m = model_for(cfg)
    out, outs = m.apply(W, v, train=False, return_intermediate=True)
    check("tiny encoder dh=88", out, g["features"])
    check("  spatial_features", outs["spatial_features"], g["spatial_features"])
    outp, _ = m.apply(W, v, train=False, frame_paddings=g["frame_paddings"])
    check("  with frame_paddings", outp, g["features_frame_paddings"])


@pytest.mark.parametrize("fuse_ln", ["1", "0"])
def test_tiny_clip_primer_hybrid_text_tower(fuse_ln, monkeypatch):
    """Text tower with norm_policy 'primer_hybrid' (models.py:146-161 giant video-text configuration) against the golden
    made by the reference's own code; with the pre-LayerNorms folded into the GEMMs (default) and standalone."""
    import videoprism_b200 as vp
    monkeypatch.setenv("VP_FUSE_LN", fuse_ln)
    g = np.load(os.path.join(G, "clip_tiny_primer.npz"))
    cfg = O.tiny_config("clip", norm_policy="primer_hybrid")
    W = O.make_synthetic_weights(cfg)
    m = model_for(cfg)
    assert list(vp.synthetic_state(m, seed=1234)) == list(W)          # same leaves, same order as the oracle's tree
    v = O.make_video(2, 4, 16, seed=16, kind="normal")
    ve, te, _ = m.apply(W, v, g["ids"], g["paddings"], train=False)
    check("primer clip video_emb", ve, g["video_emb"])
    check("primer clip text_emb", te, g["text_emb"])
    _, te_raw, _ = m.apply(W, None, g["ids"], g["paddings"], train=False, normalize=False)
    check("primer clip text_emb (raw)", te_raw, g["text_emb_raw"])


def test_large_models_full_size():
    """The large encoder (BASELINE configs[2]) and the large video-text model (configs[4]) against reference-generated goldens."""
    import videoprism_b200 as vp
    g = np.load(os.path.join(G, "large_config3.npz"))
    name = "videoprism_public_v1_large"
    m = vp.get_model(name)
    out, _ = m.apply(O.make_synthetic_weights(O.CONFIGS[name]), O.make_video(1, 16, 288, seed=5), train=False)
    check("large encoder vs reference golden", out[:, ::int(g["token_stride"])], g["features_sample"])
    del m
    g = np.load(os.path.join(G, "lvt_large_1clip_3text.npz"))
    name = "videoprism_lvt_public_v1_large"
    m = vp.get_model(name)
    ve, te, _ = m.apply(O.make_synthetic_weights(O.CONFIGS[name]), O.make_video(1, 16, 288, seed=6), g["ids"], g["paddings"], train=False)
    check("lvt large video_emb vs reference golden", ve, g["video_emb"])
    check("lvt large text_emb vs reference golden", te, g["text_emb"])
    assert np.abs(ve @ te.T - g["video_emb"] @ g["text_emb"].T).max() < 1e-3      # verify_clip_models.py:92-95
