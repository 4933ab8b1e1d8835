"""fp32 CHECK MODE (vp_create_ex + VP_FLAG_CHECK_FP32; `get_model(name, check_fp32=True)`) against the golden vectors the
reference's own module code produced (tests/golden/make_golden.py).

Tolerance = the reference's own fp32 envelope between its two implementations (Flax vs MLX):
  * max-abs < 1e-3 on features and raw embeddings  (verify_clip_models.py:92-95; observed 2.24e-4 there,
    FLAX_TO_MLX_CONVERSION_GUIDE.md:321-336),
  * max-abs < 1e-5 on l2-normalised embeddings     (observed 5.9e-6 / 1.75e-7 there, :337-358).
The production path (bf16 tensor cores) is held to per-token cosine >= 0.999 in test_golden_gpu.py; this file is the
tighter bound BASELINE.json's north star asks for."""
import os

import numpy as np
import pytest

import videoprism_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FEATURE_TOL = 1e-3
NORMALISED_TOL = 1e-5


def report(tag, got, want, tol):
    err = float(np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64)).max())
    print(f"[check-fp32] {tag}: max-abs {err:.3g} (tolerance {tol:g}, reference max {np.abs(want).max():.4g})")
    assert err < tol, f"{tag}: max-abs {err} >= {tol}"


def model_for(cfg):
    import videoprism_b200 as vp
    cls = vp.FactorizedEncoder if cfg["kind"] == "encoder" else vp.FactorizedVideoCLIP
    m = cls(**{k: v for k, v in cfg.items() if k != "kind"})
    m.check_fp32 = True
    return m


def test_check_mode_is_on():
    import videoprism_b200 as vp
    import videoprism_b200._lib as L
    m = vp.get_model("videoprism_public_v1_base", check_fp32=True)
    assert L.lib().vp_handle_flags(m._ensure_handle()) & L.VP_FLAG_CHECK_FP32
    m2 = vp.get_model("videoprism_public_v1_base", fprop_dtype=np.float32)     # an explicit float32 request selects it too
    assert L.lib().vp_handle_flags(m2._ensure_handle()) & L.VP_FLAG_CHECK_FP32
    m3 = vp.get_model("videoprism_public_v1_base")
    if os.environ.get("VP_CHECK_FP32", "0") in ("", "0"):
        assert L.lib().vp_handle_flags(m3._ensure_handle()) == 0


def test_tiny_encoder_and_clip():
    """Every output of the tiny goldens: resized position tables, frame paddings, text paddings, all intermediates."""
    g = np.load(os.path.join(G, "enc_tiny_interp.npz"))
    cfg = O.tiny_config("encoder", pos_emb_shape=(16, 16, 16))
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(2, 4, 16, seed=11, kind="normal")
    m = model_for(cfg)
    out, outs = m.apply(W, v, train=False, return_intermediate=True)
    report("tiny encoder features", out, g["features"], FEATURE_TOL)
    report("tiny encoder spatial_features", outs["spatial_features"], g["spatial_features"], FEATURE_TOL)
    outp, _ = m.apply(W, v, train=False, frame_paddings=g["frame_paddings"])
    report("tiny encoder with frame_paddings", outp, g["features_frame_paddings"], FEATURE_TOL)

    g = np.load(os.path.join(G, "clip_tiny.npz"))
    cfg = O.tiny_config("clip")
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(3, 4, 16, seed=13, kind="normal")
    m = model_for(cfg)
    ve, te, outs = m.apply(W, v, g["ids"], g["paddings"], train=False, return_intermediate=True)
    report("tiny clip video_emb (normalised)", ve, g["video_emb_norm"], NORMALISED_TOL)
    report("tiny clip text_emb (normalised)", te, g["text_emb_norm"], NORMALISED_TOL)
    for k in ("spatial_features", "spatiotemporal_features"):
        report("tiny clip " + k, outs[k], g[k], FEATURE_TOL)
    report("tiny clip frame_embeddings (normalised)", outs["frame_embeddings"], g["frame_embeddings"], NORMALISED_TOL)
    ve, te, _ = m.apply(W, v, g["ids"], g["paddings"], train=False, normalize=False)
    report("tiny clip video_emb (raw)", ve, g["video_emb_raw"], FEATURE_TOL)
    report("tiny clip text_emb (raw)", te, g["text_emb_raw"], FEATURE_TOL)


def test_base_encoder_config1():
    """BASELINE configs[0]: videoprism_public_v1_base, 1 x 16 x 288 x 288 x 3, against the reference-generated golden."""
    import videoprism_b200 as vp
    g = np.load(os.path.join(G, "base_config1.npz"))
    name = "videoprism_public_v1_base"
    m = vp.get_model(name, check_fp32=True)
    out, _ = m.apply(O.make_synthetic_weights(O.CONFIGS[name]), O.make_video(1, 16, 288, seed=0, kind="uniform"), train=False)
    report("base encoder config 1 features", out[:, :: int(g["token_stride"])], g["features_sample"], FEATURE_TOL)


def test_lvt_base_one_clip_three_queries():
    """The reference's own benchmark shape (scripts/benchmark_performance.py:70-94): 1 clip + 3 queries through the
    video-text model, normalised and raw embeddings and the similarity matrix (verify_clip_models.py:92-95)."""
    import videoprism_b200 as vp
    g = np.load(os.path.join(G, "lvt_base_1clip_3text.npz"))
    name = "videoprism_lvt_public_v1_base"
    m = vp.get_model(name, check_fp32=True)
    W = O.make_synthetic_weights(O.CONFIGS[name])
    v = O.make_video(1, 16, 288, seed=0)
    ve, te, _ = m.apply(W, v, g["ids"], g["paddings"], train=False)
    report("lvt base video_emb (normalised)", ve, g["video_emb"], NORMALISED_TOL)
    report("lvt base text_emb (normalised)", te, g["text_emb"], NORMALISED_TOL)
    report("lvt base similarity matrix", ve @ te.T, g["video_emb"] @ g["text_emb"].T, NORMALISED_TOL)
    ve, te, _ = m.apply(W, v, g["ids"], g["paddings"], train=False, normalize=False)
    report("lvt base video_emb (raw)", ve, g["video_emb_raw"], FEATURE_TOL)
    report("lvt base text_emb (raw)", te, g["text_emb_raw"], FEATURE_TOL)


def test_production_path_agrees_with_check_mode():
    """The two GPU paths against each other on one clip: the bf16 path's error is what the cosine bound says it is."""
    import videoprism_b200 as vp
    name = "videoprism_public_v1_base"
    W = O.make_synthetic_weights(O.CONFIGS[name])
    v = O.make_video(1, 16, 288, seed=3)
    ref, _ = vp.get_model(name, check_fp32=True).apply(W, v, train=False)
    got, _ = vp.get_model(name).apply(W, v, train=False)
    a = ref.reshape(-1, ref.shape[-1]).astype(np.float64); b = got.reshape(-1, got.shape[-1]).astype(np.float64)
    cos = ((a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1))).min()
    print(f"[check-fp32] bf16 path vs fp32 check mode: min per-token cosine {cos:.6f}, max-abs {np.abs(a - b).max():.4g}")
    assert cos >= 0.999
