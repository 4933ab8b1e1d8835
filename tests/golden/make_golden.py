"""Generates the golden vectors that pin the oracle (and the CUDA path) to the REFERENCE'S OWN CODE.

Runs /root/reference/videoprism/{models,encoders,layers}.py UNMODIFIED (imported from where they lie) over the
numpy stand-ins of jax / flax / einshape in oracle/refshim, through the reference's public entry point
`models.get_model(name).apply(state, ...)` (or `encoders.FactorizedEncoder(**cfg)` for the tiny unit-test
configs of encoders_test.py).  Inputs and weights are the seeded synthetic ones of oracle/videoprism_oracle.py.

    python tests/golden/make_golden.py          # needs /root/reference; writes tests/golden/*.npz

Big outputs are stored as a strided token sample plus float64 checksums so the fixtures stay small.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "refshim"), "/root/reference", os.path.join(ROOT, "oracle")]

import numpy as np

import jax.numpy as jnp
from videoprism import encoders, models
import videoprism_oracle as O

TOKEN_STRIDE = 61   # coprime with 256 and 16: samples every frame and patch column


def tree_of(flat):
    tree = {}
    for k, v in flat.items():
        node = tree
        parts = k.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = v
    return tree


def summarise(x):
    x = np.asarray(x, dtype=np.float64)
    return np.array([x.sum(), np.abs(x).sum(), (x * x).sum()], dtype=np.float64)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def tiny_cases():
    # encoders_test.py:115-181: tiny FactorizedEncoder, pos_emb (16,16,16) vs a 4x4 grid / 4 frames (2-D and 1-D
    # antialiased down-sampling of the position tables), frame paddings, return_intermediate
    cfg = O.tiny_config("encoder", pos_emb_shape=(16, 16, 16))
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(2, 4, 16, seed=11, kind="normal")
    fp = np.zeros((2, 4), np.float32); fp[0, 2:] = 1
    kw = {k: x for k, x in cfg.items() if k != "kind"}
    m = encoders.FactorizedEncoder(scan=True, **kw)
    out, outs = m.apply(tree_of(W), jnp.asarray(v), train=False, return_intermediate=True)
    outp, _ = m.apply(tree_of(W), jnp.asarray(v), train=False, frame_paddings=jnp.asarray(fp))
    save("enc_tiny_interp", features=np.asarray(out), spatial_features=np.asarray(outs["spatial_features"]),
         features_frame_paddings=np.asarray(outp), frame_paddings=fp)
    # up-sampling of both tables: pos_emb (4,4,4), 8 frames of 32x32
    cfg = O.tiny_config("encoder")
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(2, 8, 32, seed=12, kind="normal")
    m = encoders.FactorizedEncoder(scan=True, **{k: x for k, x in cfg.items() if k != "kind"})
    out, _ = m.apply(tree_of(W), jnp.asarray(v), train=False)
    save("enc_tiny_upsample", features=np.asarray(out))
    # encoders_test.py:300-371: tiny FactorizedVideoCLIP with text paddings and all intermediates
    cfg = O.tiny_config("clip")
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(3, 4, 16, seed=13, kind="normal")
    ids, pad = O.make_text(5, vocab=cfg["vocabulary_size"], max_len=8)
    pad[:, 4:] = (np.arange(4)[None, :] >= np.array([0, 1, 2, 3, 4])[:, None]).astype(np.float32)
    ids = np.where(pad > 0, 0, ids).astype(np.int32)
    m = encoders.FactorizedVideoCLIP(scan=True, enable_causal_atten=True, **{k: x for k, x in cfg.items() if k != "kind"})
    arrays = {"ids": ids, "paddings": pad}
    for normalize in (True, False):
        ve, te, outs = m.apply(tree_of(W), jnp.asarray(v), jnp.asarray(ids), jnp.asarray(pad), train=False, normalize=normalize,
                               return_intermediate=True)
        tag = "norm" if normalize else "raw"
        arrays[f"video_emb_{tag}"] = np.asarray(ve)
        arrays[f"text_emb_{tag}"] = np.asarray(te)
        if normalize:
            for k, a in outs.items():
                arrays[k] = np.asarray(a)
    save("clip_tiny", **arrays)
    classifier_case()
    giant_head_case()
    primer_case()


def giant_head_case():
    # the giant configurations (models.py:105-115) have dim_per_head = 1408 / 16 = 88: tiny encoder with the same head
    # width (model_dim 176, 2 heads), 16 tokens per frame, 5 frames, a padded frame
    cfg = O.tiny_config("encoder", model_dim=176, num_heads=2, mlp_dim=352)
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(2, 5, 16, seed=15, kind="normal")
    fp = np.zeros((2, 5), np.float32); fp[1, 3:] = 1
    m = encoders.FactorizedEncoder(scan=True, **{k: x for k, x in cfg.items() if k != "kind"})
    out, outs = m.apply(tree_of(W), jnp.asarray(v), train=False, return_intermediate=True)
    outp, _ = m.apply(tree_of(W), jnp.asarray(v), train=False, frame_paddings=jnp.asarray(fp))
    save("enc_tiny_dh88", features=np.asarray(out), spatial_features=np.asarray(outs["spatial_features"]),
         features_frame_paddings=np.asarray(outp), frame_paddings=fp)


def primer_case():
    # the giant video-text configuration (models.py:146-161): text tower with norm_policy 'primer_hybrid'
    # (pre_layer_norm + post_layer_norm around attention and FFN); vision / auxiliary stacks stay 'pre'
    cfg = O.tiny_config("clip", norm_policy="primer_hybrid")
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(2, 4, 16, seed=16, kind="normal")
    ids, pad = O.make_text(5, vocab=cfg["vocabulary_size"], max_len=8)
    pad[:, 4:] = (np.arange(4)[None, :] >= np.array([0, 1, 2, 3, 4])[:, None]).astype(np.float32)
    ids = np.where(pad > 0, 0, ids).astype(np.int32)
    m = encoders.FactorizedVideoCLIP(scan=True, enable_causal_atten=True, **{k: x for k, x in cfg.items() if k != "kind"})
    ve, te, _ = m.apply(tree_of(W), jnp.asarray(v), jnp.asarray(ids), jnp.asarray(pad), train=False)
    ve_raw, te_raw, _ = m.apply(tree_of(W), jnp.asarray(v), jnp.asarray(ids), jnp.asarray(pad), train=False, normalize=False)
    save("clip_tiny_primer", ids=ids, paddings=pad, video_emb=np.asarray(ve), text_emb=np.asarray(te),
         video_emb_raw=np.asarray(ve_raw), text_emb_raw=np.asarray(te_raw))


def classifier_case():
    # encoders_test.py:183-231: tiny FactorizedVideoClassifier (encoder + atten_pooler + projection), all intermediates,
    # with and without frame paddings
    cfg = O.tiny_config("classifier")
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(3, 4, 16, seed=14, kind="normal")
    fp = np.zeros((3, 4), np.float32); fp[1, 3:] = 1; fp[2, 1:] = 1
    enc = {k: x for k, x in cfg.items() if k not in ("kind", "num_classes")}
    m = encoders.FactorizedVideoClassifier(encoder_params=dict(scan=True, **enc), num_classes=cfg["num_classes"])
    logits, outs = m.apply(tree_of(W), jnp.asarray(v), train=False, return_intermediate=True)
    logits_p, _ = m.apply(tree_of(W), jnp.asarray(v), train=False, frame_paddings=jnp.asarray(fp))
    save("classifier_tiny", logits=np.asarray(logits), logits_frame_paddings=np.asarray(logits_p), frame_paddings=fp,
         **{k: np.asarray(a) for k, a in outs.items()})


def full_size_cases():
    # BASELINE.json configs[0]: videoprism_public_v1_base, 1x16x288x288x3 uniform clip, through models.get_model
    name = "videoprism_public_v1_base"
    cfg = O.CONFIGS[name]
    W = tree_of(O.make_synthetic_weights(cfg))
    model = models.get_model(name)
    v = O.make_video(1, 16, 288, seed=0)
    out, _ = model.apply(W, jnp.asarray(v), train=False)
    out = np.asarray(out)
    save("base_config1", features_sample=out[:, ::TOKEN_STRIDE], checksum=summarise(out), token_stride=np.array(TOKEN_STRIDE))
    # models_test.py:36-53: T=8 (temporal table 16 -> 8, antialiased), N(0, 0.1) inputs
    v = O.make_video(1, 8, 288, seed=1, kind="normal")
    out, _ = model.apply(W, jnp.asarray(v), train=False)
    out = np.asarray(out)
    save("base_T8", features_sample=out[:, ::TOKEN_STRIDE], checksum=summarise(out), token_stride=np.array(TOKEN_STRIDE))
    # the reference's own benchmark shape (scripts/benchmark_performance.py): LvT base, 1 clip + 3 text queries
    name = "videoprism_lvt_public_v1_base"
    cfg = O.CONFIGS[name]
    W = tree_of(O.make_synthetic_weights(cfg))
    model = models.get_model(name)
    v = O.make_video(1, 16, 288, seed=0)
    ids, pad = O.make_text(3)
    ve, te, _ = model.apply(W, jnp.asarray(v), jnp.asarray(ids), jnp.asarray(pad), train=False)
    ve_raw, te_raw, _ = model.apply(W, jnp.asarray(v), jnp.asarray(ids), jnp.asarray(pad), train=False, normalize=False)
    save("lvt_base_1clip_3text", video_emb=np.asarray(ve), text_emb=np.asarray(te), video_emb_raw=np.asarray(ve_raw),
         text_emb_raw=np.asarray(te_raw), ids=ids, paddings=pad)


def large_cases():
    # BASELINE.json configs[2] / [4]: the large models (24 + 4 blocks, D = 1024, 16 heads, temporal table 8 -> 16 frames,
    # bilinear up-sampling) through models.get_model: encoder on 1 clip, video-text model on 1 clip + 3 queries
    name = "videoprism_public_v1_large"
    W = tree_of(O.make_synthetic_weights(O.CONFIGS[name]))
    v = O.make_video(1, 16, 288, seed=5)
    out, _ = models.get_model(name).apply(W, jnp.asarray(v), train=False)
    out = np.asarray(out)
    save("large_config3", features_sample=out[:, ::TOKEN_STRIDE], checksum=summarise(out), token_stride=np.array(TOKEN_STRIDE))
    del W
    name = "videoprism_lvt_public_v1_large"
    W = tree_of(O.make_synthetic_weights(O.CONFIGS[name]))
    v = O.make_video(1, 16, 288, seed=6)
    ids, pad = O.make_text(3)
    ve, te, _ = models.get_model(name).apply(W, jnp.asarray(v), jnp.asarray(ids), jnp.asarray(pad), train=False)
    save("lvt_large_1clip_3text", video_emb=np.asarray(ve), text_emb=np.asarray(te), ids=ids, paddings=pad)


if __name__ == "__main__":
    if "--large-only" in sys.argv:
        large_cases()
        sys.exit(0)
    if "--classifier-only" in sys.argv:
        classifier_case()
        sys.exit(0)
    if "--primer-only" in sys.argv:
        primer_case()
        sys.exit(0)
    if "--dh88-only" in sys.argv:
        giant_head_case()
        sys.exit(0)
    tiny_cases()
    if "--tiny-only" not in sys.argv:
        full_size_cases()
        large_cases()
