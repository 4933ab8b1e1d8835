"""GPU parity of the whole forward (through the reference-facing surface and the C ABI) against the
CPU oracle on the same seeded inputs and the same random-init weights.

Tolerance (BASELINE.json north_star): per-token cosine similarity >= 0.999 for the bf16 path, with
the max-abs error reported (printed) for each case."""
import numpy as np
import pytest
import torch

import videoprism_oracle as O

pytestmark = pytest.mark.gpu

COS_MIN = 0.999


def per_token_cosine(a, b):
    a = a.reshape(-1, a.shape[-1]).astype(np.float64)
    b = b.reshape(-1, b.shape[-1]).astype(np.float64)
    return (a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1) + 1e-30)


def report(tag, got, want):
    cos = per_token_cosine(got, want)
    err = np.abs(got - want).max()
    print(f"[parity] {tag}: min cosine {cos.min():.6f}, max-abs {err:.4g}, ref max {np.abs(want).max():.4g}")
    return cos.min(), err


def make_model(cfg):
    import videoprism_b200 as vp
    kw = {k: v for k, v in cfg.items() if k != "kind"}
    cls = vp.FactorizedEncoder if cfg["kind"] == "encoder" else vp.FactorizedVideoCLIP
    return cls(**kw)


def test_synthetic_state_matches_oracle_generator():
    cfg = O.tiny_config("clip")
    m = make_model(cfg)
    import videoprism_b200 as vp
    mine = vp.synthetic_state(m, seed=1234)
    ref = O.make_synthetic_weights(cfg, seed=1234)
    assert list(mine) == list(ref)
    assert all(np.array_equal(mine[k], ref[k]) for k in ref)


@pytest.mark.parametrize("T,size,pos", [(4, 16, (4, 4, 4)), (8, 32, (4, 4, 4)), (4, 16, (16, 16, 16))])
def test_tiny_encoder_parity(T, size, pos):
    cfg = O.tiny_config("encoder", pos_emb_shape=pos)
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(3, T, size, seed=1, kind="normal")
    want, wouts = O.run_encoder(cfg, W, v, return_intermediate=True)
    m = make_model(cfg)
    got, gouts = m.apply(W, v, train=False, return_intermediate=True)
    assert got.shape == want.shape and got.dtype == np.float32
    c, _ = report(f"tiny encoder T={T} size={size} pos={pos}", got, want)
    assert c >= COS_MIN
    c2, _ = report("  spatial_features", gouts["spatial_features"], wouts["spatial_features"])
    assert c2 >= COS_MIN
    # device-buffer path returns the same numbers as the host-buffer path
    got_dev, _ = m.apply(W, torch.from_numpy(v).cuda(), train=False)
    assert np.array_equal(got_dev.cpu().numpy(), got)


def test_tiny_encoder_frame_paddings():
    cfg = O.tiny_config("encoder")
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(2, 4, 16, seed=2, kind="normal")
    fp = np.zeros((2, 4), np.float32)
    fp[0, 2:] = 1
    want, _ = O.run_encoder(cfg, W, v, frame_paddings=torch.from_numpy(fp))
    got, _ = make_model(cfg).apply(W, v, frame_paddings=fp)
    c, _ = report("tiny encoder + frame_paddings", got, want)
    assert c >= COS_MIN


def test_tiny_clip_parity():
    cfg = O.tiny_config("clip")
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(3, 4, 16, seed=3, kind="normal")
    ids, pad = O.make_text(5, vocab=cfg["vocabulary_size"], max_len=8)
    pad[:, 4:] = (np.arange(4)[None, :] >= np.array([0, 1, 2, 3, 4])[:, None]).astype(np.float32)
    ids = np.where(pad > 0, 0, ids).astype(np.int32)
    for normalize in (True, False):
        wv, wt, wo = O.run_clip(cfg, W, v, ids, pad, normalize=normalize, return_intermediate=True)
        m = make_model(cfg)
        gv, gt, go = m.apply(W, v, ids, pad, train=False, normalize=normalize, return_intermediate=True)
        assert report(f"tiny clip video emb (normalize={normalize})", gv, wv)[0] >= COS_MIN
        assert report(f"tiny clip text emb (normalize={normalize})", gt, wt)[0] >= COS_MIN
        for k in ("spatial_features", "spatiotemporal_features", "frame_embeddings"):
            assert report(f"  {k}", go[k], wo[k])[0] >= COS_MIN
    gv2, gt2, _ = m.apply(W, v, None, None)
    assert gt2 is None and gv2.shape == (3, cfg["model_dim"])
    gv3, gt3, _ = m.apply(W, None, ids, pad)
    assert gv3 is None and gt3.shape == (5, cfg["model_dim"])


def test_base_encoder_parity_config1():
    """BASELINE.json configs[0]: videoprism_public_v1_base, 1x16x288x288x3 (uniform [0,1) clip)."""
    import videoprism_b200 as vp
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(1, 16, 288, seed=0)
    want, _ = O.run_encoder(cfg, W, v)
    m = vp.get_model("videoprism_public_v1_base")
    got, _ = m.apply(W, v, train=False)
    assert got.shape == (1, 4096, 768)
    c, e = report("base encoder config-1", got, want)
    assert c >= COS_MIN
    assert m.kernel_launches > 0


def _stress_weights(kind):
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    base = O.make_synthetic_weights(cfg)
    W = dict(base)
    if kind == "proj_x3":          # all projection matrices x3 (N(0, 0.06))
        W = {k: (v * 3 if v.ndim >= 2 and "emb_var" not in k else v) for k, v in base.items()}
    elif kind == "qk_x8":          # query / key projections x8: logits reach the 50 tanh cap, softmax rows are peaked
        for k in W:
            if k.endswith("self_attention/query/w") or k.endswith("self_attention/key/w"):
                W[k] = W[k] * 8
    elif kind == "massive":        # two massive residual channels (+60 / -45 on every token), as released ViTs have
        b = W["params/patch_projection/linear/bias"].copy()
        b[[7, 300]] = [60.0, -45.0]
        W["params/patch_projection/linear/bias"] = b
    elif kind == "ln":             # LayerNorm scale x3, bias +0.5: stresses the LayerNorm folded into the GEMMs
        for k in W:
            if k.endswith("layer_norm/bias"):
                W[k] = W[k] + 0.5
            if k.endswith("layer_norm/scale"):
                W[k] = W[k] * 3
    return cfg, W


@pytest.mark.parametrize("kind,cos_min", [("proj_x3", 0.999), ("massive", 0.999), ("ln", 0.999), ("qk_x8", 0.998)])
def test_base_encoder_parity_under_harsher_weight_statistics(kind, cos_min):
    """Parity beyond the near-uniform attention of the N(0, 0.02) random init (weights of released checkpoints have logits
    that reach the cap and a few massive residual channels).  Three of the four cases hold the 0.999 bound.

    SECOND, JUSTIFIED BOUND for `qk_x8` (|logit| ~ 30..50, single keys carry a row): 0.998.  There the loss is not in the
    softmax kernel (test_attention_peaked_logits holds it to its derived bound up to the saturated cap) but in the bf16
    rounding of q and k THEMSELVES, the operands of the tensor-core q.k^T: a relative operand error of 2^-9 on a logit of 30
    is 0.06..0.1 in the exponent, i.e. several per cent in individual softmax weights.  Evidence, all in this test:
      * the fp32 check mode of this library (same kernels' algorithm, fp32 operands) agrees with the oracle to < 1e-3 max-abs
        on the same weights, so the algorithm (cap, softmax, folded LayerNorm) is right under peaked logits;
      * the reference's own algorithm evaluated with every tensor in bfloat16 (its fprop_dtype=bfloat16 mode, run through
        the oracle on the CPU: profiles/r1_stress_parity.txt) gets min cosine 0.99480 on this case: the production path
        (bf16 operands, fp32 accumulation / softmax / residual adds) loses 4x less than that."""
    import videoprism_b200 as vp
    name = "videoprism_public_v1_base"
    cfg, W = _stress_weights(kind)
    v = O.make_video(1, 16, 288, seed=3)
    want, _ = O.run_encoder(cfg, W, v)
    got, _ = vp.get_model(name).apply(W, v, train=False)
    c, e = report(f"base encoder, harsher weights [{kind}] (bf16 path)", got, want)
    assert c >= cos_min
    if kind == "qk_x8":
        chk, _ = vp.get_model(name, check_fp32=True).apply(W, v, train=False)
        c32, e32 = report(f"base encoder, harsher weights [{kind}] (fp32 check mode)", chk, want)
        assert c32 >= 0.99999 and e32 < 1e-3
        assert c > 0.9948      # better than the all-bf16 evaluation of the reference algorithm on the same case


def test_large_encoder_parity_config3_shapes():
    """BASELINE.json configs[2]: videoprism_public_v1_large (24+4 blocks, D=1024, H=16, F=4096); its temporal
    table has 8 rows and is bilinearly resized to the 16 frames (encoders.py:551-552)."""
    import videoprism_b200 as vp
    name = "videoprism_public_v1_large"
    cfg = O.CONFIGS[name]
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(1, 16, 288, seed=5)
    want, _ = O.run_encoder(cfg, W, v)
    m = vp.get_model(name)
    got, _ = m.apply(W, v, train=False)
    assert got.shape == (1, 4096, 1024)
    c, _ = report("large encoder (config-3 model), 1 clip", got, want)
    assert c >= COS_MIN


def test_lvt_large_text_and_video_parity():
    """videoprism_lvt_public_v1_large: text tower on 6 ragged queries + video side on 1 clip."""
    import videoprism_b200 as vp
    name = "videoprism_lvt_public_v1_large"
    cfg = O.CONFIGS[name]
    W = O.make_synthetic_weights(cfg)
    ids, pad = O.make_text(6)
    v = O.make_video(1, 16, 288, seed=6)
    wv, wt, _ = O.run_clip(cfg, W, v, ids, pad)
    m = vp.get_model(name)
    gv, gt, _ = m.apply(W, v, ids, pad, train=False)
    assert report("lvt large video emb", gv, wv)[0] >= COS_MIN
    assert report("lvt large text emb", gt, wt)[0] >= COS_MIN


def test_batched_forward_matches_single_clip_forward():
    """Clips are independent (encoders.py:434-436): a B=5 forward equals five B=1 forwards bit for bit, and the
    chunk-pipelined host entry point equals the device entry point."""
    import videoprism_b200 as vp
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    m = vp.get_model("videoprism_public_v1_base")
    m.load_state(O.make_synthetic_weights(cfg))
    v = O.make_video(5, 16, 288, seed=8)
    full, _ = m(v)
    dev, _ = m(torch.from_numpy(v).cuda())
    assert np.array_equal(dev.cpu().numpy(), full)
    for b in (0, 4):
        one, _ = m(v[b:b + 1])
        assert np.array_equal(one[0], full[b])


@pytest.mark.parametrize("B,T,size", [(3, 16, 216), (2, 5, 288), (1, 1, 288), (7, 16, 144)])
def test_base_encoder_ragged_shapes(B, T, size):
    """Shapes off the benchmark grid on the full-size base model: other resolutions (position table resized,
    S = 144 / 64 tokens per frame -> the generic attention kernel), odd frame counts (temporal table resized,
    T = 1: single-key softmax), batch sizes that do not fill a GEMM tile."""
    import videoprism_b200 as vp
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(B, T, size, seed=20 + B)
    want, _ = O.run_encoder(cfg, W, v)
    m = vp.get_model("videoprism_public_v1_base")
    got, _ = m.apply(W, v, train=False)
    assert got.shape == want.shape
    assert report(f"base encoder B={B} T={T} {size}x{size}", got, want)[0] >= COS_MIN


def test_base_encoder_frame_paddings_full_size():
    """frame_paddings on the full-size model: padded frames are masked as keys in the temporal stack, padded
    frames attend uniformly in the spatial stack, and FFN outputs of padded tokens are zeroed (layers.py:397-411)."""
    import videoprism_b200 as vp
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(2, 16, 288, seed=31)
    fp = np.zeros((2, 16), np.float32)
    fp[0, 10:] = 1
    fp[1, ::4] = 1
    want, _ = O.run_encoder(cfg, W, v, frame_paddings=torch.from_numpy(fp))
    got, _ = vp.get_model("videoprism_public_v1_base").apply(W, v, train=False, frame_paddings=fp)
    assert report("base encoder + frame_paddings", got, want)[0] >= COS_MIN


def test_uint8_frames_match_float_frames_bitwise():
    """uint8 frames (what cv2 decodes) take the /255 on the device: identical to feeding
    `frames.astype(float32) / 255.0` (video_utils.py:88-93), on the host and the device entry points."""
    import videoprism_b200 as vp
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    m = vp.get_model("videoprism_public_v1_base")
    m.load_state(O.make_synthetic_weights(cfg))
    u8 = np.random.default_rng(3).integers(0, 256, (3, 16, 288, 288, 3), dtype=np.uint8)
    f32 = u8.astype(np.float32) / 255.0
    a, _ = m(f32)
    b, _ = m(u8)
    c, _ = m(torch.from_numpy(u8).cuda())
    assert np.array_equal(a, b) and np.array_equal(a, c.cpu().numpy())


def test_errors_mirror_reference():
    cfg = O.tiny_config("encoder")
    m = make_model(cfg)
    W = O.make_synthetic_weights(cfg)
    with pytest.raises(ValueError):       # encoders.py:86-90
        m.apply(W, np.zeros((1, 4, 18, 18, 3), np.float32))
    with pytest.raises(AssertionError):   # encoders.py:435
        m.apply(W, np.zeros((1, 4, 16, 32, 3), np.float32))
    bad = dict(W)
    bad.pop("params/temporal_ln/bias")
    with pytest.raises(KeyError):
        make_model(cfg).apply(bad, np.zeros((1, 4, 16, 16, 3), np.float32))


def test_trace_accounts_for_every_launch_and_leaves_results_unchanged():
    """vp_trace: one event per launch, labels of the path's stages, total launch count equal to vp_kernel_launches,
    and a traced forward is bitwise equal to an untraced one."""
    import videoprism_b200 as vp
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    m = vp.get_model("videoprism_public_v1_base")
    m.load_state(O.make_synthetic_weights(cfg))
    v = torch.from_numpy(O.make_video(2, 16, 288, seed=3)).cuda()
    plain, _ = m(v)
    n0 = m.kernel_launches
    m.trace(True)
    traced, _ = m(v)
    rows = m.trace_report()
    m.trace(False)
    per_forward = m.kernel_launches - n0
    labels = {r[0]: r for r in rows}
    assert labels["TOTAL"][1] == per_forward == 84
    for k, n in (("patchify", 1), ("patch_proj", 1), ("spatial.qkv", 12), ("spatial.attn", 12), ("spatial.ffn1", 12),
                 ("temporal.ffn2", 4), ("spatial_ln", 1), ("temporal_ln", 1)):
        assert labels[k][1] == n and labels[k][2] > 0.0
    assert torch.equal(plain, traced)


def test_base_classifier_parity():
    """videoprism_vc_v1_base (models.py:200-205; encoders.py:583-653) at full size, K400 head: one 16x288x288 clip against
    the fp32 oracle.  Logits are a 768-term dot product of an O(1) LayerNorm output with N(0, 0.02) weights (|logit| ~ 0.5):
    the bf16 path is held to 3e-2 absolute and cosine >= 0.999 over the class axis; global_embeddings to the token bar."""
    import videoprism_b200 as vp
    cfg = dict(O.CONFIGS["videoprism_public_v1_base"], kind="classifier", num_classes=400)
    W = O.make_synthetic_weights(cfg)
    v = O.make_video(1, 16, 288, seed=5)
    want, wouts = O.run_classifier(cfg, W, v, return_intermediate=("global_embeddings",))
    m = vp.models.videoprism_vc_v1_base(num_classes=400)
    got, gouts = m.apply(W, v, train=False, return_intermediate=("global_embeddings",))
    assert got.shape == (1, 400) and set(gouts) == {"global_embeddings"}
    c, err = report("base classifier logits", got, want)
    assert c >= COS_MIN and err <= 3e-2
    c2, _ = report("  global_embeddings", gouts["global_embeddings"], wouts["global_embeddings"])
    assert c2 >= COS_MIN


def test_load_classifier_takes_the_encoder_of_a_video_text_checkpoint(tmp_path):
    """models_mlx.load_classifier (models_mlx.py:213-294): `vision_encoder.*` of a video-text checkpoint becomes `encoder.*`,
    pooler and head are initialised fresh, so spatiotemporal_features equal the video-text model's bit for bit."""
    import videoprism_b200 as vp
    cfg = O.tiny_config("clip")
    W = O.make_synthetic_weights(cfg)
    path = str(tmp_path / "tiny_lvt.npz")
    np.savez(path, **W)
    v = O.make_video(2, 4, 16, seed=6, kind="normal")
    clip = make_model(cfg)
    _, _, couts = clip.apply(W, v, None, None, train=False, return_intermediate=("spatiotemporal_features",))
    models = dict(vp.MODELS)
    try:
        vp.MODELS["tiny_lvt"] = lambda: make_model(cfg)
        m = vp.load_classifier("tiny_lvt", num_classes=5, weights_path=path)
    finally:
        vp.MODELS.clear(); vp.MODELS.update(models)
    logits, outs = m(v, return_intermediate=True)
    assert logits.shape == (2, 5) and np.isfinite(logits).all()
    assert np.array_equal(outs["spatiotemporal_features"], couts["spatiotemporal_features"])


def test_batches_larger_than_one_pass_are_split_on_the_device_path():
    """The device entry point runs at most 2^18 tokens per pass (32-bit indices, bounded workspace) and loops over the
    rest: a 4100-clip tiny batch (4096 + 4) equals the two halves run separately, bit for bit, including
    spatial_features and frame_paddings."""
    cfg = O.tiny_config("encoder")
    W = O.make_synthetic_weights(cfg)
    m = make_model(cfg)
    B = 4100
    v = torch.from_numpy(O.make_video(B, 4, 16, seed=9, kind="normal")).cuda()
    fp = torch.zeros((B, 4), device="cuda")
    fp[4097, 2:] = 1
    fp[5, 1:] = 1
    out, outs = m.apply(W, v, train=False, return_intermediate=True, frame_paddings=fp)
    assert out.shape == (B, 4 * 16, cfg["model_dim"])
    a, ao = m(v[:4096], return_intermediate=True, frame_paddings=fp[:4096])
    b, bo = m(v[4096:], return_intermediate=True, frame_paddings=fp[4096:])
    assert torch.equal(out[:4096], a) and torch.equal(out[4096:], b)
    assert torch.equal(outs["spatial_features"][4096:], bo["spatial_features"])
    assert torch.isfinite(out).all()


def test_config2_full_batch_properties():
    """BASELINE.json configs[1] at its full size (base encoder, 32 clips of 16x288x288x3), through properties that do not
    need a 32-clip CPU oracle run: (i) clip 0 of the batch equals the reference-generated golden of configs[0] (same clip,
    same weights) to the token bar; (ii) permuting the clips permutes the outputs bit for bit (clips are independent,
    encoders.py:434-436); (iii) the host (numpy, chunk-pipelined) and device entry points agree bit for bit;
    (iv) every output is finite and LayerNorm-scaled (temporal_ln: per-token variance of O(1))."""
    import os
    import videoprism_b200 as vp
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    m = vp.get_model("videoprism_public_v1_base")
    m.load_state(O.make_synthetic_weights(cfg))
    B = 32
    v = np.concatenate([O.make_video(1, 16, 288, seed=0), O.make_video(B - 1, 16, 288, seed=21)], axis=0)
    vd = torch.from_numpy(v).cuda()
    out, _ = m(vd)
    assert out.shape == (B, 4096, 768) and torch.isfinite(out).all()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "base_config1.npz"))
    stride = int(g["token_stride"])
    c, _ = report("config-2 batch, clip 0 vs reference golden", out[0:1, ::stride].cpu().numpy(), g["features_sample"])
    assert c >= COS_MIN
    perm = torch.from_numpy(np.random.default_rng(3).permutation(B)).cuda()
    out_p, _ = m(vd[perm].contiguous())
    assert torch.equal(out_p, out[perm])
    host, _ = m(v)
    assert np.array_equal(host, out.cpu().numpy())
    var = out.float().var(dim=-1)
    assert 0.2 < float(var.min()) and float(var.max()) < 5.0


def test_text_tower_edge_cases():
    """Ragged text batches (the reference pads to 64 and marks paddings, models.py:385-407): a query with a single real
    token, a full-length query without padding, one query alone (Q = 1) and a batch whose rows are processed identically
    whatever their neighbours are (row independence, bit for bit)."""
    cfg = O.tiny_config("clip")
    W = O.make_synthetic_weights(cfg)
    m = make_model(cfg)
    L = 8
    rng = np.random.default_rng(4)
    ids = rng.integers(1, cfg["vocabulary_size"], (5, L)).astype(np.int32)
    lens = np.array([1, L, 3, L - 1, 2])
    pad = (np.arange(L)[None, :] >= lens[:, None]).astype(np.float32)
    ids = np.where(pad > 0, 0, ids).astype(np.int32)
    _, want, _ = O.run_clip(cfg, W, None, ids, pad)
    _, got, _ = m.apply(W, None, ids, pad, train=False)
    assert got.shape == (5, cfg["model_dim"])
    assert report("text tower, ragged lengths 1..L", got, want)[0] >= COS_MIN
    np.testing.assert_allclose(np.linalg.norm(got, axis=-1), 1.0, atol=1e-5)
    for q in (0, 1, 4):
        _, one, _ = m(None, ids[q:q + 1], pad[q:q + 1])
        assert np.array_equal(one[0], got[q])
    _, raw, _ = m(None, ids, pad, normalize=False)
    _, wraw, _ = O.run_clip(cfg, W, None, ids, pad, normalize=False)
    assert report("text tower, unnormalised", raw, wraw)[0] >= COS_MIN
    v_none, t_none, outs = m(None, None, None)
    assert v_none is None and t_none is None and outs == {}
    with pytest.raises(AssertionError):
        m(None, ids, None)            # encoders.py:888


def test_fprop_dtype_bfloat16_returns_bf16_features():
    """models.get_model(name, fprop_dtype=jnp.bfloat16) (models.py:283-301): bfloat16 features on the device path, equal
    to the float32 features rounded once (the same LayerNorm result, stored narrower)."""
    import videoprism_b200 as vp
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    W = O.make_synthetic_weights(cfg)
    v = torch.from_numpy(O.make_video(2, 16, 288, seed=12)).cuda()
    m32 = vp.get_model("videoprism_public_v1_base")
    m16 = vp.get_model("videoprism_public_v1_base", fprop_dtype=torch.bfloat16)
    f32, _ = m32.apply(W, v, train=False)
    b16, _ = m16.apply(W, v, train=False)
    assert f32.dtype == torch.float32 and b16.dtype == torch.bfloat16 and b16.shape == f32.shape
    assert torch.equal(b16, f32.to(torch.bfloat16))
    host, _ = m16.apply(W, v.cpu().numpy(), train=False)      # numpy callers keep float32
    assert host.dtype == np.float32 and np.array_equal(host, f32.cpu().numpy())


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_two_devices():
    """A handle is bound to the device that is current when it is created (include/videoprism_b200.h); one process may
    hold handles on several GPUs.  Kernel attributes (dynamic shared memory opt-in) are per device: the second device must
    run the >48 KB kernels too, and give bitwise the same features."""
    import videoprism_b200 as vp
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    W = O.make_synthetic_weights(cfg)
    v = torch.from_numpy(O.make_video(2, 16, 288, seed=31))
    outs = []
    for dev in (0, 1):
        with torch.cuda.device(dev):
            m = vp.get_model("videoprism_public_v1_base")
            m.load_state(W)
            o, _ = m(v.cuda(dev))
            torch.cuda.synchronize(dev)
            outs.append(o.cpu())
    assert torch.equal(outs[0], outs[1])
    assert m.device_index == 1
    with pytest.raises(ValueError, match="lives on cuda:1"):
        m(v.cuda(0))
    # video-text model on the second device: long-sequence attention, pooler, text tower
    cfgc = O.tiny_config("clip")
    Wc = O.make_synthetic_weights(cfgc)
    ids, pad = O.make_text(3, vocab=cfgc["vocabulary_size"], max_len=8)
    vt = O.make_video(2, 4, 16, seed=32, kind="normal")
    res = []
    for dev in (0, 1):
        with torch.cuda.device(dev):
            m = make_model(cfgc)
            res.append(m.apply(Wc, vt, ids, pad, train=False)[:2])
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


def test_benchmark_performance_script_runs_with_synthetic_fallbacks():
    """scripts/benchmark_performance.py (the reference's benchmark command line): on a box without the mp4, the checkpoint
    and the SentencePiece model it falls back to synthetic frames / weights / token ids and still reports timings."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "benchmark_performance.py"), "--runs", "2", "--warmup", "1",
                        "--video-path", "/nonexistent.mp4", "--text-tokenizer", "/nonexistent.model", "--normalize"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, HF_HUB_OFFLINE="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "mean=" in r.stdout and "video embeddings (1, 768), text embeddings (3, 768)" in r.stdout
    assert "synthetic uniform frames" in r.stdout and "synthetic token ids" in r.stdout


def test_release_workspace_frees_memory_and_the_next_forward_is_unchanged():
    import videoprism_b200 as vp
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    m = vp.get_model("videoprism_public_v1_base")
    m.load_state(O.make_synthetic_weights(cfg))
    v = torch.from_numpy(O.make_video(4, 16, 288, seed=41)).cuda()
    a, _ = m(v)
    torch.cuda.synchronize()
    free_before, _ = torch.cuda.mem_get_info()
    m.release_workspace()
    free_after, _ = torch.cuda.mem_get_info()
    assert free_after - free_before > 150 * 2**20      # 4 clips: x, n, qkv, u, patches ~ 250 MB
    b, _ = m(v[:2])
    assert torch.equal(a[:2], b)
