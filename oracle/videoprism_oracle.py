"""CPU oracle for the VideoPrism FactorizedEncoder / FactorizedVideoCLIP forward.

TEST INFRASTRUCTURE ONLY.  This module is a CPU restatement (PyTorch-CPU, fp32 by
default, fp64 on request) of the reference Flax implementation
(`/root/reference/videoprism/layers.py`, `encoders.py`, `models.py`).  It is the
checker for the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it; the product
package (`videoprism-mlx_b200/`) never does.

Pinning status: the reference's own tests hold no numeric goldens for this path
 and jax/flax are not installable here, so XLA itself cannot be
run.  The oracle is instead pinned against the reference's *own module code*
(`layers.py`/`encoders.py`, imported unmodified) executed over a numpy stand-in
for the jax/flax primitives it calls (`oracle/refshim/`, generator
`tests/golden/make_golden.py`); the vectors are committed under `tests/golden/`.
See DESIGN.md §Oracle for what that does and does not prove.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Collection, Dict, List, Optional, Tuple

import numpy as np
import torch

# ----------------------------------------------------------------------------
# Configs (reference: videoprism/models.py:82-161, MODELS :224-233,
#          TEXT_TOKENIZERS vocab_size 32000 :56-61, TEXT_MAX_LEN :54)
# ----------------------------------------------------------------------------
TEXT_MAX_LEN = 64

CONFIGS = {
    "videoprism_public_v1_base": dict(
        kind="encoder", patch_size=18, pos_emb_shape=(16, 16, 16), model_dim=768,
        num_spatial_layers=12, num_temporal_layers=4, num_heads=12, mlp_dim=3072,
        atten_logit_cap=50.0),
    "videoprism_public_v1_large": dict(
        kind="encoder", patch_size=18, pos_emb_shape=(8, 16, 16), model_dim=1024,
        num_spatial_layers=24, num_temporal_layers=4, num_heads=16, mlp_dim=4096,
        atten_logit_cap=50.0),
    "videoprism_lvt_public_v1_base": dict(
        kind="clip", patch_size=18, pos_emb_shape=(16, 16, 16), model_dim=768,
        num_spatial_layers=12, num_temporal_layers=4, num_heads=12, mlp_dim=3072,
        num_auxiliary_layers=2, num_unimodal_layers=12, vocabulary_size=32000,
        atten_logit_cap=50.0),
    "videoprism_lvt_public_v1_large": dict(
        kind="clip", patch_size=18, pos_emb_shape=(8, 16, 16), model_dim=1024,
        num_spatial_layers=24, num_temporal_layers=4, num_heads=16, mlp_dim=4096,
        num_auxiliary_layers=2, num_unimodal_layers=12, vocabulary_size=32000,
        atten_logit_cap=50.0),
}


def tiny_config(kind: str = "encoder", **over) -> dict:
    """The tiny shapes the reference unit tests use (encoders_test.py:129-158,
    :300-336) scaled so that dim_per_head stays a multiple of 16."""
    cfg = dict(kind=kind, patch_size=4, pos_emb_shape=(4, 4, 4), model_dim=64,
               num_spatial_layers=2, num_temporal_layers=2, num_heads=2, mlp_dim=128,
               atten_logit_cap=50.0)
    if kind == "clip":
        cfg.update(num_auxiliary_layers=1, num_unimodal_layers=2, vocabulary_size=128)
    if kind == "classifier":
        cfg.update(num_classes=10)
    cfg.update(over)
    return cfg


# ----------------------------------------------------------------------------
# Parameter tree (reference: SURVEY.md §3.4; names from encoders.py / layers.py;
# "repeated" = nn.scan-stacked leading [L] axis, layers.py:925-936,
# convert_weights.py:188-198; '/'-joined keys utils.py:84-105)
# ----------------------------------------------------------------------------
def _stack_specs(prefix: str, L: int, D: int, H: int, F: int, norm_policy: str = "pre") -> List[Tuple[str, tuple, str]]:
    dh = D // H
    p = prefix + "/x_layers"
    if norm_policy == "primer_hybrid":
        base = _stack_specs(prefix, L, D, H, F)
        out = []
        for key, shape, kind in base:
            if "/layer_norm/" in key:
                out.append((key.replace("/layer_norm/", "/pre_layer_norm/"), shape, kind))
                if key.endswith("/bias"):   # after the pre_layer_norm pair of this sub-layer
                    out.append((key.replace("/layer_norm/bias", "/post_layer_norm/scale"), shape, "ln_scale"))
                    out.append((key.replace("/layer_norm/bias", "/post_layer_norm/bias"), shape, "bias"))
            else:
                out.append((key, shape, kind))
        return out
    return [
        (p + "/layer_norm/scale", (L, D), "ln_scale"),
        (p + "/layer_norm/bias", (L, D), "bias"),
        (p + "/self_attention/query/w", (L, D, H, dh), "matrix"),
        (p + "/self_attention/query/b", (L, H, dh), "bias"),
        (p + "/self_attention/key/w", (L, D, H, dh), "matrix"),
        (p + "/self_attention/key/b", (L, H, dh), "bias"),
        (p + "/self_attention/value/w", (L, D, H, dh), "matrix"),
        (p + "/self_attention/value/b", (L, H, dh), "bias"),
        (p + "/self_attention/post/w", (L, D, H, dh), "matrix"),
        (p + "/self_attention/post/b", (L, D), "bias"),
        (p + "/ff_layer/layer_norm/scale", (L, D), "ln_scale"),
        (p + "/ff_layer/layer_norm/bias", (L, D), "bias"),
        (p + "/ff_layer/ffn_layer1/linear/kernel", (L, D, F), "matrix"),
        (p + "/ff_layer/ffn_layer1/linear/bias", (L, F), "bias"),
        (p + "/ff_layer/ffn_layer2/linear/kernel", (L, F, D), "matrix"),
        (p + "/ff_layer/ffn_layer2/linear/bias", (L, D), "bias"),
    ]


def _encoder_specs(prefix: str, cfg: dict) -> List[Tuple[str, tuple, str]]:
    D, H, F, P = cfg["model_dim"], cfg["num_heads"], cfg["mlp_dim"], cfg["patch_size"]
    tp, hp, wp = cfg["pos_emb_shape"]
    s = [
        (prefix + "/patch_projection/linear/kernel", (P * P * 3, D), "matrix"),
        (prefix + "/patch_projection/linear/bias", (D,), "bias"),
        (prefix + "/spatial_pos_emb/emb_var", (hp * wp, D), "matrix"),
        (prefix + "/temporal_pos_emb/emb_var", (tp, D), "matrix"),
    ]
    s += _stack_specs(prefix + "/spatial_encoder/transformers_stack", cfg["num_spatial_layers"], D, H, F)
    s += _stack_specs(prefix + "/temporal_encoder/transformers_stack", cfg["num_temporal_layers"], D, H, F)
    s += [
        (prefix + "/spatial_ln/scale", (D,), "ln_scale"),
        (prefix + "/spatial_ln/bias", (D,), "bias"),
        (prefix + "/temporal_ln/scale", (D,), "ln_scale"),
        (prefix + "/temporal_ln/bias", (D,), "bias"),
    ]
    return s


def _pooler_specs(pp: str, D: int, H: int, ph: int) -> List[Tuple[str, tuple, str]]:
    """AttenTokenPoolingLayer leaves (layers.py:1044-1136; 12 leaves, layers_test.py:282)."""
    return [
        (pp + "/pooling_attention_query", (1, D), "matrix"),
        (pp + "/pooling_attention/query/w", (D, H, ph), "matrix"),
        (pp + "/pooling_attention/query/b", (H, ph), "bias"),
        (pp + "/pooling_attention/key/w", (D, H, ph), "matrix"),
        (pp + "/pooling_attention/key/b", (H, ph), "bias"),
        (pp + "/pooling_attention/value/w", (D, H, ph), "matrix"),
        (pp + "/pooling_attention/value/b", (H, ph), "bias"),
        (pp + "/pooling_attention/post/w", (D, H, ph), "matrix"),
        (pp + "/pooling_attention/post/b", (D,), "bias"),
        (pp + "/pooling_attention/per_dim_scale/per_dim_scale", (ph,), "pds"),
        (pp + "/pooling_attention_layer_norm/scale", (D,), "ln_scale"),
        (pp + "/pooling_attention_layer_norm/bias", (D,), "bias"),
    ]


def param_specs(cfg: dict) -> List[Tuple[str, tuple, str]]:
    """Ordered (flax_key, shape, kind) list for a config."""
    if cfg["kind"] == "encoder":
        return _encoder_specs("params", cfg)
    D, H, F = cfg["model_dim"], cfg["num_heads"], cfg["mlp_dim"]
    if cfg["kind"] == "classifier":
        # FactorizedVideoClassifier (encoders.py:583-653): encoder, atten_pooler with hidden_dim = model_dim (:631-638,
        # so dim_per_head = D/H), projection Dense to num_classes (:643-650)
        s = _encoder_specs("params/encoder", cfg)
        s += _pooler_specs("params/atten_pooler", D, H, D // H)
        s += [("params/projection/linear/kernel", (D, cfg["num_classes"]), "matrix"),
              ("params/projection/linear/bias", (cfg["num_classes"],), "bias")]
        return s
    s = _encoder_specs("params/vision_encoder", cfg)
    if cfg["num_auxiliary_layers"] > 0:
        s += _stack_specs("params/auxiliary_encoder/transformers_stack", cfg["num_auxiliary_layers"], D, H, F)
    ph = 4 * D // H  # pooler: hidden_dim = 4*D (encoders.py:861), dim_per_head = hidden/H (layers.py:708-713)
    s += _pooler_specs("params/contrastive_vision_pooler", D, H, ph)
    tp = "params/text_encoder"
    s += [
        (tp + "/token_emb/emb_var", (cfg["vocabulary_size"], D), "matrix"),
        (tp + "/cls_emb", (1, 1, D), "matrix"),
    ]
    # text tower: mlp_dim = 4*model_dim (encoders.py:897)
    s += _stack_specs(tp + "/unimodal_transformer", cfg["num_unimodal_layers"], D, H, 4 * D, cfg.get("norm_policy", "pre"))
    s += [
        (tp + "/unimodal_ln/scale", (D,), "ln_scale"),
        (tp + "/unimodal_ln/bias", (D,), "bias"),
    ]
    return s


def make_synthetic_weights(cfg: dict, seed: int = 1234) -> Dict[str, np.ndarray]:
    """Random-init fp32 weights in the Flax key layout.

    matrices / embeddings N(0, 0.02); biases N(0, 0.02); LayerNorm scale N(0, 0.1)
    (effective 1 + scale); per_dim_scale N(0, 0.1).  Deliberately richer than
    Flax's own init (zeros for biases / LN scale, layers.py:248-266) so that
    bias / scale bugs are visible.
    """
    rng = np.random.default_rng(seed)
    std = {"matrix": 0.02, "bias": 0.02, "ln_scale": 0.1, "pds": 0.1}
    out = {}
    for key, shape, kind in param_specs(cfg):
        out[key] = (rng.standard_normal(shape, dtype=np.float32) * np.float32(std[kind])).astype(np.float32)
    return out


def make_video(batch: int, frames: int = 16, size: int = 288, seed: int = 0, kind: str = "uniform") -> np.ndarray:
    """Synthetic clips: uniform [0,1) (documented input range,
    README.md:171) or N(0, 0.1) as the reference tests use (models_test.py:39-41)."""
    rng = np.random.default_rng(seed)
    shape = (batch, frames, size, size, 3)
    if kind == "uniform":
        return rng.random(shape, dtype=np.float32)
    return (rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1)).astype(np.float32)


def make_text(queries: int, vocab: int = 32000, max_len: int = TEXT_MAX_LEN,
              seed_ids: int = 2, seed_len: int = 3) -> Tuple[np.ndarray, np.ndarray]:
    """Synthetic token ids / paddings in the format of models.tokenize_texts
    (models.py:385-407): right-padded, ids past the length 0, paddings 1.0."""
    ids = np.random.default_rng(seed_ids).integers(1, vocab, (queries, max_len), dtype=np.int32)
    lens = np.random.default_rng(seed_len).integers(4, 33, (queries,))
    pos = np.arange(max_len)[None, :]
    pad = (pos >= lens[:, None]).astype(np.float32)
    ids = np.where(pad > 0, 0, ids).astype(np.int32)
    return ids, pad


# ----------------------------------------------------------------------------
# Layers (reference: videoprism/layers.py)
# ----------------------------------------------------------------------------
def _neg(dtype: torch.dtype) -> float:
    """layers.py:39-48 `_get_large_negative_number`: -0.7 * finfo(dtype).max."""
    return -0.7 * torch.finfo(dtype).max


def layer_norm(x: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """layers.py:237-270: biased variance, eps inside rsqrt, (1 + scale), + bias."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    y = (x - mean) * torch.rsqrt(var + eps)
    return y * (scale + 1.0) + bias


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """layers.py:31: jax.nn.gelu(approximate=False)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def paddings_to_mask(paddings: torch.Tensor) -> torch.Tensor:
    """layers.py:75-89: [B,S] -> [B,1,1,S] additive mask."""
    return paddings[:, None, None, :] * _neg(paddings.dtype)


def causal_mask(S: int, dtype: torch.dtype) -> torch.Tensor:
    """layers.py:92-108: [1,1,S,S], (row < col) * large_negative."""
    idx = torch.arange(S)
    m = (idx[:, None] < idx[None, :]).to(dtype) * _neg(dtype)
    return m[None, None]


def merge_masks(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """layers.py:111-152: expand the 1-D mask to 2-D via min(mask, mask^T), then min."""
    def expand_t(km):
        return torch.minimum(km.transpose(-1, -2), km)
    if a.shape[-2] != b.shape[-2]:
        if a.shape[-2] == 1:
            a = expand_t(a)
        else:
            assert b.shape[-2] == 1
            b = expand_t(b)
    return torch.minimum(a, b)


def attention_masks_for_fprop(x: torch.Tensor, paddings: torch.Tensor, causal: bool) -> torch.Tensor:
    """layers.py:155-179."""
    m = paddings_to_mask(paddings.to(x.dtype))
    if causal:
        m = merge_masks(m, causal_mask(x.shape[-2], x.dtype))
    return m


def dot_attention(q, k, v, mask, cap: float, per_dim_scale: Optional[torch.Tensor], hidden_dim: int):
    """layers.py:601-661 `_dot_atten` for q [B,T,N,H], k/v [B,S,N,H], mask [1|B,1,1|T,S].

    query scale: layers.py:569-584 (dh^-0.5 where dh = hidden_dim // num_heads, or
    PerDimScale :502-527); logits :596-599; cap :586-594 (before masking);
    fp32 softmax :651-654 (here: in the oracle's dtype, fp32 or fp64); context :660.
    """
    n_heads = q.shape[-2]
    if per_dim_scale is not None:
        dim = q.shape[-1]
        scale = (1.442695041 / math.sqrt(dim)) * torch.nn.functional.softplus(per_dim_scale)
        q = q * scale
    else:
        q = q * (hidden_dim // n_heads) ** -0.5
    logits = torch.einsum("btnh,bsnh->bnts", q, k)
    if cap and cap > 0.0:
        logits = cap * torch.tanh(logits / cap)
    # mask semantics layers.py:51-72; the large-negative constant is the float32 one
    # because the reference casts logits to float32 before masking (:651-653).
    neg32 = -0.7 * torch.finfo(torch.float32).max
    mask_thr = _neg(mask.dtype) * 0.5
    logits = torch.where(mask >= mask_thr, logits, torch.full_like(logits, neg32))
    probs = torch.softmax(logits, dim=-1)
    return torch.einsum("bnts,bsnh->btnh", probs, v), probs


def attention_layer(xq, xkv, p: dict, mask, cap: float, per_dim_scale=None, hidden_dim: Optional[int] = None):
    """layers.py:686-746 DotProductAttention.__call__ with AttentionProjection :455-499.
    p: query/key/value {w [D,N,H], b [N,H]}, post {w [D,N,H], b [D]}."""
    q = torch.einsum("btd,dnh->btnh", xq, p["query/w"]) + p["query/b"]
    k = torch.einsum("bsd,dnh->bsnh", xkv, p["key/w"]) + p["key/b"]
    v = torch.einsum("bsd,dnh->bsnh", xkv, p["value/w"]) + p["value/b"]
    n_heads, dh = p["query/w"].shape[-2:]
    hd = hidden_dim if hidden_dim is not None else n_heads * dh
    ctx, _ = dot_attention(q, k, v, mask, cap, per_dim_scale, hd)
    return torch.einsum("btnh,dnh->btd", ctx, p["post/w"]) + p["post/b"]


def transformer_block(x, p: dict, mask, paddings, cap: float, act) -> torch.Tensor:
    """layers.py:797-872 + TransformerFeedForward :371-430, norm_policy 'pre' (every released model) or 'primer_hybrid'
    (the giant video-text model's text tower, models.py:155), told apart by the parameter names as the reference names
    them ('layer_norm' vs 'pre_layer_norm' + 'post_layer_norm', layers.py:819-822,:846-849,:388-391,:414-417)."""
    primer = "pre_layer_norm/scale" in p
    pre = "pre_layer_norm" if primer else "layer_norm"
    n = layer_norm(x, p[pre + "/scale"], p[pre + "/bias"])
    att = {k[len("self_attention/"):]: v for k, v in p.items() if k.startswith("self_attention/")}
    a = attention_layer(n, n, att, mask, cap)                         # :827-844
    if primer:
        a = layer_norm(a, p["post_layer_norm/scale"], p["post_layer_norm/bias"])       # :846-847
    y = x + a                                                         # :855
    m = layer_norm(y, p["ff_layer/" + pre + "/scale"], p["ff_layer/" + pre + "/bias"])   # :388-391
    keep = (1.0 - paddings)[..., None]
    u = act(m @ p["ff_layer/ffn_layer1/linear/kernel"] + p["ff_layer/ffn_layer1/linear/bias"]) * keep   # :394-398
    w = (u @ p["ff_layer/ffn_layer2/linear/kernel"] + p["ff_layer/ffn_layer2/linear/bias"]) * keep      # :405-411
    if primer:
        w = layer_norm(w, p["ff_layer/post_layer_norm/scale"], p["ff_layer/post_layer_norm/bias"])     # :414-415
    return y + w                                                      # :425


def stacked_transformer(x, stack: dict, paddings, cap: float, act, causal: bool) -> torch.Tensor:
    """layers.py:989-1041 with the nn.scan over the leading [L] axis (:875-937)."""
    mask = attention_masks_for_fprop(x, paddings, causal)
    L = next(iter(stack.values())).shape[0]
    for l in range(L):
        x = transformer_block(x, {k: v[l] for k, v in stack.items()}, mask, paddings, cap, act)
    return x


# ----------------------------------------------------------------------------
# jax.image.resize(..., 'bilinear') (called at encoders.py:124-126, :157-161)
# ----------------------------------------------------------------------------
def _resize_weights(n_in: int, n_out: int, dtype) -> torch.Tensor:
    """Restates jax.image.scale_and_translate's weight matrix for the triangle
    kernel with antialias=True (jax default): half-pixel centres, kernel widened by
    1/scale when down-sampling, per-output renormalisation, samples outside the
    input zeroed.  [n_in, n_out]."""
    scale = n_out / n_in
    inv_scale = 1.0 / scale
    kernel_scale = max(inv_scale, 1.0)
    sample_f = (torch.arange(n_out, dtype=dtype) + 0.5) * inv_scale - 0.5
    x = (sample_f[None, :] - torch.arange(n_in, dtype=dtype)[:, None]).abs() / kernel_scale
    w = torch.clamp(1.0 - x, min=0.0)
    tot = w.sum(dim=0, keepdim=True)
    w = torch.where(tot.abs() > 1000.0 * float(np.finfo(np.float32).eps), w / torch.where(tot != 0, tot, torch.ones_like(tot)), torch.zeros_like(w))
    ok = ((sample_f >= -0.5) & (sample_f <= n_in - 0.5))[None, :]
    return torch.where(ok, w, torch.zeros_like(w))


def interpolate_emb_1d(emb: torch.Tensor, target_len: int) -> torch.Tensor:
    """encoders.py:107-128: emb [N,D] -> [target_len, D]."""
    return _resize_weights(emb.shape[0], target_len, emb.dtype).T @ emb


def interpolate_emb_2d(emb: torch.Tensor, src: Tuple[int, int], dst: Tuple[int, int]) -> torch.Tensor:
    """encoders.py:131-165: emb [H1*W1, D] -> [H2*W2, D] (separable)."""
    D = emb.shape[-1]
    g = emb.reshape(src[0], src[1], D)
    g = torch.einsum("hwd,hH->Hwd", g, _resize_weights(src[0], dst[0], emb.dtype))
    g = torch.einsum("hwd,wW->hWd", g, _resize_weights(src[1], dst[1], emb.dtype))
    return g.reshape(dst[0] * dst[1], D)


# ----------------------------------------------------------------------------
# Encoders (reference: videoprism/encoders.py)
# ----------------------------------------------------------------------------
def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """encoders.py:50-67."""
    return x / torch.sqrt((x * x).sum(dim=-1, keepdim=True) + eps)


def image_to_patch(x: torch.Tensor, p: int) -> torch.Tensor:
    """encoders.py:70-104: '... (m p)(n q) c -> ... (m n)(p q c)'."""
    B, H, W, C = x.shape
    assert H % p == 0 and W % p == 0
    m, n = H // p, W // p
    return x.reshape(B, m, p, n, p, C).permute(0, 1, 3, 2, 4, 5).reshape(B, m * n, p * p * C)


def _sub(tree: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    prefix = prefix.rstrip("/") + "/"
    return {k[len(prefix):]: v for k, v in tree.items() if k.startswith(prefix)}


def to_torch(weights: Dict[str, np.ndarray], dtype=torch.float32) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in weights.items()}


def encoder_forward(cfg: dict, W: Dict[str, torch.Tensor], video: torch.Tensor,
                    return_intermediate: bool | Collection[str] = False,
                    frame_paddings: Optional[torch.Tensor] = None,
                    prefix: str = "params") -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """FactorizedEncoder.__call__ + encode_with_patches (encoders.py:411-580).
    video [B,T,H,W,3] -> ([B, T*N, D], outputs)."""
    dtype = video.dtype
    P = _sub(W, prefix)
    B, T, H, Wd, C = video.shape
    assert H == Wd                                                         # :435
    cap = cfg["atten_logit_cap"]
    patches = image_to_patch(video.reshape(B * T, H, Wd, C), cfg["patch_size"])   # :436-439
    N = patches.shape[1]
    pp = None
    if frame_paddings is not None:                                         # :440-447
        assert tuple(frame_paddings.shape) == (B, T)
        pp = frame_paddings.reshape(B * T, 1).repeat(1, N).to(dtype)
    x = patches @ P["patch_projection/linear/kernel"] + P["patch_projection/linear/bias"]   # :488-494
    sp_shape = tuple(cfg["pos_emb_shape"][-2:])
    sp = P["spatial_pos_emb/emb_var"][: sp_shape[0] * sp_shape[1]]          # :499-505 (one-hot matmul == slice)
    g = (H // cfg["patch_size"], Wd // cfg["patch_size"])
    if sp_shape != g:
        sp = interpolate_emb_2d(sp, sp_shape, g)                           # :508-513
    x = x + sp[None]                                                       # :514
    zeros_s = torch.zeros(x.shape[:-1], dtype=dtype) if pp is None else pp  # :367-368
    x = stacked_transformer(x, _sub(P, "spatial_encoder/transformers_stack/x_layers"), zeros_s, cap, gelu_erf, False)
    x = layer_norm(x, P["spatial_ln/scale"], P["spatial_ln/bias"])          # :528-530
    spatial = x
    D = x.shape[-1]
    x = x.reshape(B, T, N, D).permute(0, 2, 1, 3).reshape(B * N, T, D)      # :535
    tpad = None
    if pp is not None:
        tpad = pp.reshape(B, T, N).permute(0, 2, 1).reshape(B * N, T)       # :537-540
    Tpos = cfg["pos_emb_shape"][0]
    te = P["temporal_pos_emb/emb_var"][:Tpos]                               # :543-550
    if Tpos != T:
        te = interpolate_emb_1d(te, T)                                     # :551-552
    x = x + te[None]                                                       # :553
    zeros_t = torch.zeros(x.shape[:-1], dtype=dtype) if tpad is None else tpad
    x = stacked_transformer(x, _sub(P, "temporal_encoder/transformers_stack/x_layers"), zeros_t, cap, gelu_erf, False)
    x = layer_norm(x, P["temporal_ln/scale"], P["temporal_ln/bias"])        # :567-569
    x = x.reshape(B, N, T, D).permute(0, 2, 1, 3).reshape(B, T * N, D)      # :570-572
    outs = {}
    if _contains(return_intermediate, "spatial_features"):                 # :575-578
        outs["spatial_features"] = spatial.reshape(B, T * N, D)
    return x, outs


def _contains(coll, key: str) -> bool:
    """encoders.py:36-47."""
    return coll if isinstance(coll, bool) else key in coll


def atten_token_pool(W: Dict[str, torch.Tensor], tokens: torch.Tensor, n_heads: int, hidden_dim: Optional[int] = None) -> torch.Tensor:
    """AttenTokenPoolingLayer (layers.py:1072-1136): 1 learned query, hidden 4*D unless given (:1088),
    PerDimScale on q, no logit cap, LayerNorm after.  tokens [B,S,D] -> [B,1,D]."""
    B, S, D = tokens.shape
    q = W["pooling_attention_query"][None].expand(B, -1, -1)               # :1093-1101
    mask = paddings_to_mask(torch.zeros(B, S, dtype=tokens.dtype))         # :1102-1104
    att = _sub(W, "pooling_attention")
    out = attention_layer(q, tokens, att, mask, 0.0,
                          per_dim_scale=att["per_dim_scale/per_dim_scale"], hidden_dim=hidden_dim or 4 * D)
    return layer_norm(out, W["pooling_attention_layer_norm/scale"], W["pooling_attention_layer_norm/bias"])  # :1124-1129


def sinusoidal_pos_emb(S: int, D: int, dtype) -> torch.Tensor:
    """PositionalEmbedding (encoders.py:240-266): [S, D] = [sin | cos]; the timescale
    arithmetic is float32 in the reference regardless of fprop dtype."""
    pos = torch.arange(S, dtype=torch.float32)
    nts = D // 2
    inc = math.log(10000.0 / 1.0) / max(float(nts) - 1.0, 1.0)
    inv = torch.exp(torch.arange(nts, dtype=torch.float32) * torch.tensor(-inc, dtype=torch.float32))
    st = pos[:, None] * inv[None, :]
    e = torch.cat([torch.sin(st), torch.cos(st)], dim=-1)
    if D % 2:
        e = torch.nn.functional.pad(e, (0, 1))
    return e.to(dtype)


def text_encoder_forward(cfg: dict, W: Dict[str, torch.Tensor], ids: torch.Tensor, paddings: torch.Tensor,
                         prefix: str = "params/text_encoder") -> torch.Tensor:
    """TextEncoder.__call__ (encoders.py:693-759): ids [Q,L] int, paddings [Q,L] -> [Q, L+1, D]."""
    P = _sub(W, prefix)
    dtype = P["cls_emb"].dtype
    Q, L = ids.shape
    D = cfg["model_dim"]
    x = P["token_emb/emb_var"][ids.long()] * (D ** 0.5) + sinusoidal_pos_emb(L, D, dtype)[None]   # :708-722
    cls = P["cls_emb"].expand(Q, -1, -1) * (D ** 0.5)                       # :724-734
    x = torch.cat([x, cls], dim=1)                                         # :735
    pad = torch.cat([paddings.to(dtype), torch.zeros(Q, 1, dtype=dtype)], dim=1)   # :737-740
    x = stacked_transformer(x, _sub(P, "unimodal_transformer/x_layers"), pad, cfg["atten_logit_cap"], torch.relu, True)
    return layer_norm(x, P["unimodal_ln/scale"], P["unimodal_ln/bias"])     # :756-758


def clip_forward(cfg: dict, W: Dict[str, torch.Tensor], video: Optional[torch.Tensor] = None,
                 text_ids: Optional[torch.Tensor] = None, text_paddings: Optional[torch.Tensor] = None,
                 normalize: bool = True, return_intermediate: bool | Collection[str] = False,
                 frame_paddings: Optional[torch.Tensor] = None):
    """FactorizedVideoCLIP.__call__ (encoders.py:784-910)."""
    v_emb, t_emb, outs = None, None, {}
    if video is not None:
        T = video.shape[1]
        f, vo = encoder_forward(cfg, W, video, return_intermediate, frame_paddings, prefix="params/vision_encoder")
        outs.update(vo)
        if _contains(return_intermediate, "spatiotemporal_features"):      # :843-844
            outs["spatiotemporal_features"] = f
        if cfg["num_auxiliary_layers"] > 0:                                # :846-857
            aux = _sub(W, "params/auxiliary_encoder/transformers_stack/x_layers")
            f = stacked_transformer(f, aux, torch.zeros(f.shape[:-1], dtype=f.dtype), cfg["atten_logit_cap"], gelu_erf, False)
        pw = _sub(W, "params/contrastive_vision_pooler")
        v_emb = atten_token_pool(pw, f, cfg["num_heads"])[:, 0]             # :859-870
        if normalize:
            v_emb = l2_normalize(v_emb)                                    # :871-872
        if _contains(return_intermediate, "frame_embeddings"):            # :874-885
            B, TN, D = f.shape
            fe = atten_token_pool(pw, f.reshape(B * T, TN // T, D), cfg["num_heads"])[:, 0].reshape(B, T, D)
            if normalize:
                fe = l2_normalize(fe)
            outs["frame_embeddings"] = fe
    if text_ids is not None:
        assert text_paddings is not None                                   # :888
        tf = text_encoder_forward(cfg, W, text_ids, text_paddings)
        t_emb = tf[:, -1]                                                  # :906
        if normalize:
            t_emb = l2_normalize(t_emb)
    return v_emb, t_emb, outs


def classifier_forward(cfg: dict, W: Dict[str, torch.Tensor], video: torch.Tensor,
                       return_intermediate: bool | Collection[str] = False, frame_paddings: Optional[torch.Tensor] = None):
    """FactorizedVideoClassifier.__call__ (encoders.py:596-653)."""
    f, outs = encoder_forward(cfg, W, video, return_intermediate, frame_paddings, prefix="params/encoder")    # :616-627
    if _contains(return_intermediate, "spatiotemporal_features"):                                           # :628-629
        outs["spatiotemporal_features"] = f
    emb = atten_token_pool(_sub(W, "params/atten_pooler"), f, cfg["num_heads"], hidden_dim=cfg["model_dim"])[:, 0]  # :631-639
    if _contains(return_intermediate, "global_embeddings"):                                                 # :641-642
        outs["global_embeddings"] = emb
    logits = emb @ W["params/projection/linear/kernel"] + W["params/projection/linear/bias"]                # :643-650
    return logits, outs


# ----------------------------------------------------------------------------
# Convenience entry points used by tests / bench
# ----------------------------------------------------------------------------
def run_encoder(cfg: dict, weights: Dict[str, np.ndarray], video: np.ndarray, dtype=torch.float32, **kw):
    W = to_torch(weights, dtype)
    with torch.no_grad():
        prefix = "params" if cfg["kind"] == "encoder" else "params/vision_encoder"
        out, outs = encoder_forward(cfg, W, torch.from_numpy(video).to(dtype), prefix=prefix, **kw)
    return out.numpy(), {k: v.numpy() for k, v in outs.items()}


def run_clip(cfg: dict, weights: Dict[str, np.ndarray], video=None, ids=None, paddings=None, dtype=torch.float32, **kw):
    W = to_torch(weights, dtype)
    with torch.no_grad():
        v, t, outs = clip_forward(
            cfg, W,
            None if video is None else torch.from_numpy(video).to(dtype),
            None if ids is None else torch.from_numpy(ids),
            None if paddings is None else torch.from_numpy(paddings).to(dtype), **kw)
    return (None if v is None else v.numpy(), None if t is None else t.numpy(),
            {k: o.numpy() for k, o in outs.items()})


def run_classifier(cfg: dict, weights: Dict[str, np.ndarray], video: np.ndarray, dtype=torch.float32, **kw):
    W = to_torch(weights, dtype)
    with torch.no_grad():
        logits, outs = classifier_forward(cfg, W, torch.from_numpy(video).to(dtype), **kw)
    return logits.numpy(), {k: v.numpy() for k, v in outs.items()}


def count_params(cfg: dict) -> int:
    return sum(int(np.prod(s)) for _, s, _ in param_specs(cfg))
