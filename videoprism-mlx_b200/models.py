"""Reference-facing Python surface of the B200 path.

Mirrors, name for name, the two call surfaces of the reference:

* Flax style (`videoprism/models.py`): `get_model(name[, fprop_dtype])` (:268-303),
  `load_pretrained_weights(name[, checkpoint_path])` (:306-336), `has_model` (:255-265),
  `model.apply(state, video, train=False[, return_intermediate, frame_paddings])`
  (`encoders.py:411-456`) and, for the video-text models,
  `model.apply(state, video_or_None, text_ids_or_None, text_paddings_or_None, train=False,
  normalize=True, ...)` (`encoders.py:784-910`).
* MLX style (`videoprism/models_mlx.py`): `load_video_encoder(name[, weights_path])` (:146-210),
  `load_model(name[, weights_path])` (:91-143); the returned object is called directly.

All arithmetic happens in `libvideoprism_b200.so` (C ABI, `include/videoprism_b200.h`).  Inputs may be
numpy arrays (host buffers: the library does the H2D / D2H copies, results are numpy) or torch CUDA
tensors (device buffers: work is enqueued on torch's current stream, results are torch tensors).
"""
from __future__ import annotations

import ctypes as C
import functools
import os
from typing import Any, Callable, Collection, Dict, Mapping, Optional, Tuple

import numpy as np

from . import _lib
from . import tokenizers
from . import utils

K400_NUM_CLASSES: int = 400   # models.py:51-52
SSV2_NUM_CLASSES: int = 174
TEXT_MAX_LEN: int = 64  # models.py:54
TEXT_TOKENIZERS = {"c4_en": {"model_path": "gs://t5-data/vocabs/cc_en.32000/sentencepiece.model", "vocab_size": 32_000}}

# models.py:62-80
CHECKPOINTS = {
    "videoprism_public_v1_base": ("google/videoprism-base-f16r288", "flax_base_f16r288_repeated.npz"),
    "videoprism_public_v1_large": ("google/videoprism-large-f8r288", "flax_large_f8r288_repeated.npz"),
    "videoprism_lvt_public_v1_base": ("google/videoprism-lvt-base-f16r288", "flax_lvt_base_f16r288_repeated.npz"),
    "videoprism_lvt_public_v1_large": ("google/videoprism-lvt-large-f8r288", "flax_lvt_large_f8r288_repeated.npz"),
}

# models.py:82-161.  The giant configurations have no MODELS entry and no released checkpoint; the encoder / classifier run
# (dim_per_head = 88 goes through the generic attention kernels), the video-text giant needs norm_policy 'primer_hybrid'.
CONFIGS = {
    "videoprism_v1_base": dict(patch_size=18, pos_emb_shape=(16, 16, 16), model_dim=768, num_spatial_layers=12,
                               num_temporal_layers=4, num_heads=12, mlp_dim=3072, atten_logit_cap=50.0, scan=True),
    "videoprism_v1_large": dict(patch_size=18, pos_emb_shape=(8, 16, 16), model_dim=1024, num_spatial_layers=24,
                                num_temporal_layers=4, num_heads=16, mlp_dim=4096, atten_logit_cap=50.0, scan=True),
    "videoprism_v1_giant": dict(patch_size=18, pos_emb_shape=(8, 16, 16), model_dim=1408, num_spatial_layers=40,
                                num_temporal_layers=4, num_heads=16, mlp_dim=6144, atten_logit_cap=50.0, scan=True),
    "videoprism_lvt_v1_giant": dict(patch_size=18, pos_emb_shape=(8, 16, 16), num_spatial_layers=40, num_temporal_layers=4,
                                    mlp_dim=6144, num_auxiliary_layers=2, enable_causal_atten=True, num_unimodal_layers=16,
                                    norm_policy="primer_hybrid", model_dim=1408, num_heads=16, atten_logit_cap=50.0, scan=True),
    "videoprism_lvt_v1_base": dict(patch_size=18, pos_emb_shape=(16, 16, 16), num_spatial_layers=12, num_temporal_layers=4,
                                   mlp_dim=3072, num_auxiliary_layers=2, enable_causal_atten=True, num_unimodal_layers=12,
                                   norm_policy="pre", model_dim=768, num_heads=12, atten_logit_cap=50.0, scan=True),
    "videoprism_lvt_v1_large": dict(patch_size=18, pos_emb_shape=(8, 16, 16), num_spatial_layers=24, num_temporal_layers=4,
                                    mlp_dim=4096, num_auxiliary_layers=2, enable_causal_atten=True, num_unimodal_layers=12,
                                    norm_policy="pre", model_dim=1024, num_heads=16, atten_logit_cap=50.0, scan=True),
}


def _contains(collection, key: str) -> bool:
    """encoders.py:36-47."""
    return collection if isinstance(collection, bool) else key in collection


def _flatten(tree: Mapping[str, Any], prefix: str = "") -> Dict[str, np.ndarray]:
    """Nested param tree -> '/'-joined keys (inverse of utils.recover_tree, utils.py:84-105)."""
    out: Dict[str, np.ndarray] = {}
    for k, v in tree.items():
        key = f"{prefix}/{k}" if prefix else str(k)
        if isinstance(v, Mapping):
            out.update(_flatten(v, key))
        else:
            out[key] = v
    return out


def _recover_tree(flat: Mapping[str, Any]) -> Dict[str, Any]:
    """utils.recover_tree (utils.py:84-105): '/'-joined keys -> nested dict."""
    tree: Dict[str, Any] = {}
    for k, v in flat.items():
        node = tree
        parts = k.split("/")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = v
    return tree


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class _Module:
    """Common plumbing: owns the vp_handle, uploads a Flax-layout param tree once."""

    _kind = _lib.VP_KIND_ENCODER

    def __init__(self, **config):
        self.config = dict(config)
        # get_model(fprop_dtype=...): None (default) = the production path (bf16 tensor cores, fp32 accumulation, fp32
        # features out); bfloat16 = the same with bf16 features out; an EXPLICIT float32 selects the fp32 check mode
        self.fprop_dtype = None
        # fp32 check mode (vp_create_ex + VP_FLAG_CHECK_FP32): the whole forward in float32 on the CUDA cores, ~50x slower,
        # held to the reference's own fp32 envelope.  Also switched on for every model by VP_CHECK_FP32=1.
        self.check_fp32 = os.environ.get("VP_CHECK_FP32", "0") not in ("", "0")
        self._handle: Optional[C.c_void_p] = None
        self._loaded_state_id: Optional[int] = None
        self._loaded_state_ref = None
        self._loaded_fingerprint = None
        self._inflight: Dict[int, Any] = {}

    # -- handle -------------------------------------------------------------------------------
    def _vp_config(self) -> _lib.VpConfig:
        c = self.config
        t, h, w = c["pos_emb_shape"]
        return _lib.VpConfig(
            kind=self._kind, patch_size=c["patch_size"], pos_emb_t=t, pos_emb_h=h, pos_emb_w=w, model_dim=c["model_dim"],
            num_spatial_layers=c["num_spatial_layers"], num_temporal_layers=c["num_temporal_layers"], num_heads=c["num_heads"],
            mlp_dim=c["mlp_dim"], atten_logit_cap=float(c.get("atten_logit_cap", 0.0)),
            num_auxiliary_layers=int(c.get("num_auxiliary_layers", 0)), num_unimodal_layers=int(c.get("num_unimodal_layers", 0)),
            vocabulary_size=int(c.get("vocabulary_size", 0)), num_classes=int(c.get("num_classes", 0)),
            text_norm_policy={"pre": 0, "primer_hybrid": 1}[c.get("norm_policy", "pre")] if self._kind == _lib.VP_KIND_CLIP else 0)

    def _ensure_handle(self):
        if self._handle is None:
            policy = self.config.get("norm_policy", "pre")
            # the video-text model hands norm_policy to its TEXT tower only (encoders.py:899; vision stacks: 'pre', :832,:853)
            allowed = ("pre", "primer_hybrid") if self._kind == _lib.VP_KIND_CLIP else ("pre",)
            if policy not in allowed:
                raise NotImplementedError(f"norm_policy={policy!r} is not implemented (supported here: {allowed})")
            if self._kind == _lib.VP_KIND_CLIP and not self.config.get("enable_causal_atten", True):
                # encoders.py:682,:751: the text tower's StackedTransformer takes this flag; every released configuration sets
                # it (models.py:116-161) and the device path's text attention is causal, so the other value is refused loudly
                raise NotImplementedError("enable_causal_atten=False (a non-causal text tower) is not implemented")
            lib = _lib.lib()
            cfg = self._vp_config()
            h = C.c_void_p()
            # torch's device context is lazy (no cudaSetDevice before the device has a context), so the ordinal is passed
            # explicitly: the model lives on torch's current device at the time of its first use
            device = -1
            try:
                import torch
                if torch.cuda.is_available():
                    device = int(torch.cuda.current_device())
            except ImportError:
                pass
            d = self.fprop_dtype
            explicit_f32 = d is not None and "float32" in (getattr(d, "__name__", None) or str(d))
            flags = _lib.VP_FLAG_CHECK_FP32 if (self.check_fp32 or explicit_f32) else 0
            _lib.check(lib.vp_create_ex(C.byref(cfg), device, flags, C.byref(h)), None)
            self._handle = h
        return self._handle

    @property
    def device_index(self) -> int:
        """CUDA device ordinal the model (weights, workspace, kernels) lives on."""
        return int(_lib.lib().vp_handle_device(self._ensure_handle()))

    def _check_device(self, tensor) -> None:
        if tensor.device.index != self.device_index:
            raise ValueError(f"input is on cuda:{tensor.device.index} but the model lives on cuda:{self.device_index} "
                             "(create / load it under `with torch.cuda.device(...)` of the device it should run on)")

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                _lib.lib().vp_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def param_shapes(self) -> Dict[str, Tuple[int, ...]]:
        """Ordered {flax_key: shape} of the parameter leaves this model expects."""
        lib = _lib.lib()
        h = self._ensure_handle()
        out = {}
        for i in range(lib.vp_num_weights(h)):
            key = lib.vp_weight_key(h, i).decode()
            out[key] = tuple(int(lib.vp_weight_dim(h, i, a)) for a in range(lib.vp_weight_ndim(h, i)))
        return out

    def load_state(self, variables: Mapping[str, Any]) -> "_Module":
        """Uploads a Flax-layout param tree: nested {'params': {...}} (what load_pretrained_weights
        returns) or a flat {'params/...': array} mapping."""
        lib = _lib.lib()
        h = self._ensure_handle()
        flat = _flatten(variables) if any(isinstance(v, Mapping) for v in variables.values()) else dict(variables)
        expected = self.param_shapes()
        missing = [k for k in expected if k not in flat]
        if missing:
            raise KeyError(f"{len(missing)} parameters missing from the checkpoint, e.g. {missing[:3]}")
        for key in expected:
            arr = flat[key]
            if _is_torch(arr):
                arr = arr.detach().cpu().numpy()
            arr = np.ascontiguousarray(np.asarray(arr), dtype=np.float32)
            shape = (C.c_int64 * arr.ndim)(*arr.shape)
            _lib.check(lib.vp_set_weight(h, key.encode(), arr.ctypes.data_as(C.c_void_p), shape, arr.ndim), h)
        _lib.check(lib.vp_finalize(h), h)
        self._loaded_state_id = id(variables)
        self._loaded_state_ref = variables  # keeps id() stable
        self._loaded_fingerprint = self._fingerprint(variables)
        return self

    @staticmethod
    def _fingerprint(variables):
        """Identity of every leaf (key, object id, data pointer, shape): changes when a leaf of the tree is REPLACED, also in a
        dict that is mutated in place.  Element-wise writes into an uploaded array are not detected (that would mean hashing
        hundreds of MB per call): call load_state() again after such writes."""
        flat = _flatten(variables) if any(isinstance(v, Mapping) for v in variables.values()) else variables
        out = []
        for k in sorted(flat):
            v = flat[k]
            ptr = v.data_ptr() if _is_torch(v) else (v.ctypes.data if isinstance(v, np.ndarray) else 0)
            out.append((k, id(v), ptr, tuple(getattr(v, "shape", ()))))
        return tuple(out)

    def _bind(self, variables):
        if variables is None:
            return
        if id(variables) != self._loaded_state_id or self._fingerprint(variables) != self._loaded_fingerprint:
            self.load_state(variables)

    def apply(self, variables, *args, **kwargs):
        """Flax-style entry: `model.apply(state, ...)`.  The state is uploaded on first use."""
        kwargs.pop("rngs", None)
        kwargs.pop("mutable", None)
        self._bind(variables)
        return self(*args, **kwargs)

    def wait(self, ticket: int) -> None:
        """Blocks until the asynchronous host call `ticket` (forward_async / embed_video_async) has delivered its results."""
        _lib.check(_lib.lib().vp_wait(self._ensure_handle(), C.c_uint64(int(ticket))), self._handle)
        for k in [k for k in self._inflight if k <= ticket]:
            del self._inflight[k]

    def release_workspace(self) -> None:
        """Frees the activation workspace (sized for the largest batch seen so far); weights stay loaded.  The next call
        allocates again: for servers that see one large batch and then go back to small ones."""
        if self._handle is not None:
            _lib.check(_lib.lib().vp_release_workspace(self._handle), self._handle)

    @property
    def kernel_launches(self) -> int:
        return int(_lib.lib().vp_kernel_launches(self._handle)) if self._handle is not None else 0

    def trace(self, enable: bool = True) -> None:
        """Start / stop the in-situ kernel timeline (one CUDA event after every launch; vp_trace)."""
        _lib.check(_lib.lib().vp_trace(self._ensure_handle(), 1 if enable else 0), self._handle)

    def trace_report(self):
        """[(label, launches, total_ms)] since trace(True), in first-launch order, plus the ("TOTAL", n, ms) row."""
        buf = C.create_string_buffer(1 << 16)
        _lib.check(_lib.lib().vp_trace_report(self._handle, buf, len(buf)), self._handle)
        rows = [ln.split() for ln in buf.value.decode().splitlines() if ln.strip()]
        return [(r[0], int(r[1]), float(r[2])) for r in rows]

    # -- helpers --------------------------------------------------------------------------------
    @staticmethod
    def _stream_ptr() -> int:
        import torch
        return int(torch.cuda.current_stream().cuda_stream)

    def _wants_bf16(self) -> bool:
        """`get_model(name, fprop_dtype=jnp.bfloat16)` (models.py:283-301) makes the reference emit bfloat16 features; the
        device-buffer path honours that (the features leave the last LayerNorm in bf16, half the write traffic).  Host
        (numpy) results stay float32: numpy has no bfloat16."""
        d = self.fprop_dtype
        return d is not None and "bfloat16" in (getattr(d, "__name__", None) or str(d))

    def _check_video(self, inputs) -> Tuple[int, int, int, int]:
        if inputs.ndim != 5 or inputs.shape[-1] != 3:
            raise ValueError(f"inputs must be [B, T, H, W, 3], got {tuple(inputs.shape)}")
        b, t, hh, ww, _ = (int(s) for s in inputs.shape)
        if hh != ww:
            raise AssertionError("h == w")  # encoders.py:435
        p = self.config["patch_size"]
        if hh % p or ww % p:
            raise ValueError(f"Image height ({hh}) and width ({ww}) should be multiples of patch_size ({p}).")  # encoders.py:86-90
        return b, t, hh, ww


class FactorizedEncoder(_Module):
    """Drop-in for encoders.FactorizedEncoder (encoders.py:391-580) / encoders_mlx.FactorizedEncoder."""

    _kind = _lib.VP_KIND_ENCODER

    def __call__(self, inputs, train: bool = False, return_intermediate: bool | Collection[str] = False, frame_paddings=None,
                 out=None):
        """`out` (extension, host path only): a preallocated float32 C-contiguous `[B, T*N, D]` numpy array to receive
        the features, e.g. from `pinned_empty` so the device-to-host copy is a true asynchronous DMA."""
        del train  # dropout probabilities are 0 in every released config: train has no effect on the forward
        lib = _lib.lib()
        h = self._ensure_handle()
        b, t, hh, ww = self._check_video(inputs)
        p = self.config["patch_size"]
        n, d = (hh // p) * (ww // p), self.config["model_dim"]
        want_spatial = _contains(return_intermediate, "spatial_features")
        if frame_paddings is not None and tuple(frame_paddings.shape) != (b, t):
            raise AssertionError("frame_paddings.shape == (b, t)")  # encoders.py:442
        if _is_torch(inputs):
            import torch
            if not inputs.is_cuda:
                raise ValueError("torch inputs must live on a CUDA device (pass numpy arrays for host buffers)")
            self._check_device(inputs)
            is_u8 = inputs.dtype == torch.uint8     # raw decoded frames: / 255 happens on the device (video_utils.py:88-93)
            x = inputs.contiguous() if is_u8 else inputs.to(torch.float32).contiguous()
            fwd = lib.vp_encoder_forward_u8 if is_u8 else lib.vp_encoder_forward
            bf16_out = self._wants_bf16() and not want_spatial
            out = torch.empty((b, t * n, d), dtype=torch.bfloat16 if bf16_out else torch.float32, device=x.device)
            sp = torch.empty_like(out) if want_spatial else None
            fp = None if frame_paddings is None else torch.as_tensor(frame_paddings, device=x.device).to(torch.float32).contiguous()
            with torch.cuda.device(x.device):
                _lib.check(fwd(h, x.data_ptr(), b, t, hh, ww, None if fp is None else fp.data_ptr(),
                                                  out.data_ptr(), None if sp is None else sp.data_ptr(),
                                                  _lib.VP_BF16 if bf16_out else _lib.VP_F32, self._stream_ptr()), h)
            outs = {"spatial_features": sp} if want_spatial else {}
            return out, outs
        # blocking host call (vp_encoder_forward_host): its chunk schedule ramps at both ends, because nothing outside this call
        # overlaps its first copy in and its last copy out; forward_async below is the form whose calls overlap each other
        is_u8 = np.asarray(inputs).dtype == np.uint8
        x = np.ascontiguousarray(np.asarray(inputs), dtype=np.uint8 if is_u8 else np.float32)
        fwd_host = lib.vp_encoder_forward_host_u8 if is_u8 else lib.vp_encoder_forward_host
        if out is None:
            out = np.empty((b, t * n, d), dtype=np.float32)
        elif out.shape != (b, t * n, d) or out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError(f"out must be a C-contiguous float32 array of shape {(b, t * n, d)}")
        sp = np.empty_like(out) if want_spatial else None
        fp = None if frame_paddings is None else np.ascontiguousarray(np.asarray(frame_paddings), dtype=np.float32)
        _lib.check(fwd_host(
            h, x.ctypes.data_as(C.c_void_p), b, t, hh, ww, None if fp is None else fp.ctypes.data_as(C.c_void_p),
            out.ctypes.data_as(C.c_void_p), None if sp is None else sp.ctypes.data_as(C.c_void_p), None), h)
        outs = {"spatial_features": sp} if want_spatial else {}
        return out, outs

    def forward_async(self, inputs, return_intermediate: bool | Collection[str] = False, frame_paddings=None, out=None,
                      bf16_features: bool = False):
        """Host buffers only (extension; the reference's counterpart is JAX's asynchronous dispatch): enqueues the
        chunk-pipelined H2D copies, the forward and the D2H copies and returns `(ticket, out, outs)` at once; `wait(ticket)`
        (= block_until_ready) returns when `out` holds the features.  Calls pipeline on the device: the copies of one call
        overlap the forward of its neighbours, so a loop `t = forward_async(x[i], out=o[i % 2]); wait(prev); prev = t` is bound
        by the forward alone.  `inputs` (float32 or uint8 frames) and `out` should be page-locked (`pinned_empty`) and must not
        be touched before `wait` returns.  `bf16_features=True`: `out` is a uint16 array holding bfloat16 bit patterns (numpy
        has no bfloat16; `torch.from_numpy(out).view(torch.bfloat16)`): half the device-to-host bytes."""
        if _is_torch(inputs):
            raise ValueError("forward_async takes host (numpy) buffers; CUDA tensors are already asynchronous through __call__")
        b, t, hh, ww = self._check_video(inputs)
        lib = _lib.lib()
        h = self._ensure_handle()
        p = self.config["patch_size"]
        n, d = (hh // p) * (ww // p), self.config["model_dim"]
        want_spatial = _contains(return_intermediate, "spatial_features")
        if want_spatial and bf16_features:
            raise ValueError("spatial_features are float32: not available together with bf16_features")
        if frame_paddings is not None and tuple(frame_paddings.shape) != (b, t):
            raise AssertionError("frame_paddings.shape == (b, t)")  # encoders.py:442
        is_u8 = np.asarray(inputs).dtype == np.uint8
        x = np.ascontiguousarray(np.asarray(inputs), dtype=np.uint8 if is_u8 else np.float32)
        odt = np.uint16 if bf16_features else np.float32
        if out is None:
            out = np.empty((b, t * n, d), dtype=odt)
        elif out.shape != (b, t * n, d) or out.dtype != odt or not out.flags["C_CONTIGUOUS"]:
            raise ValueError(f"out must be a C-contiguous {np.dtype(odt).name} array of shape {(b, t * n, d)}")
        sp = np.empty_like(out) if want_spatial else None
        fp = None if frame_paddings is None else np.ascontiguousarray(np.asarray(frame_paddings), dtype=np.float32)
        ticket = C.c_uint64(0)
        _lib.check(lib.vp_encoder_forward_host_async(
            h, x.ctypes.data_as(C.c_void_p), _lib.VP_U8 if is_u8 else _lib.VP_F32, b, t, hh, ww,
            None if fp is None else fp.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
            None if sp is None else sp.ctypes.data_as(C.c_void_p), _lib.VP_BF16 if bf16_features else _lib.VP_F32, None, C.byref(ticket)), h)
        self._inflight[ticket.value] = (x, fp, out, sp)     # keeps the host buffers alive until wait()
        outs = {"spatial_features": sp} if want_spatial else {}
        return ticket.value, out, outs


class FactorizedVideoCLIP(_Module):
    """Drop-in for encoders.FactorizedVideoCLIP (encoders.py:762-910) / encoders_mlx.FactorizedVideoCLIP."""

    _kind = _lib.VP_KIND_CLIP

    def __call__(self, inputs=None, text_token_ids=None, text_paddings=None, train: bool = False, normalize: bool = True,
                 return_intermediate: bool | Collection[str] = False, frame_paddings=None):
        del train
        lib = _lib.lib()
        h = self._ensure_handle()
        d = self.config["model_dim"]
        video_emb, text_emb, outs = None, None, {}
        use_torch = _is_torch(inputs) or _is_torch(text_token_ids)
        if use_torch:
            import torch
        if inputs is not None:
            b, t, hh, ww = self._check_video(inputs)
            p = self.config["patch_size"]
            n = (hh // p) * (ww // p)
            names = [k for k in ("spatial_features", "spatiotemporal_features", "frame_embeddings") if _contains(return_intermediate, k)]
            shapes = {"spatial_features": (b, t * n, d), "spatiotemporal_features": (b, t * n, d), "frame_embeddings": (b, t, d)}
            if _is_torch(inputs):
                self._check_device(inputs)
                x = inputs.to(torch.float32).contiguous()
                video_emb = torch.empty((b, d), dtype=torch.float32, device=x.device)
                bufs = {k: torch.empty(shapes[k], dtype=torch.float32, device=x.device) for k in names}
                fp = None if frame_paddings is None else torch.as_tensor(frame_paddings, device=x.device).to(torch.float32).contiguous()
                ptr = lambda k: bufs[k].data_ptr() if k in bufs else None
                with torch.cuda.device(x.device):
                    _lib.check(lib.vp_clip_video_forward(
                        h, x.data_ptr(), b, t, hh, ww, None if fp is None else fp.data_ptr(), int(bool(normalize)), video_emb.data_ptr(),
                        ptr("spatial_features"), ptr("spatiotemporal_features"), ptr("frame_embeddings"), self._stream_ptr()), h)
                outs.update(bufs)
            else:
                if names:
                    # host path with intermediates: stage through torch device buffers
                    import torch
                    xv = torch.from_numpy(np.ascontiguousarray(np.asarray(inputs), dtype=np.float32)).cuda(self.device_index)
                    fpv = None if frame_paddings is None else torch.from_numpy(np.asarray(frame_paddings, dtype=np.float32)).cuda(self.device_index)
                    v, _, o = self(xv, None, None, normalize=normalize, return_intermediate=return_intermediate, frame_paddings=fpv)
                    video_emb = v.cpu().numpy()
                    outs.update({k: a.cpu().numpy() for k, a in o.items()})
                else:
                    if frame_paddings is None and np.asarray(inputs).dtype != np.uint8:
                        x = np.ascontiguousarray(np.asarray(inputs), dtype=np.float32)      # blocking call, ramped chunk schedule
                        video_emb = np.empty((b, d), dtype=np.float32)
                        _lib.check(lib.vp_clip_video_forward_host(h, x.ctypes.data_as(C.c_void_p), b, t, hh, ww, int(bool(normalize)),
                                                                  video_emb.ctypes.data_as(C.c_void_p), None), h)
                    else:                                                                    # uint8 frames / frame paddings
                        ticket, video_emb = self.embed_video_async(inputs, frame_paddings=frame_paddings, normalize=normalize)
                        self.wait(ticket)
        if text_token_ids is not None:
            assert text_paddings is not None, "Text paddings are required."  # encoders.py:888
            if text_token_ids.ndim != 2 or tuple(text_paddings.shape) != tuple(text_token_ids.shape):
                raise ValueError("text_token_ids and text_paddings must both be [Q, L]")
            q, length = (int(s) for s in text_token_ids.shape)
            if _is_torch(text_token_ids):
                self._check_device(text_token_ids)
                ids = text_token_ids.to(torch.int32).contiguous()
                pad = torch.as_tensor(text_paddings, device=ids.device).to(torch.float32).contiguous()
                text_emb = torch.empty((q, d), dtype=torch.float32, device=ids.device)
                with torch.cuda.device(ids.device):
                    _lib.check(lib.vp_clip_text_forward(h, ids.data_ptr(), pad.data_ptr(), q, length, int(bool(normalize)),
                                                        text_emb.data_ptr(), self._stream_ptr()), h)
            else:
                ids = np.ascontiguousarray(np.asarray(text_token_ids), dtype=np.int32)
                pad = np.ascontiguousarray(np.asarray(text_paddings), dtype=np.float32)
                text_emb = np.empty((q, d), dtype=np.float32)
                _lib.check(lib.vp_clip_text_forward_host(h, ids.ctypes.data_as(C.c_void_p), pad.ctypes.data_as(C.c_void_p), q, length,
                                                         int(bool(normalize)), text_emb.ctypes.data_as(C.c_void_p), None), h)
        return video_emb, text_emb, outs


    # (method of FactorizedVideoCLIP, attached below)
def _embed_video_async(self, inputs, frame_paddings=None, normalize: bool = True, out=None):
    """Host buffers only (extension): pooled video embeddings `[B, D]` of float32 or uint8 clips through the chunk-pipelined
    asynchronous entry point (vp_clip_video_forward_host_async); returns `(ticket, out)`, `wait(ticket)` for the result.
    Consecutive calls overlap their host-to-device copies with the neighbouring call's forward."""
    lib = _lib.lib()
    h = self._ensure_handle()
    b, t, hh, ww = self._check_video(inputs)
    d = self.config["model_dim"]
    if frame_paddings is not None and tuple(frame_paddings.shape) != (b, t):
        raise AssertionError("frame_paddings.shape == (b, t)")  # encoders.py:442
    is_u8 = np.asarray(inputs).dtype == np.uint8
    x = np.ascontiguousarray(np.asarray(inputs), dtype=np.uint8 if is_u8 else np.float32)
    if out is None:
        out = np.empty((b, d), dtype=np.float32)
    elif out.shape != (b, d) or out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"]:
        raise ValueError(f"out must be a C-contiguous float32 array of shape {(b, d)}")
    fp = None if frame_paddings is None else np.ascontiguousarray(np.asarray(frame_paddings), dtype=np.float32)
    ticket = C.c_uint64(0)
    _lib.check(lib.vp_clip_video_forward_host_async(
        h, x.ctypes.data_as(C.c_void_p), _lib.VP_U8 if is_u8 else _lib.VP_F32, b, t, hh, ww,
        None if fp is None else fp.ctypes.data_as(C.c_void_p), int(bool(normalize)), out.ctypes.data_as(C.c_void_p), None, C.byref(ticket)), h)
    self._inflight[ticket.value] = (x, fp, out)
    return ticket.value, out


FactorizedVideoCLIP.embed_video_async = _embed_video_async


class FactorizedVideoClassifier(_Module):
    """Drop-in for encoders.FactorizedVideoClassifier (encoders.py:583-653: `encoder_params=..., num_classes=...`) and
    encoders_mlx.FactorizedVideoClassifier (models_mlx.py:259: encoder keywords + `num_classes`).  Parameter tree:
    `params/encoder/...`, `params/atten_pooler/...` (hidden_dim = model_dim), `params/projection/linear/{kernel,bias}`."""

    _kind = _lib.VP_KIND_CLASSIFIER

    def __init__(self, encoder_params: Optional[Mapping[str, Any]] = None, num_classes: int = 0, **config):
        cfg = dict(encoder_params or {})
        cfg.update(config)
        if int(num_classes) <= 0:
            raise ValueError("num_classes must be a positive integer")
        cfg["num_classes"] = int(num_classes)
        super().__init__(**cfg)
        self.encoder_params = {k: v for k, v in cfg.items() if k != "num_classes"}
        self.num_classes = int(num_classes)

    def __call__(self, inputs, train: bool = False, return_intermediate: bool | Collection[str] = False, frame_paddings=None):
        del train
        lib = _lib.lib()
        h = self._ensure_handle()
        b, t, hh, ww = self._check_video(inputs)
        p = self.config["patch_size"]
        n, d = (hh // p) * (ww // p), self.config["model_dim"]
        if frame_paddings is not None and tuple(frame_paddings.shape) != (b, t):
            raise AssertionError("frame_paddings.shape == (b, t)")  # encoders.py:442
        import torch
        host = not _is_torch(inputs)
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(inputs), dtype=np.float32)).cuda(self.device_index) if host else inputs
        if not x.is_cuda:
            raise ValueError("torch inputs must live on a CUDA device (pass numpy arrays for host buffers)")
        self._check_device(x)
        x = x.to(torch.float32).contiguous()
        names = [k for k in ("spatial_features", "spatiotemporal_features", "global_embeddings") if _contains(return_intermediate, k)]
        shapes = {"spatial_features": (b, t * n, d), "spatiotemporal_features": (b, t * n, d), "global_embeddings": (b, d)}
        bufs = {k: torch.empty(shapes[k], dtype=torch.float32, device=x.device) for k in names}
        logits = torch.empty((b, self.num_classes), dtype=torch.float32, device=x.device)
        fp = None if frame_paddings is None else torch.as_tensor(np.asarray(frame_paddings) if host else frame_paddings,
                                                                 device=x.device).to(torch.float32).contiguous()
        ptr = lambda k: bufs[k].data_ptr() if k in bufs else None
        with torch.cuda.device(x.device):
            _lib.check(lib.vp_classifier_forward(h, x.data_ptr(), b, t, hh, ww, None if fp is None else fp.data_ptr(), logits.data_ptr(),
                                                 ptr("global_embeddings"), ptr("spatial_features"), ptr("spatiotemporal_features"),
                                                 self._stream_ptr()), h)
        if host:
            return logits.cpu().numpy(), {k: a.cpu().numpy() for k, a in bufs.items()}
        return logits, bufs


# ------------------------------------------------------------------------------ registry (models.py:164-233)
def videoprism_v1_base():
    return FactorizedEncoder(**CONFIGS["videoprism_v1_base"])


def videoprism_v1_large():
    return FactorizedEncoder(**CONFIGS["videoprism_v1_large"])


def videoprism_v1_giant():
    """models.py:174-176."""
    return FactorizedEncoder(**CONFIGS["videoprism_v1_giant"])


def videoprism_lvt_v1_giant(text_tokenizer: str = "c4_en"):
    """models.py:193-197 (text tower with norm_policy 'primer_hybrid', 16 blocks; no checkpoint was released)."""
    config = dict(CONFIGS["videoprism_lvt_v1_giant"])
    config["vocabulary_size"] = TEXT_TOKENIZERS[text_tokenizer]["vocab_size"]
    return FactorizedVideoCLIP(**config)


def videoprism_vc_v1_giant(num_classes: int):
    """models.py:216-221."""
    return FactorizedVideoClassifier(encoder_params=CONFIGS["videoprism_v1_giant"], num_classes=num_classes)


def videoprism_lvt_v1_base(text_tokenizer: str = "c4_en"):
    config = dict(CONFIGS["videoprism_lvt_v1_base"])
    config["vocabulary_size"] = TEXT_TOKENIZERS[text_tokenizer]["vocab_size"]
    return FactorizedVideoCLIP(**config)


def videoprism_lvt_v1_large(text_tokenizer: str = "c4_en"):
    config = dict(CONFIGS["videoprism_lvt_v1_large"])
    config["vocabulary_size"] = TEXT_TOKENIZERS[text_tokenizer]["vocab_size"]
    return FactorizedVideoCLIP(**config)


def videoprism_vc_v1_base(num_classes: int):
    """models.py:200-205."""
    return FactorizedVideoClassifier(encoder_params=CONFIGS["videoprism_v1_base"], num_classes=num_classes)


def videoprism_vc_v1_large(num_classes: int):
    """models.py:208-213."""
    return FactorizedVideoClassifier(encoder_params=CONFIGS["videoprism_v1_large"], num_classes=num_classes)


MODELS = {
    "videoprism_public_v1_base": videoprism_v1_base,
    "videoprism_public_v1_large": videoprism_v1_large,
    "videoprism_lvt_public_v1_base": functools.partial(videoprism_lvt_v1_base, text_tokenizer="c4_en"),
    "videoprism_lvt_public_v1_large": functools.partial(videoprism_lvt_v1_large, text_tokenizer="c4_en"),
}


def _get_model_name_by_hf_model_id(model_id: str) -> Optional[str]:
    """models.py:236-252."""
    for model_name, value in CHECKPOINTS.items():
        if isinstance(value, tuple) and value[0] == model_id:
            return model_name
    return None


def has_model(model_name: str, models: Optional[Mapping[str, Callable]] = None) -> bool:
    """models.py:255-265."""
    models = models or MODELS
    if model_name.startswith("google/"):
        model_name = _get_model_name_by_hf_model_id(model_name)
    return model_name is not None and model_name in models


def get_model(model_name: Optional[str], model_fn: Optional[Callable] = None, models: Optional[Mapping[str, Callable]] = None,
              fprop_dtype=None, check_fp32: Optional[bool] = None):
    """models.get_model (models.py:268-303): name (or HF id) -> configured module (no weights yet).
    `check_fp32=True` (extension; also `fprop_dtype=float32` given explicitly, or VP_CHECK_FP32=1) selects the fp32 check mode."""
    if model_fn is None:
        assert model_name is not None
        models = models or MODELS
        if model_name.startswith("google/"):
            resolved = _get_model_name_by_hf_model_id(model_name)
            if resolved is None:
                raise ValueError(f"Failed to find model name with `{model_name}`.")
            model_name = resolved
        if model_name not in models:
            raise ValueError(f"Model `{model_name}` not found.")
        model_fn = models[model_name]
    model = model_fn()
    if fprop_dtype is not None:
        model.fprop_dtype = fprop_dtype
    if check_fp32 is not None:
        model.check_fp32 = bool(check_fp32)
    return model


def load_checkpoint(checkpoint_path: str) -> Dict[str, Any]:
    """utils.load_checkpoint (utils.py:145-169): npz of '/'-joined keys -> nested tree of numpy arrays."""
    if not os.path.exists(checkpoint_path):
        raise FileNotFoundError(checkpoint_path)
    with np.load(checkpoint_path, allow_pickle=False) as data:
        flat = {k: data[k] for k in data.files}
    return _recover_tree(flat)


def load_pretrained_weights(model_name: Optional[str], checkpoint_path: Optional[str] = None,
                            checkpoints: Optional[Mapping[str, Any]] = None):
    """models.load_pretrained_weights (models.py:306-336).  Returns the nested {'params': ...} tree."""
    checkpoints = checkpoints or CHECKPOINTS
    if checkpoint_path is None:
        assert model_name is not None
        if model_name.startswith("google/"):
            model_name = _get_model_name_by_hf_model_id(model_name)
        repo_id, filename = checkpoints[model_name]
        import huggingface_hub  # needs network access; offline callers pass checkpoint_path
        checkpoint_path = huggingface_hub.hf_hub_download(repo_id=repo_id, filename=filename)
    return load_checkpoint(checkpoint_path)


def load_text_tokenizer(name: str) -> tokenizers.Tokenizer:
    """models.load_text_tokenizer (models.py:339-352)."""
    if name not in TEXT_TOKENIZERS:
        raise ValueError(f"Text tokenizer `{name}` not found.")
    return tokenizers.SentencePieceTokenizer(TEXT_TOKENIZERS[name]["model_path"])


def tokenize_texts(tokenizer: tokenizers.Tokenizer, inputs, max_length: int = TEXT_MAX_LEN, add_bos: Optional[bool] = None,
                   canonicalize: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """models.tokenize_texts (models.py:355-407): texts -> (ids [Q, max_length] int32 zero-padded, paddings [Q, max_length]
    float32 with 1 = padding), truncating long texts; the format `vp_clip_text_forward` takes."""
    if canonicalize:
        inputs = [utils.canonicalize_text(t) for t in inputs]
    if add_bos is None:
        add_bos = tokenizer.bos_token >= 0
    rows = tokenizer.to_int(list(inputs), bos=add_bos, eos=False)
    ids = np.zeros((len(rows), max_length), dtype=np.int32)
    paddings = np.ones((len(rows), max_length), dtype=np.float32)
    for r, toks in enumerate(rows):
        n = min(len(toks), max_length)
        ids[r, :n] = toks[:n]
        paddings[r, :n] = 0.0
    return ids, paddings


def synthetic_state(model, seed: int = 1234) -> Dict[str, np.ndarray]:
    """Random-init fp32 parameters in the Flax key layout for a model (or model name): matrices,
    embeddings and biases N(0, 0.02), LayerNorm scale N(0, 0.1) (effective 1 + scale),
    per_dim_scale N(0, 0.1).  No checkpoint is reachable offline."""
    if isinstance(model, str):
        model = get_model(model)
    rng = np.random.default_rng(seed)
    out = {}
    for key, shape in model.param_shapes().items():
        leaf = key.rsplit("/", 1)[-1]
        std = 0.1 if leaf in ("scale", "per_dim_scale") else 0.02
        out[key] = (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)
    return out


# ------------------------------------------------------------------------------ MLX-style loaders (models_mlx.py)
def _default_weights_path(model_name: str) -> str:
    for ext in (".npz",):
        p = os.path.join("weights", f"{model_name}{ext}")
        if os.path.exists(p):
            return p
    return os.path.join("weights", f"{model_name}.npz")


def load_video_encoder(model_name: str, weights_path: Optional[str] = None, state: Optional[Mapping[str, Any]] = None) -> FactorizedEncoder:
    """models_mlx.load_video_encoder (models_mlx.py:146-210): encoder with weights loaded; call it as
    `features, outputs = enc(video)`.  Video-text names are rejected as in the reference (:169-174)."""
    if "lvt" in model_name:
        raise ValueError(f"'{model_name}' is a video-text model. Use load_model() instead, or use a video encoder backbone like "
                         "'videoprism_public_v1_base'")
    if model_name not in MODELS:
        raise ValueError(f"Model '{model_name}' not found. Available models: {', '.join(MODELS)}")
    model = get_model(model_name)
    if state is None:
        path = weights_path or _default_weights_path(model_name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"Weights not found at {path}")  # models_mlx.py:191-196
        state = load_checkpoint(path)
    model.load_state(state)
    return model


def load_model(model_name: str, weights_path: Optional[str] = None, state: Optional[Mapping[str, Any]] = None) -> FactorizedVideoCLIP:
    """models_mlx.load_model (models_mlx.py:91-143): video-text model with weights loaded; call it as
    `video_emb, text_emb, outputs = model(video, text_ids, text_paddings)`."""
    if model_name not in MODELS:
        raise ValueError(f"Model '{model_name}' not found. Available models: {', '.join(MODELS)}")
    model = get_model(model_name)
    if not isinstance(model, FactorizedVideoCLIP):
        raise ValueError(f"'{model_name}' is a video encoder backbone. Use load_video_encoder() instead.")
    if state is None:
        path = weights_path or _default_weights_path(model_name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"Weights not found at {path}")
        state = load_checkpoint(path)
    model.load_state(state)
    return model



def get_model_config(model_name: str) -> dict:
    """models_mlx.get_model_config (models_mlx.py:72-88): a copy of the named configuration."""
    if model_name not in MODELS:
        raise ValueError(f"Model '{model_name}' not found. Available models: {', '.join(MODELS)}")
    return dict(get_model(model_name).config)


def load_classifier(model_name: str, num_classes: int, weights_path: Optional[str] = None,
                    state: Optional[Mapping[str, Any]] = None, seed: int = 0) -> FactorizedVideoClassifier:
    """models_mlx.load_classifier (models_mlx.py:213-294): encoder backbone of `model_name` + attention pooling + a
    `num_classes` projection.  As in the reference, a checkpoint only provides the ENCODER (`params/...` of an encoder
    checkpoint or `params/vision_encoder/...` of a video-text one, mapped to `params/encoder/...`); pooler and head are
    freshly initialised (Flax defaults: N(0, 0.02)-style kernels, zero biases / LayerNorm offsets) unless `state` holds them."""
    config = get_model_config(model_name)
    keys = ("patch_size", "pos_emb_shape", "model_dim", "num_spatial_layers", "num_temporal_layers", "num_heads", "mlp_dim",
            "atten_logit_cap")
    model = FactorizedVideoClassifier(encoder_params={k: config[k] for k in keys}, num_classes=num_classes)
    flat: Dict[str, Any] = {}
    if state is None:
        path = weights_path or _default_weights_path(model_name)
        if os.path.exists(path):
            state = load_checkpoint(path)
    if state is not None:
        src = _flatten(state) if any(isinstance(v, Mapping) for v in state.values()) else dict(state)
        for k, v in src.items():
            if k.startswith("params/vision_encoder/"):
                flat["params/encoder/" + k[len("params/vision_encoder/"):]] = v
            elif k.startswith(("params/encoder/", "params/atten_pooler/", "params/projection/")):
                flat[k] = v
            elif k.startswith(("params/patch_projection", "params/spatial_", "params/temporal_")):
                flat["params/encoder/" + k[len("params/"):]] = v
    rng = np.random.default_rng(seed)
    for key, shape in model.param_shapes().items():
        if key in flat:
            continue
        leaf = key.rsplit("/", 1)[-1]
        if leaf in ("bias", "b", "scale", "per_dim_scale"):
            flat[key] = np.zeros(shape, np.float32)
        else:
            flat[key] = (rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02)).astype(np.float32)
    model.load_state(flat)
    return model


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """A numpy array backed by page-locked host memory (cudaHostAlloc through torch): H2D / D2H copies of such
    buffers are asynchronous DMAs, which is what lets the host entry points overlap copies with compute.  The array's
    `.base` chain holds the torch tensor, so the pinned block is freed when the array is garbage collected."""
    import torch
    kinds = {np.float32: torch.float32, np.int32: torch.int32, np.uint8: torch.uint8, np.uint16: torch.uint16}
    t = torch.empty(tuple(shape), dtype=kinds[np.dtype(dtype).type], pin_memory=True)
    return t.numpy()


def compute_similarity_matrix(video_emb, text_emb):
    """README.md:81 / colab `compute_similarity_matrix`: sim = video_emb @ text_emb.T on the device."""
    import torch
    lib = _lib.lib()
    v = torch.as_tensor(video_emb).to(torch.float32).cuda().contiguous()
    t = torch.as_tensor(text_emb).to(torch.float32).cuda().contiguous()
    sim = torch.empty((v.shape[0], t.shape[0]), dtype=torch.float32, device=v.device)
    _lib.check(lib.vp_similarity(v.data_ptr(), t.data_ptr(), sim.data_ptr(), v.shape[0], t.shape[0], v.shape[1],
                                 int(torch.cuda.current_stream().cuda_stream)), None)
    return sim
