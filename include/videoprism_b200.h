/* videoprism_b200 — C ABI of the B200-native VideoPrism forward path.
 *
 * The reference (tmoroney/videoprism-mlx) has no FFI: its hot path sits behind a Python call
 * surface.  Each entry point below names the reference interface it replaces; the Python shims in
 * `videoprism-mlx_b200/models.py` present that surface on top of this ABI.
 *
 * Conventions: every function returns 0 on success and a negative vp_status on failure (message via
 * vp_last_error); no exception crosses the ABI; the caller owns all I/O buffers and the stream; the
 * handle owns the repacked weights and its workspace; all device work is enqueued on the given
 * stream.  In steady state (same shapes as an earlier call) the device-buffer entry points and the
 * *_host_async ones do not synchronise with the host; the FIRST call with a new shape may: the
 * workspace grows with cudaMalloc / cudaFree and resized position tables are uploaded with a stream
 * synchronisation (so capture CUDA graphs only after one warm-up call per shape).  The synchronous
 * *_host variants wait once, for the device->host copy.  A handle is bound to one CUDA device (the current one at vp_create,
 * or the one named in vp_create_on_device); every entry point selects that device for the duration of
 * the call and restores the caller's current device before returning.  A handle is not thread-safe,
 * and because all forwards of a handle share its workspace, consecutive calls on one handle must be
 * ordered on the device (same stream, or streams the caller orders with events).  There is no CPU
 * fallback: without a CUDA device every compute call fails.
 */
#ifndef VIDEOPRISM_B200_H_
#define VIDEOPRISM_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define VP_API __attribute__((visibility("default")))
#else
#define VP_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vp_handle vp_handle;

typedef enum vp_status {
  VP_OK = 0,
  VP_ERR_INVALID = -1,     /* bad argument / shape (reference: assert / ValueError) */
  VP_ERR_KEY = -2,         /* unknown or wrongly shaped parameter key */
  VP_ERR_INCOMPLETE = -3,  /* vp_finalize / forward before all parameters were set */
  VP_ERR_CUDA = -4,        /* CUDA runtime / driver failure */
  VP_ERR_UNSUPPORTED = -5
} vp_status;

typedef enum vp_dtype { VP_F32 = 0, VP_BF16 = 1, VP_I32 = 2, VP_U8 = 3 } vp_dtype;

typedef enum vp_model_kind {
  VP_KIND_ENCODER = 0, /* encoders.FactorizedEncoder   (videoprism/encoders.py:391-580) */
  VP_KIND_CLIP = 1,    /* encoders.FactorizedVideoCLIP (videoprism/encoders.py:762-910) */
  VP_KIND_CLASSIFIER = 2 /* encoders.FactorizedVideoClassifier (videoprism/encoders.py:583-653) */
} vp_model_kind;

/* Mirrors the CONFIGS dict entries of videoprism/models.py:82-161 (and MODEL_CONFIGS,
 * videoprism/models_mlx.py:14-69). */
typedef struct vp_config {
  int kind;                /* vp_model_kind */
  int patch_size;          /* 18 */
  int pos_emb_t, pos_emb_h, pos_emb_w; /* pos_emb_shape */
  int model_dim;
  int num_spatial_layers;
  int num_temporal_layers;
  int num_heads;
  int mlp_dim;
  float atten_logit_cap;   /* 50.0 */
  int num_auxiliary_layers; /* CLIP only */
  int num_unimodal_layers;  /* CLIP only */
  int vocabulary_size;      /* CLIP only */
  int num_classes;          /* classifier only (videoprism/models.py:200-221) */
  int text_norm_policy;     /* CLIP text tower only: 0 = 'pre' (base / large), 1 = 'primer_hybrid' (giant, videoprism/models.py:155;
                               videoprism/layers.py:819-820,:846-847,:388-389,:414-415).  The vision stacks are always 'pre'
                               (videoprism/encoders.py:832,:853) */
} vp_config;

/* -- lifecycle: replaces models.get_model (videoprism/models.py:268-303) and
 *    models_mlx.load_video_encoder / load_model (videoprism/models_mlx.py:91-210) -------------- */
VP_API int vp_create(const vp_config* cfg, vp_handle** out);   /* on the calling thread's current CUDA device */
/* Same on an explicit device ordinal (device < 0: the current device).  Frameworks that select devices lazily (PyTorch
 * does not call cudaSetDevice for a device without a context yet) should use this one: the handle, its weights, its
 * workspace and every kernel it launches live on `device`; all buffers passed to its entry points must too. */
VP_API int vp_create_on_device(const vp_config* cfg, int device, vp_handle** out);
VP_API int vp_handle_device(const vp_handle* h);                /* device ordinal the handle is bound to */
/* fp32 CHECK MODE.  models.get_model(name) computes in float32 unless fprop_dtype says otherwise (videoprism/layers.py:182-205);
 * the production path of this library computes on the bf16 tensor cores.  A handle created with VP_FLAG_CHECK_FP32 (or any
 * handle of a process started with VP_CHECK_FP32=1) runs the SAME entry points entirely in float32 on the CUDA cores: fp32
 * residual stream, LayerNorm output, GEMM operands / accumulators, softmax and P.  It is ~50x slower and exists to hold the
 * implementation to the reference's own fp32 envelope (max-abs <= 1e-3 on features, <= 1e-5 on normalised embeddings:
 * FLAX_TO_MLX_CONVERSION_GUIDE.md:321-358, verify_clip_models.py:92-95).  Supports norm_policy 'pre' (every released model). */
#define VP_FLAG_CHECK_FP32 1u
VP_API int vp_create_ex(const vp_config* cfg, int device, unsigned flags, vp_handle** out);
VP_API int vp_handle_flags(const vp_handle* h);
VP_API void vp_destroy(vp_handle* h);
VP_API const char* vp_last_error(const vp_handle* h); /* h may be NULL: last error of a failed vp_create */

/* -- parameters: replaces models.load_pretrained_weights + utils.load_checkpoint / recover_tree
 *    (videoprism/models.py:306-336, videoprism/utils.py:84-105,:145-169).  `flax_key` is the
 *    '/'-joined key of the Flax checkpoint ("params/spatial_encoder/transformers_stack/x_layers/
 *    self_attention/query/w", scan-stacked leading [L] axis); `data` is fp32, host or device memory,
 *    C-contiguous with the given shape.  The handle repacks into its own bf16 / fp32 layouts
 *    immediately; `data` may be freed on return. */
VP_API int vp_set_weight(vp_handle* h, const char* flax_key, const void* data, const int64_t* shape, int ndim);
VP_API int vp_num_weights(const vp_handle* h);                    /* number of parameter leaves expected */
VP_API const char* vp_weight_key(const vp_handle* h, int index);  /* i-th expected flax key */
VP_API int vp_weight_ndim(const vp_handle* h, int index);
VP_API int64_t vp_weight_dim(const vp_handle* h, int index, int axis);
VP_API int vp_finalize(vp_handle* h);                             /* checks completeness, precomputes constants */

/* -- FactorizedEncoder.__call__ (videoprism/encoders.py:411-456, encode_with_patches :458-580).
 *    video [B,T,H,W,3] fp32 device memory, values as the reference expects ([0,1]).
 *    out_features [B, T*N, D] (out_dtype VP_F32 or VP_BF16).  spatial_features may be NULL, else
 *    receives outputs['spatial_features'] [B, T*N, D] in float32 (it needs out_dtype == VP_F32: asking
 *    for it together with VP_BF16 returns VP_ERR_UNSUPPORTED).  frame_paddings may be NULL, else
 *    [B,T] fp32 device memory (1 = padded frame).  `stream` is a cudaStream_t. */
VP_API int vp_encoder_forward(vp_handle* h, const float* video, int B, int T, int H, int W, const float* frame_paddings,
                       void* out_features, void* spatial_features, int out_dtype, void* stream);
/* Same call with HOST buffers (pageable or pinned): H2D copy of the clip batch, forward, D2H copy of
 * the features, one stream synchronisation at the end.  Outputs are fp32. */
VP_API int vp_encoder_forward_host(vp_handle* h, const float* video, int B, int T, int H, int W, const float* frame_paddings,
                            float* out_features, float* spatial_features, void* stream);

/* Asynchronous host-buffer call: enqueues the chunk-pipelined H2D copies / forward / D2H copies and RETURNS; *ticket names
 * the call for vp_wait, which blocks until its results are in out_features (and spatial_features).  Consecutive calls
 * pipeline on the device: the H2D of call k+1 overlaps the forward of call k, the D2H of call k the forward of call k+1
 * (this is how a serving loop hides PCIe entirely).  video: in_dtype VP_F32 or VP_U8; out_features: out_dtype VP_F32 or
 * VP_BF16 (half the D2H bytes; spatial_features is fp32 and requires VP_F32).  The host buffers must stay valid (and
 * should be page-locked) until vp_wait returns.  At most 8 calls may be outstanding.  Replaces the same reference call as
 * vp_encoder_forward (`model.apply(state, video, ...)`; JAX dispatch is asynchronous too, block_until_ready = vp_wait). */
VP_API int vp_encoder_forward_host_async(vp_handle* h, const void* video, int in_dtype, int B, int T, int H, int W,
                                  const float* frame_paddings, void* out_features, float* spatial_features, int out_dtype,
                                  void* stream, uint64_t* ticket);
VP_API int vp_wait(vp_handle* h, uint64_t ticket);

/* Same two calls for uint8 frames [B,T,H,W,3] (0..255) as cv2 decodes them: the `astype(float32) / 255.0` of
 * video_utils.load_video (videoprism/video_utils.py:88-93) happens on the device inside the patchify kernel,
 * bit-identically, so the host-to-device copy is 4x smaller. */
VP_API int vp_encoder_forward_u8(vp_handle* h, const uint8_t* video, int B, int T, int H, int W, const float* frame_paddings,
                          void* out_features, void* spatial_features, int out_dtype, void* stream);
VP_API int vp_encoder_forward_host_u8(vp_handle* h, const uint8_t* video, int B, int T, int H, int W, const float* frame_paddings,
                               float* out_features, float* spatial_features, void* stream);

/* -- FactorizedVideoCLIP.__call__ (videoprism/encoders.py:784-910), video side:
 *    vision_encoder -> auxiliary_encoder -> contrastive_vision_pooler -> (l2 normalise).
 *    video_emb [B,D] fp32.  Optional outputs (NULL to skip), all fp32: spatial_features [B,T*N,D],
 *    spatiotemporal_features [B,T*N,D], frame_embeddings [B,T,D]. */
VP_API int vp_clip_video_forward(vp_handle* h, const float* video, int B, int T, int H, int W, const float* frame_paddings,
                          int normalize, float* video_emb, float* spatial_features, float* spatiotemporal_features,
                          float* frame_embeddings, void* stream);
/* -- text side: text_encoder (TextEncoder, videoprism/encoders.py:656-759) -> last token -> (l2 normalise).
 *    ids [Q,L] int32, paddings [Q,L] fp32 (1 = pad), device memory; text_emb [Q,D] fp32. */
VP_API int vp_clip_text_forward(vp_handle* h, const int32_t* ids, const float* paddings, int Q, int L, int normalize,
                         float* text_emb, void* stream);
VP_API int vp_clip_video_forward_u8(vp_handle* h, const uint8_t* video, int B, int T, int H, int W, const float* frame_paddings,
                             int normalize, float* video_emb, float* spatial_features, float* spatiotemporal_features,
                             float* frame_embeddings, void* stream);   /* uint8 frames, / 255 on the device */
VP_API int vp_clip_video_forward_host(vp_handle* h, const float* video, int B, int T, int H, int W, int normalize,
                               float* video_emb, void* stream);
/* chunk-pipelined, asynchronous host-buffer form (see vp_encoder_forward_host_async): fp32 or uint8 frames, optional
 * frame_paddings [B,T] (host), pooled video embeddings [B,D] fp32 to host memory; vp_wait(ticket) for the result */
VP_API int vp_clip_video_forward_host_async(vp_handle* h, const void* video, int in_dtype, int B, int T, int H, int W,
                                     const float* frame_paddings, int normalize, float* video_emb, void* stream, uint64_t* ticket);
VP_API int vp_clip_text_forward_host(vp_handle* h, const int32_t* ids, const float* paddings, int Q, int L, int normalize,
                              float* text_emb, void* stream);

/* -- FactorizedVideoClassifier.__call__ (videoprism/encoders.py:596-653; models_mlx.load_classifier,
 *    videoprism/models_mlx.py:213-294): encoder -> atten_pooler (hidden_dim = model_dim) -> projection.
 *    logits [B, num_classes] fp32.  Optional outputs (NULL to skip), fp32: global_embeddings [B,D],
 *    spatial_features / spatiotemporal_features [B,T*N,D].  Device memory throughout. */
VP_API int vp_classifier_forward(vp_handle* h, const float* video, int B, int T, int H, int W, const float* frame_paddings,
                          float* logits, float* global_embeddings, float* spatial_features, float* spatiotemporal_features,
                          void* stream);

/* -- retrieval similarity (README.md:81; colab compute_similarity_matrix): sim[i,j] = v[i] . t[j].
 *    v [Nv,D], t [Nt,D], sim [Nv,Nt], fp32 device memory. */
VP_API int vp_similarity(const float* v, const float* t, float* sim, int Nv, int Nt, int D, void* stream);

/* -- introspection ------------------------------------------------------------------------------- */
VP_API size_t vp_workspace_bytes(const vp_handle* h, int B, int T, int H, int W); /* device bytes a forward of this shape needs */
/* Frees the activation workspace (it grows to the largest batch seen and is otherwise kept until vp_destroy); weights stay.
 * Waits for the handle's device to be idle.  The next forward allocates what it needs again. */
VP_API int vp_release_workspace(vp_handle* h);
VP_API int64_t vp_kernel_launches(const vp_handle* h);  /* kernels launched by this handle so far */
VP_API int vp_device_sm_count(void);                    /* < 0 when no CUDA device is usable */
/* In-situ timeline (diagnostic; the reference's counterpart is scripts/benchmark_performance.py's per-stage timers):
 * vp_trace(h, 1) starts recording one CUDA event after every kernel this handle launches (on the launch stream),
 * vp_trace_report synchronises and writes "label count total_ms" lines (+ "TOTAL n ms") into buf, vp_trace(h, 0) stops. */
VP_API int vp_trace(vp_handle* h, int enable);
VP_API int vp_trace_report(vp_handle* h, char* buf, int cap);

/* -- frame ingest (videoprism/video_utils.py:20-94 load_video, :97-127 _center_crop_resize), device memory:
 *    decoded RGB uint8 frames [T, H, W, 3] of any size -> uint8 [T, target_size, target_size, 3], bit-exact with the
 *    cv2.resize (INTER_LINEAR, fixed point) + centre crop the reference performs on the host.  resize_mode 0 =
 *    "center_crop" (shortest side -> target_size, then crop), 1 = "resize" (may distort).  The result feeds
 *    vp_encoder_forward_u8, which applies the /255 of video_utils.py:91. */
VP_API int vp_resize_frames_u8(const uint8_t* frames, int T, int H, int W, uint8_t* out, int target_size, int resize_mode,
                        void* stream);

/* -- kernel-level entry points (device pointers; used by the parity tests and micro-benchmarks) ---
 *    C[M,N] = A[M,K] * Wt[N,K]^T (+bias[N]) ; act 0 none, 1 exact GELU, 2 ReLU ; optional bf16 residual. */
VP_API int vp_gemm_bf16(const void* A, int lda, const void* Wt, int ldb, void* C, int ldc, int M, int N, int K,
                 const float* bias, int act, const void* resid, int ldr, const float* row_scale,
                 const float* pos_table, int pos_period, int out_f32, void* stream);
/* LayerNorm folded into a projection (how the engine runs every LN -> Dense pair of a Transformer block,
 * layers.py:822,:391 + :486-488,:304-312):  A holds RAW rows, Wt = (gamma1 (.) W)^T from vp_fold_ln_weight, and the
 * epilogue applies  rstd[m]*acc - rstd[m]*mean[m]*ln_colsum[n] + bias[n]  from per-row partial (sum, sum of
 * squares) ln_stats_in [M][ln_slots][2] (added in slot order).  stats_out (optional, [M][vp_gemm_stats_slots(N)][2])
 * receives the same partial statistics of the stored bf16 output rows, one slot per (column tile, epilogue half),
 * each written once: deterministic, no atomics. */
VP_API int vp_gemm_bf16_ln(const void* A, int lda, const void* Wt, int ldb, void* C, int ldc, int M, int N, int K,
                    const float* bias, int act, const void* resid, int ldr, const float* ln_stats_in, int ln_slots,
                    const float* ln_colsum, int ln_dim, float* stats_out, void* stream);
VP_API int vp_gemm_stats_slots(int N);
VP_API int vp_row_stats(const void* x, int ldx, float* stats, int M, int D, void* stream);
/* dst bf16 [N, ldk] = (gamma1[k] * src[k,n] * scale)^T ; colsum[n] = sum_k dst[n,k] ; bias_out = beta.src*scale + bias_in*scale */
VP_API int vp_fold_ln_weight(const float* src, const float* gamma1, const float* beta, const float* bias_in, void* dst,
                      float* colsum, float* bias_out, int K, int N, int ldk, float scale, void* stream);
/* y = LayerNorm(x) * gamma1 + beta over the last dim, x bf16 [M,D]; y_bf16 / y_f32 may each be NULL. */
VP_API int vp_layernorm(const void* x, int ldx, const float* gamma1, const float* beta, void* y_bf16, float* y_f32,
                 const float* add_table, int add_div, int add_mod, int M, int D, void* stream);
/* patches [BT*(H/p)*(W/p), ldo] bf16 from video [BT,H,W,3] fp32 */
VP_API int vp_patchify(const float* video, void* out, int ldo, int BT, int H, int W, int p, void* stream);
/* attention over a packed q|k|v buffer; see csrc/kernels.h AttnArgs for the row mapping.
 * causal: bit 0 = causal mask; bit 1 = force the mma.sync kernel (tests: bypass the tcgen05 S=256 kernel) */
VP_API int vp_attention(const void* q, const void* k, const void* v, int ld, void* out, int ldo, int num_seq, int S,
                 int group, int heads, int dh, float cap, const float* key_pad, int causal, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDEOPRISM_B200_H_ */
