// Engine: owns the repacked weights and the workspace of one model on one device and enqueues the
// forward pass (encoder / CLIP video side / CLIP text side) as a fixed sequence of kernels on the
// caller's stream.  Exposes the C ABI declared in include/videoprism_b200.h.
//
// Activation layout: tokens stay in [B, T, N, D] order (row m = (b*T + t)*N + n) through BOTH the
// spatial and the temporal stack.  GEMMs and LayerNorms are per-token, so only the attention kernel
// needs to know which rows form a sequence (AttnArgs::group); the reference's einshape transposes
// (encoders.py:535, :570-572) are therefore never materialised.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/videoprism_b200.h"
#include "check_fp32.h"
#include "kernels.h"

namespace vp {
int num_sms();
cudaError_t launch_pool(cudaStream_t s, const bf16* x, int num_seq, int S, int D, int H, int dh, const bf16* wkq /*[32,D] hi|lo*/,
                        const bf16* wv /*[D, H*dh]*/, const float* bv, const bf16* wpost /*[D, H*dh]*/, const float* bpost,
                        const float* ln_g1, const float* ln_b, int normalize, float* scratch, float* out, int64_t* launches);
size_t pool_scratch_floats(int num_seq, int S, int D, int H, int dh);
cudaError_t launch_pad_expand(cudaStream_t s, const float* frame_pad, float* pad_tok, float* keep_tok, float* pad_tube, int B,
                              int T, int N);
cudaError_t launch_similarity(cudaStream_t s, const float* v, const float* t, float* sim, int Nv, int Nt, int D);
}  // namespace vp

using vp::bf16;

namespace {

std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  uint64_t* gen = nullptr;   // bumped whenever the buffer moves: captured CUDA graphs that hold the old pointer are stale
  // Grows (never shrinks) the buffer.  cudaFree / cudaMalloc synchronise with the device, so a buffer still in use by
  // queued work is never pulled away; `zero` clears a NEW allocation on the caller's stream `st`, i.e. ordered before
  // the kernels the caller enqueues next on that stream (a memset on the legacy default stream would not be ordered
  // against cudaStreamNonBlocking streams).
  cudaError_t ensure(size_t n, bool zero = false, cudaStream_t st = nullptr) {
    if (n <= bytes) return cudaSuccess;
    if (gen) ++*gen;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) return e;
    bytes = n;
    if (zero) e = cudaMemsetAsync(p, 0, n, st);
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; if (gen) ++*gen; }
};

// A forward captured as a CUDA graph: replaying it costs one launch instead of ~84 (+ ~250 tensor-map encodes), which is
// what a 1-clip forward is bound by.  The key holds everything the captured launches depend on (entry, shapes, dtypes and
// every caller pointer); `gen` is the workspace generation at capture time.
struct GraphEntry {
  std::vector<uint64_t> key;
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;
  uint64_t gen = 0, stamp = 0;
};

struct StackWeights {
  int L = 0, D = 0, H = 0, F = 0;
  bf16 *wqkv = nullptr, *wo = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *bqkv = nullptr, *bo = nullptr, *b1 = nullptr, *b2 = nullptr;
  float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  // LayerNorm folded into the QKV / FFN1 projections (see GemmEpilogue::ln_stats_in): (gamma1 (.) W)^T in bf16,
  // column sums of the rounded weights, and beta.W + b; built by finalize_stack from the fp32 staging copies below.
  bf16 *wqkv_ln = nullptr, *w1_ln = nullptr;
  float *c1_qkv = nullptr, *c2_qkv = nullptr, *c1_ffn1 = nullptr, *c2_ffn1 = nullptr;
  float *f_wqkv = nullptr, *f_bqkv = nullptr, *f_w1 = nullptr, *f_b1 = nullptr;   // fp32 [L][3][D][D], [L][3][D], [L][D][F], [L][F] (unscaled)
  // norm_policy 'primer_hybrid' (layers.py:819-820,:846-847,:388-389,:414-415): ln1 / ln2 are the 'pre_layer_norm's and the
  // attention / FFN outputs pass through these 'post_layer_norm's before the residual add
  bool primer = false;
  float *lnp1_g = nullptr, *lnp1_b = nullptr, *lnp2_g = nullptr, *lnp2_b = nullptr;
  // fp32 check mode only (check_fp32.cu): the two projections that are otherwise kept in bf16 alone.  post.w [L][D_out][D] is
  // [N, K]; ffn_layer2 kernel [L][F][D] is [K, N].  (q/k/v and ffn_layer1 use the f_* copies above.)
  float *c_wo = nullptr, *c_w2 = nullptr;
};

struct ParamSpec {
  std::string key;
  std::vector<int64_t> shape;
  std::function<cudaError_t(const float* staged, cudaStream_t s)> repack;
  bool set = false;
};

}  // namespace

struct vp_handle {
  vp_config cfg;
  int device = 0;
  std::string err;
  std::vector<ParamSpec> specs;
  std::unordered_map<std::string, int> spec_index;
  std::vector<void*> owned;  // device allocations holding weights
  bool finalized = false;
  int64_t launches = 0;
  DevBuf staging;
  // in-situ timeline (vp_trace): one CUDA event after every launch, labelled; off by default
  bool trace_on = false;
  std::vector<std::pair<const char*, cudaEvent_t>> trace;
  void mark(cudaStream_t st, const char* label, bool counts = true) {
    if (counts) launches++;
    if (!trace_on) return;
    cudaEvent_t ev = nullptr;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, st);
    trace.emplace_back(label, ev);
  }

  // encoder
  bf16* w_patch = nullptr; float* b_patch = nullptr; int k_patch = 0, k_patch_pad = 0;
  std::vector<float> h_spatial_pos, h_temporal_pos;  // host copies for (re)interpolation
  float* d_spatial_pos = nullptr; int spatial_grid_h = 0, spatial_grid_w = 0;
  bf16* d_spatial_pos_bf16 = nullptr;   // the same table in bf16: TMA-loaded like a residual tile by the patch projection
  float* d_temporal_pos = nullptr; int temporal_len = 0;
  StackWeights spatial, temporal, aux, text;
  std::vector<StackWeights*> stacks;   // the stacks this model has (for finalize_stack)
  bool fuse_ln = true;                 // LayerNorm folded into the QKV / FFN1 GEMMs (VP_FUSE_LN=0 disables)
  bool check_fp32 = false;             // VP_FLAG_CHECK_FP32: the whole forward in float32 on the CUDA cores (check_fp32.cu)
  float* c_wpatch = nullptr;           // check mode: patch projection kernel [k_patch, D] fp32
  float *c_pool_wkq = nullptr, *c_pool_wv = nullptr, *c_pool_wpost = nullptr;   // [D, H] / [D, H*ph] / [D, H*ph] fp32
  DevBuf c_x, c_n, c_qkv, c_u, c_patch, c_pool;                                   // fp32 workspace
  float *sp_ln_g = nullptr, *sp_ln_b = nullptr, *tp_ln_g = nullptr, *tp_ln_b = nullptr;
  // pooler (collapsed single-query form, see finalize_pooler)
  std::vector<float> h_pool_query, h_pool_wq, h_pool_bq, h_pool_wk, h_pool_pds;
  bf16* pool_wkq = nullptr;   // [32, D]: rows h = bf16(wkq[h]), rows H + h = bf16(wkq[h] - that): split-precision score weights
  bf16* pool_wv = nullptr; float* pool_bv = nullptr; bf16* pool_wpost = nullptr;
  float *pool_bpost = nullptr, *pool_ln_g = nullptr, *pool_ln_b = nullptr;
  int pool_ph = 0;                     // pooler dim_per_head: 4*D/H for the video-text pooler, D/H for the classifier's
  // classifier head (FactorizedVideoClassifier, encoders.py:643-650)
  float *cls_w = nullptr, *cls_b = nullptr;
  // text
  float* tok_emb = nullptr; float* cls_emb = nullptr; float* d_pe = nullptr; int pe_len = 0;
  float *uni_ln_g = nullptr, *uni_ln_b = nullptr;

  // workspace
  DevBuf ws_x, ws_n, ws_qkv, ws_u, ws_patch, ws_misc, ws_io_in, ws_io_out, ws_pool, ws_stats;
  // host-buffer pipeline (vp_encoder_forward_host)
  bool pipe_init = false;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  bool comp_rec[2] = {false, false}, out_rec[2] = {false, false};   // ev_comp / ev_out of the slot have been recorded at least once
  uint64_t chunk_seq = 0;     // chunks enqueued so far, over ALL host calls: slot = chunk_seq & 1, so consecutive calls pipeline
  size_t pipe_in_stride = 0, pipe_out_stride = 0;   // bytes between the two staging slots: only ever grow (see host_pipeline)
  int pipe_last_mode = -1;    // the output staging buffer is laid out per mode (features: two slots; embeddings: one block)
  static constexpr int kTickets = 8;
  cudaEvent_t ev_done[kTickets] = {};   // ev_done[t % kTickets]: results of the call with ticket t are in the caller's buffers
  uint64_t next_ticket = 1;
  int host_chunk_clips = 0;   // 0 = automatic
  // CUDA graphs (VP_GRAPHS=0 disables): the second call with a given key is captured, later ones replay
  bool use_graphs = true, capturing = false;
  uint64_t ws_generation = 0, graph_clock = 0;
  std::vector<GraphEntry> graphs;
  std::vector<std::vector<uint64_t>> graph_seen, graph_bad;
  cudaStream_t s_cap = nullptr;
  size_t stats_stride = 0;    // floats between the two LayerNorm-statistics buffers

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

namespace {

#define CK(call)                                                                                    \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return h->fail(VP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

template <typename T>
cudaError_t dev_alloc(vp_handle* h, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
  if (e != cudaSuccess) return e;
  h->owned.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return cudaSuccess;
}

void add_spec(vp_handle* h, const std::string& key, std::vector<int64_t> shape,
              std::function<cudaError_t(const float*, cudaStream_t)> fn) {
  ParamSpec s;
  s.key = key; s.shape = std::move(shape); s.repack = std::move(fn);
  h->spec_index[key] = static_cast<int>(h->specs.size());
  h->specs.push_back(std::move(s));
}

cudaError_t copy_f32(const float* src, float* dst, size_t n, cudaStream_t s) {
  return cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s);
}

// Parameter tree of one scan-stacked Transformer stack (layers.py:797-872).
cudaError_t add_stack(vp_handle* h, const std::string& prefix, StackWeights* w, int L, int D, int H, int F, bool primer = false) {
  w->L = L; w->D = D; w->H = H; w->F = F; w->primer = primer;
  const int dh = D / H;
  cudaError_t e;
  if ((e = dev_alloc(h, &w->wqkv, (size_t)L * 3 * D * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->wo, (size_t)L * D * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->w1, (size_t)L * F * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->w2, (size_t)L * D * F)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->bqkv, (size_t)L * 3 * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->bo, (size_t)L * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->b1, (size_t)L * F)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->b2, (size_t)L * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->ln1_g, (size_t)L * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->ln1_b, (size_t)L * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->ln2_g, (size_t)L * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->ln2_b, (size_t)L * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->wqkv_ln, (size_t)L * 3 * D * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->w1_ln, (size_t)L * F * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->c1_qkv, (size_t)L * 3 * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->c2_qkv, (size_t)L * 3 * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->c1_ffn1, (size_t)L * F)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->c2_ffn1, (size_t)L * F)) != cudaSuccess) return e;
  // fp32 copies of the projections that get a LayerNorm folded in; kept for the life of the handle so that
  // vp_finalize can re-fold after any later vp_set_weight (LN scale / bias and the kernels arrive in any order)
  if ((e = dev_alloc(h, &w->f_wqkv, (size_t)L * 3 * D * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->f_bqkv, (size_t)L * 3 * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->f_w1, (size_t)L * D * F)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &w->f_b1, (size_t)L * F)) != cudaSuccess) return e;
  if (h->check_fp32) {
    if ((e = dev_alloc(h, &w->c_wo, (size_t)L * D * D)) != cudaSuccess) return e;
    if ((e = dev_alloc(h, &w->c_w2, (size_t)L * F * D)) != cudaSuccess) return e;
  }
  if (primer) {
    if ((e = dev_alloc(h, &w->lnp1_g, (size_t)L * D)) != cudaSuccess) return e;
    if ((e = dev_alloc(h, &w->lnp1_b, (size_t)L * D)) != cudaSuccess) return e;
    if ((e = dev_alloc(h, &w->lnp2_g, (size_t)L * D)) != cudaSuccess) return e;
    if ((e = dev_alloc(h, &w->lnp2_b, (size_t)L * D)) != cudaSuccess) return e;
  }
  h->stacks.push_back(w);
  const std::string p = prefix + "/x_layers";
  const std::string pre = primer ? "pre_layer_norm" : "layer_norm";
  // query scale dh^-0.5 (layers.py:569-584, internal_enable_per_dim_scale=False) is folded into Wq, bq.
  const float qscale = 1.0f / sqrtf(static_cast<float>(dh));
  StackWeights ww = *w;
  add_spec(h, p + "/" + pre + "/scale", {L, D}, [ww, L, D](const float* s, cudaStream_t st) {
    return vp::launch_affine_f32(st, s, ww.ln1_g, (size_t)L * D, 1.0f, 1.0f); });
  add_spec(h, p + "/" + pre + "/bias", {L, D}, [ww, L, D](const float* s, cudaStream_t st) {
    return copy_f32(s, ww.ln1_b, (size_t)L * D, st); });
  if (primer) {
    add_spec(h, p + "/post_layer_norm/scale", {L, D}, [ww, L, D](const float* s, cudaStream_t st) {
      return vp::launch_affine_f32(st, s, ww.lnp1_g, (size_t)L * D, 1.0f, 1.0f); });
    add_spec(h, p + "/post_layer_norm/bias", {L, D}, [ww, L, D](const float* s, cudaStream_t st) {
      return copy_f32(s, ww.lnp1_b, (size_t)L * D, st); });
  }
  const char* names[3] = {"query", "key", "value"};
  for (int i = 0; i < 3; ++i) {
    const float sc = (i == 0) ? qscale : 1.0f;
    add_spec(h, p + "/self_attention/" + names[i] + "/w", {L, D, H, dh}, [ww, L, D, i, sc](const float* s, cudaStream_t st) {
      for (int l = 0; l < L; ++l) {   // [D, (H dh)] -> rows [i*D, (i+1)*D) of the fused [3D, D] K-major weight
        cudaError_t e = vp::launch_transpose_cast(st, s + (size_t)l * D * D, ww.wqkv + (size_t)l * 3 * D * D + (size_t)i * D * D,
                                                  D, D, D, sc);
        if (e != cudaSuccess) return e;
        if (ww.f_wqkv != nullptr) {
          e = copy_f32(s + (size_t)l * D * D, ww.f_wqkv + ((size_t)l * 3 + i) * D * D, (size_t)D * D, st);
          if (e != cudaSuccess) return e;
        }
      }
      return cudaSuccess; });
    add_spec(h, p + "/self_attention/" + names[i] + "/b", {L, H, dh}, [ww, L, D, i, sc](const float* s, cudaStream_t st) {
      for (int l = 0; l < L; ++l) {
        cudaError_t e = vp::launch_affine_f32(st, s + (size_t)l * D, ww.bqkv + (size_t)l * 3 * D + (size_t)i * D, D, sc, 0.f);
        if (e != cudaSuccess) return e;
        if (ww.f_bqkv != nullptr) {
          e = copy_f32(s + (size_t)l * D, ww.f_bqkv + ((size_t)l * 3 + i) * D, D, st);
          if (e != cudaSuccess) return e;
        }
      }
      return cudaSuccess; });
  }
  // post.w is [D_out, (H dh)]: already the K-major [N, K] layout the GEMM wants (layers.py:483).
  add_spec(h, p + "/self_attention/post/w", {L, D, H, dh}, [ww, L, D](const float* s, cudaStream_t st) {
    if (ww.c_wo != nullptr) { cudaError_t e = copy_f32(s, ww.c_wo, (size_t)L * D * D, st); if (e != cudaSuccess) return e; }
    return vp::launch_cast_bf16(st, s, ww.wo, (size_t)L * D * D, 1.0f); });
  add_spec(h, p + "/self_attention/post/b", {L, D}, [ww, L, D](const float* s, cudaStream_t st) {
    return copy_f32(s, ww.bo, (size_t)L * D, st); });
  add_spec(h, p + "/ff_layer/" + pre + "/scale", {L, D}, [ww, L, D](const float* s, cudaStream_t st) {
    return vp::launch_affine_f32(st, s, ww.ln2_g, (size_t)L * D, 1.0f, 1.0f); });
  add_spec(h, p + "/ff_layer/" + pre + "/bias", {L, D}, [ww, L, D](const float* s, cudaStream_t st) {
    return copy_f32(s, ww.ln2_b, (size_t)L * D, st); });
  if (primer) {
    add_spec(h, p + "/ff_layer/post_layer_norm/scale", {L, D}, [ww, L, D](const float* s, cudaStream_t st) {
      return vp::launch_affine_f32(st, s, ww.lnp2_g, (size_t)L * D, 1.0f, 1.0f); });
    add_spec(h, p + "/ff_layer/post_layer_norm/bias", {L, D}, [ww, L, D](const float* s, cudaStream_t st) {
      return copy_f32(s, ww.lnp2_b, (size_t)L * D, st); });
  }
  add_spec(h, p + "/ff_layer/ffn_layer1/linear/kernel", {L, D, F}, [ww, L, D, F](const float* s, cudaStream_t st) {
    for (int l = 0; l < L; ++l) {
      cudaError_t e = vp::launch_transpose_cast(st, s + (size_t)l * D * F, ww.w1 + (size_t)l * F * D, D, F, D, 1.0f);
      if (e != cudaSuccess) return e;
    }
    return ww.f_w1 != nullptr ? copy_f32(s, ww.f_w1, (size_t)L * D * F, st) : cudaSuccess; });
  add_spec(h, p + "/ff_layer/ffn_layer1/linear/bias", {L, F}, [ww, L, F](const float* s, cudaStream_t st) {
    cudaError_t e = copy_f32(s, ww.b1, (size_t)L * F, st);
    if (e != cudaSuccess) return e;
    return ww.f_b1 != nullptr ? copy_f32(s, ww.f_b1, (size_t)L * F, st) : cudaSuccess; });
  add_spec(h, p + "/ff_layer/ffn_layer2/linear/kernel", {L, F, D}, [ww, L, D, F](const float* s, cudaStream_t st) {
    for (int l = 0; l < L; ++l) {
      cudaError_t e = vp::launch_transpose_cast(st, s + (size_t)l * F * D, ww.w2 + (size_t)l * D * F, F, D, F, 1.0f);
      if (e != cudaSuccess) return e;
    }
    return ww.c_w2 != nullptr ? copy_f32(s, ww.c_w2, (size_t)L * F * D, st) : cudaSuccess; });
  add_spec(h, p + "/ff_layer/ffn_layer2/linear/bias", {L, D}, [ww, L, D](const float* s, cudaStream_t st) {
    return copy_f32(s, ww.b2, (size_t)L * D, st); });
  return cudaSuccess;
}

cudaError_t add_ln(vp_handle* h, const std::string& prefix, float** g, float** b, int D) {
  cudaError_t e;
  if ((e = dev_alloc(h, g, D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, b, D)) != cudaSuccess) return e;
  float* gg = *g; float* bb = *b;
  add_spec(h, prefix + "/scale", {D}, [gg, D](const float* s, cudaStream_t st) { return vp::launch_affine_f32(st, s, gg, D, 1.0f, 1.0f); });
  add_spec(h, prefix + "/bias", {D}, [bb, D](const float* s, cudaStream_t st) { return copy_f32(s, bb, D, st); });
  return cudaSuccess;
}

cudaError_t host_copy(std::vector<float>* dst, const float* staged, size_t n, cudaStream_t st) {
  dst->resize(n);
  cudaError_t e = cudaMemcpyAsync(dst->data(), staged, n * sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(st);
}

cudaError_t add_encoder(vp_handle* h, const std::string& prefix) {
  const vp_config& c = h->cfg;
  const int D = c.model_dim, P = c.patch_size;
  h->k_patch = P * P * 3;
  h->k_patch_pad = (h->k_patch + 63) / 64 * 64;
  cudaError_t e;
  if ((e = dev_alloc(h, &h->w_patch, (size_t)D * h->k_patch_pad)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &h->b_patch, D)) != cudaSuccess) return e;
  vp_handle* hh = h;
  if (h->check_fp32 && (e = dev_alloc(h, &h->c_wpatch, (size_t)h->k_patch * D)) != cudaSuccess) return e;
  add_spec(h, prefix + "/patch_projection/linear/kernel", {h->k_patch, D}, [hh, D](const float* s, cudaStream_t st) {
    if (hh->c_wpatch != nullptr) { cudaError_t e = copy_f32(s, hh->c_wpatch, (size_t)hh->k_patch * D, st); if (e != cudaSuccess) return e; }
    return vp::launch_transpose_cast(st, s, hh->w_patch, hh->k_patch, D, hh->k_patch_pad, 1.0f); });
  add_spec(h, prefix + "/patch_projection/linear/bias", {D}, [hh, D](const float* s, cudaStream_t st) {
    return copy_f32(s, hh->b_patch, D, st); });
  add_spec(h, prefix + "/spatial_pos_emb/emb_var", {(int64_t)c.pos_emb_h * c.pos_emb_w, D}, [hh, D](const float* s, cudaStream_t st) {
    hh->spatial_grid_h = hh->spatial_grid_w = 0;
    return host_copy(&hh->h_spatial_pos, s, (size_t)hh->cfg.pos_emb_h * hh->cfg.pos_emb_w * D, st); });
  add_spec(h, prefix + "/temporal_pos_emb/emb_var", {c.pos_emb_t, D}, [hh, D](const float* s, cudaStream_t st) {
    hh->temporal_len = 0;
    return host_copy(&hh->h_temporal_pos, s, (size_t)hh->cfg.pos_emb_t * D, st); });
  if ((e = add_stack(h, prefix + "/spatial_encoder/transformers_stack", &h->spatial, c.num_spatial_layers, D, c.num_heads, c.mlp_dim)) != cudaSuccess) return e;
  if ((e = add_stack(h, prefix + "/temporal_encoder/transformers_stack", &h->temporal, c.num_temporal_layers, D, c.num_heads, c.mlp_dim)) != cudaSuccess) return e;
  if ((e = add_ln(h, prefix + "/spatial_ln", &h->sp_ln_g, &h->sp_ln_b, D)) != cudaSuccess) return e;
  if ((e = add_ln(h, prefix + "/temporal_ln", &h->tp_ln_g, &h->tp_ln_b, D)) != cudaSuccess) return e;
  return cudaSuccess;
}

// AttenTokenPoolingLayer parameters (layers.py:1044-1136) under `pp`; ph = hidden_dim / num_heads.
cudaError_t add_pooler(vp_handle* h, const std::string& pp, int ph) {
  const int D = h->cfg.model_dim, H = h->cfg.num_heads;
  h->pool_ph = ph;
  cudaError_t e;
  vp_handle* hh = h;
  if ((e = dev_alloc(h, &h->pool_wkq, (size_t)32 * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &h->pool_wv, (size_t)H * ph * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &h->pool_bv, (size_t)H * ph)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &h->pool_wpost, (size_t)D * H * ph)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &h->pool_bpost, D)) != cudaSuccess) return e;
  add_spec(h, pp + "/pooling_attention_query", {1, D}, [hh, D](const float* s, cudaStream_t st) { return host_copy(&hh->h_pool_query, s, D, st); });
  add_spec(h, pp + "/pooling_attention/query/w", {D, H, ph}, [hh, D, H, ph](const float* s, cudaStream_t st) { return host_copy(&hh->h_pool_wq, s, (size_t)D * H * ph, st); });
  add_spec(h, pp + "/pooling_attention/query/b", {H, ph}, [hh, H, ph](const float* s, cudaStream_t st) { return host_copy(&hh->h_pool_bq, s, (size_t)H * ph, st); });
  add_spec(h, pp + "/pooling_attention/key/w", {D, H, ph}, [hh, D, H, ph](const float* s, cudaStream_t st) { return host_copy(&hh->h_pool_wk, s, (size_t)D * H * ph, st); });
  // key bias shifts every score of a head by the same constant -> cancels in the softmax; accepted, unused.
  add_spec(h, pp + "/pooling_attention/key/b", {H, ph}, [](const float*, cudaStream_t) { return cudaSuccess; });
  if (h->check_fp32) {
    if ((e = dev_alloc(h, &h->c_pool_wkq, (size_t)D * H)) != cudaSuccess) return e;
    if ((e = dev_alloc(h, &h->c_pool_wv, (size_t)D * H * ph)) != cudaSuccess) return e;
    if ((e = dev_alloc(h, &h->c_pool_wpost, (size_t)D * H * ph)) != cudaSuccess) return e;
  }
  add_spec(h, pp + "/pooling_attention/value/w", {D, H, ph}, [hh, D, H, ph](const float* s, cudaStream_t st) {
    if (hh->c_pool_wv != nullptr) { cudaError_t e = copy_f32(s, hh->c_pool_wv, (size_t)D * H * ph, st); if (e != cudaSuccess) return e; }
    return vp::launch_cast_bf16(st, s, hh->pool_wv, (size_t)D * H * ph, 1.0f); });  // kept [D, (H dh)]: pool_ctx_kernel reads it j-contiguous
  add_spec(h, pp + "/pooling_attention/value/b", {H, ph}, [hh, H, ph](const float* s, cudaStream_t st) { return copy_f32(s, hh->pool_bv, (size_t)H * ph, st); });
  add_spec(h, pp + "/pooling_attention/post/w", {D, H, ph}, [hh, D, H, ph](const float* s, cudaStream_t st) {
    if (hh->c_pool_wpost != nullptr) { cudaError_t e = copy_f32(s, hh->c_pool_wpost, (size_t)D * H * ph, st); if (e != cudaSuccess) return e; }
    return vp::launch_cast_bf16(st, s, hh->pool_wpost, (size_t)D * H * ph, 1.0f); });
  add_spec(h, pp + "/pooling_attention/post/b", {D}, [hh, D](const float* s, cudaStream_t st) { return copy_f32(s, hh->pool_bpost, D, st); });
  add_spec(h, pp + "/pooling_attention/per_dim_scale/per_dim_scale", {ph}, [hh, ph](const float* s, cudaStream_t st) { return host_copy(&hh->h_pool_pds, s, ph, st); });
  if ((e = add_ln(h, pp + "/pooling_attention_layer_norm", &h->pool_ln_g, &h->pool_ln_b, D)) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t add_clip_extras(vp_handle* h) {
  const vp_config& c = h->cfg;
  const int D = c.model_dim, H = c.num_heads;
  const int ph = 4 * D / H;  // pooler dim_per_head: hidden 4*D over H heads (encoders.py:861, layers.py:708-713)
  cudaError_t e;
  if (c.num_auxiliary_layers > 0) {
    if ((e = add_stack(h, "params/auxiliary_encoder/transformers_stack", &h->aux, c.num_auxiliary_layers, D, H, c.mlp_dim)) != cudaSuccess) return e;
  }
  if ((e = add_pooler(h, "params/contrastive_vision_pooler", ph)) != cudaSuccess) return e;
  vp_handle* hh = h;
  // text tower (encoders.py:656-759): mlp_dim = 4 * model_dim (:897)
  const std::string tp = "params/text_encoder";
  if ((e = dev_alloc(h, &h->tok_emb, (size_t)c.vocabulary_size * D)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &h->cls_emb, D)) != cudaSuccess) return e;
  add_spec(h, tp + "/token_emb/emb_var", {c.vocabulary_size, D}, [hh, D](const float* s, cudaStream_t st) {
    return copy_f32(s, hh->tok_emb, (size_t)hh->cfg.vocabulary_size * D, st); });
  add_spec(h, tp + "/cls_emb", {1, 1, D}, [hh, D](const float* s, cudaStream_t st) { return copy_f32(s, hh->cls_emb, D, st); });
  if ((e = add_stack(h, tp + "/unimodal_transformer", &h->text, c.num_unimodal_layers, D, H, 4 * D, c.text_norm_policy == 1)) != cudaSuccess) return e;
  if ((e = add_ln(h, tp + "/unimodal_ln", &h->uni_ln_g, &h->uni_ln_b, D)) != cudaSuccess) return e;
  return cudaSuccess;
}

// jax.image.resize(..., 'bilinear') weights (antialias=True default): triangle kernel, half-pixel
// centres, widened by 1/scale when down-sampling, renormalised per output (encoders.py:124-126,:157-161).
std::vector<float> resize_weights(int n_in, int n_out) {
  std::vector<float> w((size_t)n_in * n_out, 0.f);
  const double inv_scale = (double)n_in / n_out;
  const double kscale = inv_scale > 1.0 ? inv_scale : 1.0;
  for (int o = 0; o < n_out; ++o) {
    const double sf = (o + 0.5) * inv_scale - 0.5;
    double tot = 0.0;
    for (int i = 0; i < n_in; ++i) {
      double x = fabs(sf - i) / kscale;
      double v = x < 1.0 ? 1.0 - x : 0.0;
      w[(size_t)i * n_out + o] = (float)v;
      tot += v;
    }
    const bool inside = sf >= -0.5 && sf <= n_in - 0.5;
    for (int i = 0; i < n_in; ++i) {
      float& v = w[(size_t)i * n_out + o];
      v = (inside && fabs(tot) > 1000.0 * 1.1920929e-7) ? (float)(v / tot) : 0.f;
    }
  }
  return w;
}

int prepare_pos_tables(vp_handle* h, int T, int gh, int gw, cudaStream_t st) {
  const vp_config& c = h->cfg;
  const int D = c.model_dim;
  if (h->spatial_grid_h != gh || h->spatial_grid_w != gw) {
    std::vector<float> tab;
    if (gh == c.pos_emb_h && gw == c.pos_emb_w) {
      tab = h->h_spatial_pos;
    } else {  // _interpolate_emb_2d, encoders.py:131-165 (separable)
      std::vector<float> wh = resize_weights(c.pos_emb_h, gh), ww = resize_weights(c.pos_emb_w, gw);
      std::vector<float> tmp((size_t)gh * c.pos_emb_w * D, 0.f);
      for (int o = 0; o < gh; ++o)
        for (int i = 0; i < c.pos_emb_h; ++i) {
          const float w = wh[(size_t)i * gh + o];
          if (w == 0.f) continue;
          for (int x = 0; x < c.pos_emb_w * D; ++x) tmp[(size_t)o * c.pos_emb_w * D + x] += w * h->h_spatial_pos[(size_t)i * c.pos_emb_w * D + x];
        }
      tab.assign((size_t)gh * gw * D, 0.f);
      for (int y = 0; y < gh; ++y)
        for (int o = 0; o < gw; ++o)
          for (int i = 0; i < c.pos_emb_w; ++i) {
            const float w = ww[(size_t)i * gw + o];
            if (w == 0.f) continue;
            for (int d = 0; d < D; ++d) tab[((size_t)y * gw + o) * D + d] += w * tmp[((size_t)y * c.pos_emb_w + i) * D + d];
          }
    }
    ++h->ws_generation;   // the tables move: captured graphs are stale
    if (h->d_spatial_pos) { cudaFree(h->d_spatial_pos); h->d_spatial_pos = nullptr; }
    CK(cudaMalloc(&h->d_spatial_pos, tab.size() * sizeof(float)));
    CK(cudaMemcpyAsync(h->d_spatial_pos, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    std::vector<bf16> tab16(tab.size());
    for (size_t i = 0; i < tab.size(); ++i) tab16[i] = __float2bfloat16(tab[i]);
    if (h->d_spatial_pos_bf16) { cudaFree(h->d_spatial_pos_bf16); h->d_spatial_pos_bf16 = nullptr; }
    CK(cudaMalloc(&h->d_spatial_pos_bf16, tab16.size() * sizeof(bf16)));
    CK(cudaMemcpyAsync(h->d_spatial_pos_bf16, tab16.data(), tab16.size() * sizeof(bf16), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    h->spatial_grid_h = gh; h->spatial_grid_w = gw;
  }
  if (h->temporal_len != T) {
    std::vector<float> tab;
    if (T == c.pos_emb_t) {
      tab = h->h_temporal_pos;
    } else {  // _interpolate_emb_1d, encoders.py:107-128
      std::vector<float> wt = resize_weights(c.pos_emb_t, T);
      tab.assign((size_t)T * D, 0.f);
      for (int o = 0; o < T; ++o)
        for (int i = 0; i < c.pos_emb_t; ++i) {
          const float w = wt[(size_t)i * T + o];
          if (w == 0.f) continue;
          for (int d = 0; d < D; ++d) tab[(size_t)o * D + d] += w * h->h_temporal_pos[(size_t)i * D + d];
        }
    }
    ++h->ws_generation;
    if (h->d_temporal_pos) { cudaFree(h->d_temporal_pos); h->d_temporal_pos = nullptr; }
    CK(cudaMalloc(&h->d_temporal_pos, tab.size() * sizeof(float)));
    CK(cudaMemcpyAsync(h->d_temporal_pos, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    h->temporal_len = T;
  }
  return VP_OK;
}

// FactorizedVideoClassifier (encoders.py:583-653): encoder under 'params/encoder', AttenTokenPoolingLayer 'atten_pooler'
// with hidden_dim = model_dim (:631-638), FeedForward 'projection' to num_classes (:643-650).
cudaError_t add_classifier_extras(vp_handle* h) {
  const vp_config& c = h->cfg;
  const int D = c.model_dim, NC = c.num_classes;
  cudaError_t e;
  if ((e = add_pooler(h, "params/atten_pooler", D / c.num_heads)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &h->cls_w, (size_t)D * NC)) != cudaSuccess) return e;
  if ((e = dev_alloc(h, &h->cls_b, NC)) != cudaSuccess) return e;
  vp_handle* hh = h;
  add_spec(h, "params/projection/linear/kernel", {D, NC}, [hh, D, NC](const float* s, cudaStream_t st) { return copy_f32(s, hh->cls_w, (size_t)D * NC, st); });
  add_spec(h, "params/projection/linear/bias", {NC}, [hh, NC](const float* s, cudaStream_t st) { return copy_f32(s, hh->cls_b, NC, st); });
  return cudaSuccess;
}

// Pooler constants.  The pooling query is a learned parameter, so the projected, PerDimScale'd query
// qh[h,:] = (query . Wq[:,h,:] + bq[h,:]) * 1.442695041/sqrt(dh) * softplus(pds)   (layers.py:502-527,:1093)
// is a constant, and scores[s,h] = x[s] . (Wk[:,h,:] qh[h,:]) + const(h): fold to wkq [H, D].
int finalize_pooler(vp_handle* h) {
  const vp_config& c = h->cfg;
  const int D = c.model_dim, H = c.num_heads, ph = h->pool_ph;
  std::vector<double> qh((size_t)H * ph);
  for (int hh = 0; hh < H; ++hh)
    for (int j = 0; j < ph; ++j) {
      double acc = h->h_pool_bq[(size_t)hh * ph + j];
      for (int d = 0; d < D; ++d) acc += (double)h->h_pool_query[d] * h->h_pool_wq[((size_t)d * H + hh) * ph + j];
      const double p = h->h_pool_pds[j];
      const double softplus = p > 30.0 ? p : log1p(exp(p));
      qh[(size_t)hh * ph + j] = acc * (1.442695041 / sqrt((double)ph)) * softplus;
    }
  std::vector<float> wkq((size_t)H * D);
  for (int hh = 0; hh < H; ++hh)
    for (int d = 0; d < D; ++d) {
      double acc = 0.0;
      for (int j = 0; j < ph; ++j) acc += (double)h->h_pool_wk[((size_t)d * H + hh) * ph + j] * qh[(size_t)hh * ph + j];
      wkq[(size_t)hh * D + d] = (float)acc;
    }
  // The scores run on the tensor core (pooling.cu): wkq = hi + lo with both parts in bf16 keeps ~16 mantissa bits of the
  // folded weight, so the scores stay as accurate as an fp32 dot product of the bf16 tokens would be.
  if (2 * H > 32) return h->fail(VP_ERR_UNSUPPORTED, "pooling head supports at most 16 heads");
  std::vector<bf16> hl((size_t)32 * D, __float2bfloat16(0.f));
  for (int hh = 0; hh < H; ++hh)
    for (int d = 0; d < D; ++d) {
      const float w = wkq[(size_t)hh * D + d];
      const bf16 hi = __float2bfloat16(w);
      hl[(size_t)hh * D + d] = hi;
      hl[(size_t)(H + hh) * D + d] = __float2bfloat16(w - __bfloat162float(hi));
    }
  CK(cudaMemcpy(h->pool_wkq, hl.data(), hl.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  if (h->c_pool_wkq != nullptr) {   // check mode: the folded score weights in fp32, [D, H]
    std::vector<float> t((size_t)D * H);
    for (int hh = 0; hh < H; ++hh)
      for (int d = 0; d < D; ++d) t[(size_t)d * H + hh] = wkq[(size_t)hh * D + d];
    CK(cudaMemcpy(h->c_pool_wkq, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  // the host copies of the query / key projections stay for the life of the handle (like f_wqkv / f_w1): vp_finalize
  // re-folds them after any later vp_set_weight, whichever leaf changed
  return VP_OK;
}

// PositionalEmbedding (encoders.py:240-266), float32 arithmetic as the reference.
int prepare_pe(vp_handle* h, int L, cudaStream_t st) {
  if (h->pe_len == L) return VP_OK;
  const int D = h->cfg.model_dim, nts = D / 2;
  std::vector<float> pe((size_t)L * D, 0.f);
  const float inc = (float)(log(10000.0) / fmax((double)((float)nts - 1.0f), 1.0));
  for (int p = 0; p < L; ++p)
    for (int i = 0; i < nts; ++i) {
      const float inv = expf((float)i * -inc);
      const float stime = (float)p * inv;
      pe[(size_t)p * D + i] = sinf(stime);
      pe[(size_t)p * D + nts + i] = cosf(stime);
    }
  ++h->ws_generation;
  if (h->d_pe) { cudaFree(h->d_pe); h->d_pe = nullptr; }
  CK(cudaMalloc(&h->d_pe, pe.size() * sizeof(float)));
  CK(cudaMemcpyAsync(h->d_pe, pe.data(), pe.size() * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));
  h->pe_len = L;
  return VP_OK;
}

// trace labels of one stack: layernorm, qkv, attention, out-proj, ffn1, ffn2
const char* const kTagSpatial[6] = {"spatial.ln", "spatial.qkv", "spatial.attn", "spatial.outproj", "spatial.ffn1", "spatial.ffn2"};
const char* const kTagTemporal[6] = {"temporal.ln", "temporal.qkv", "temporal.attn", "temporal.outproj", "temporal.ffn1", "temporal.ffn2"};
const char* const kTagAux[6] = {"aux.ln", "aux.qkv", "aux.attn", "aux.outproj", "aux.ffn1", "aux.ffn2"};
const char* const kTagText[6] = {"text.ln", "text.qkv", "text.attn", "text.outproj", "text.ffn1", "text.ffn2"};

struct SeqLayout {
  int num_seq, S, group, causal;
  const float* key_pad;    // [num_seq, S] or null
  const float* row_scale;  // [M] or null
  int pad_whole_seq = 0;   // key_pad is constant within a sequence (frame paddings seen by the spatial stack)
};

// One Transformer stack (layers.py:989-1041) over the bf16 residual stream x [M, D] (in place).
// Fused-LN form (default): on entry stats_a holds (sum, sum of squares) of every row of x; each block is
//   QKV GEMM on raw x with LN1 folded in -> attention -> out-proj + residual (emits stats of the new x into stats_b)
//   -> FFN1 on raw x with LN2 folded in (+act) -> FFN2 + residual (emits stats into stats_a);
// on exit stats_a again describes x.  With fuse_ln off, the two LayerNorms run as standalone kernels.
int run_stack(vp_handle* h, const StackWeights& w, bf16* x, int M, const SeqLayout& sl, int act, cudaStream_t st, float* stats_a,
              int slots_a, float* stats_b, const char* const* tag) {
  const int D = w.D, F = w.F, H = w.H;
  bf16* n = static_cast<bf16*>(h->ws_n.p);
  bf16* qkv = static_cast<bf16*>(h->ws_qkv.p);
  bf16* u = static_cast<bf16*>(h->ws_u.p);
  const bool fuse = h->fuse_ln && stats_a != nullptr;
  const int gslots = vp::gemm_stats_slots(D);   // slots the residual GEMMs (N = D) write
  for (int l = 0; l < w.L; ++l) {
    vp::LnArgs ln;
    ln.x = x; ln.ldx = D; ln.y_bf16 = n; ln.M = M; ln.D = D;
    vp::GemmEpilogue e1;
    if (fuse) {
      e1.bias = w.c2_qkv + (size_t)l * 3 * D; e1.ln_stats_in = stats_a; e1.ln_slots = (l == 0) ? slots_a : (w.primer ? 1 : gslots);
      e1.ln_colsum = w.c1_qkv + (size_t)l * 3 * D; e1.ln_dim = D;
      CK(vp::launch_gemm(st, x, D, w.wqkv_ln + (size_t)l * 3 * D * D, D, qkv, 3 * D, M, 3 * D, D, e1)); h->mark(st, tag[1]);
    } else {
      ln.gamma1 = w.ln1_g + (size_t)l * D; ln.beta = w.ln1_b + (size_t)l * D;
      CK(vp::launch_layernorm(st, ln)); h->mark(st, tag[0]);
      e1.bias = w.bqkv + (size_t)l * 3 * D;
      CK(vp::launch_gemm(st, n, D, w.wqkv + (size_t)l * 3 * D * D, D, qkv, 3 * D, M, 3 * D, D, e1)); h->mark(st, tag[1]);
    }
    vp::AttnArgs at;
    at.q = qkv; at.k = qkv + D; at.v = qkv + 2 * D; at.ld = 3 * D; at.out = n; at.ldo = D;
    at.num_seq = sl.num_seq; at.S = sl.S; at.group = sl.group; at.heads = H; at.dh = D / H;
    at.cap = h->cfg.atten_logit_cap; at.key_pad = sl.key_pad; at.causal = sl.causal; at.pad_whole_seq = sl.pad_whole_seq;
    int n_attn = 1;
    at.launched = &n_attn;
    CK(vp::launch_attention(st, at)); h->mark(st, tag[2]);
    h->launches += n_attn - 1;
    vp::GemmEpilogue e2;
    e2.bias = w.bo + (size_t)l * D;
    if (w.primer) {
      // 'primer_hybrid': x += post_layer_norm(attention output) (layers.py:846-855); the projection lands in the (idle)
      // FFN hidden buffer, the LayerNorm kernel adds the residual and emits the row statistics for the folded FFN1
      CK(vp::launch_gemm(st, n, D, w.wo + (size_t)l * D * D, D, u, D, M, D, D, e2)); h->mark(st, tag[3]);
      vp::LnArgs lp;
      lp.x = u; lp.ldx = D; lp.gamma1 = w.lnp1_g + (size_t)l * D; lp.beta = w.lnp1_b + (size_t)l * D; lp.y_bf16 = x;
      lp.resid = x; lp.ldr = D; lp.M = M; lp.D = D; lp.stats_out = fuse ? stats_b : nullptr;
      CK(vp::launch_layernorm(st, lp)); h->mark(st, tag[0]);
    } else {
      e2.resid = x; e2.ldr = D;
      if (fuse) e2.stats_out = stats_b;
      CK(vp::launch_gemm(st, n, D, w.wo + (size_t)l * D * D, D, x, D, M, D, D, e2)); h->mark(st, tag[3]);
    }
    vp::GemmEpilogue e3;
    e3.act = act; e3.row_scale = sl.row_scale;
    if (fuse) {
      e3.bias = w.c2_ffn1 + (size_t)l * F; e3.ln_stats_in = stats_b; e3.ln_slots = w.primer ? 1 : gslots;
      e3.ln_colsum = w.c1_ffn1 + (size_t)l * F; e3.ln_dim = D;
      CK(vp::launch_gemm(st, x, D, w.w1_ln + (size_t)l * F * D, D, u, F, M, F, D, e3)); h->mark(st, tag[4]);
    } else {
      ln.gamma1 = w.ln2_g + (size_t)l * D; ln.beta = w.ln2_b + (size_t)l * D;
      CK(vp::launch_layernorm(st, ln)); h->mark(st, tag[0]);
      e3.bias = w.b1 + (size_t)l * F;
      CK(vp::launch_gemm(st, n, D, w.w1 + (size_t)l * F * D, D, u, F, M, F, D, e3)); h->mark(st, tag[4]);
    }
    vp::GemmEpilogue e4;
    e4.bias = w.b2 + (size_t)l * D; e4.row_scale = sl.row_scale;
    if (w.primer) {   // x += post_layer_norm(FFN output, zeroed on padded tokens) (layers.py:410-424)
      CK(vp::launch_gemm(st, u, F, w.w2 + (size_t)l * D * F, F, n, D, M, D, F, e4)); h->mark(st, tag[5]);
      vp::LnArgs lp;
      lp.x = n; lp.ldx = D; lp.gamma1 = w.lnp2_g + (size_t)l * D; lp.beta = w.lnp2_b + (size_t)l * D; lp.y_bf16 = x;
      lp.resid = x; lp.ldr = D; lp.M = M; lp.D = D; lp.stats_out = fuse ? stats_a : nullptr;
      CK(vp::launch_layernorm(st, lp)); h->mark(st, tag[0]);
    } else {
      e4.resid = x; e4.ldr = D;
      if (fuse) e4.stats_out = stats_a;
      CK(vp::launch_gemm(st, u, F, w.w2 + (size_t)l * D * F, F, x, D, M, D, F, e4)); h->mark(st, tag[5]);
    }
  }
  return VP_OK;
}

int ensure_workspace(vp_handle* h, size_t M, int D, int F) {
  CK(h->ws_x.ensure(M * D * sizeof(bf16)));
  CK(h->ws_n.ensure(M * D * sizeof(bf16)));
  CK(h->ws_qkv.ensure(M * 3 * D * sizeof(bf16)));
  CK(h->ws_u.ensure(M * F * sizeof(bf16)));
  const int sl = vp::gemm_stats_slots(D) > 16 ? vp::gemm_stats_slots(D) : 16;
  h->stats_stride = M * 2 * (size_t)sl + 64;
  CK(h->ws_stats.ensure(2 * h->stats_stride * sizeof(float)));   // two buffers of [M][slots][2]
  return VP_OK;
}

// ===================================================================== fp32 check mode (check_fp32.cu)
// The same forward, as the reference writes it, entirely in float32: x (residual stream), LayerNorm output, q/k/v,
// attention context and the FFN hidden layer are fp32 buffers; every GEMM is an fp32 FFMA GEMM on the fp32 weights.
int ensure_workspace_f32(vp_handle* h, size_t M, int D, int F) {
  CK(h->c_x.ensure(M * D * sizeof(float)));
  CK(h->c_n.ensure(M * D * sizeof(float)));
  CK(h->c_qkv.ensure(M * 3 * D * sizeof(float)));
  CK(h->c_u.ensure(M * F * sizeof(float)));
  return VP_OK;
}

// One 'pre'-LayerNorm Transformer stack (layers.py:797-872, :989-1041) over the fp32 residual stream x [M, D], in place.
int run_stack_f32(vp_handle* h, const StackWeights& w, float* x, int M, const SeqLayout& sl, int act, cudaStream_t st,
                  const char* const* tag) {
  if (w.primer) return h->fail(VP_ERR_UNSUPPORTED, "fp32 check mode supports norm_policy 'pre' only");
  const int D = w.D, F = w.F, H = w.H;
  float* n = static_cast<float*>(h->c_n.p);
  float* qkv = static_cast<float*>(h->c_qkv.p);
  float* u = static_cast<float*>(h->c_u.p);
  const float qscale = 1.0f / sqrtf(static_cast<float>(D / H));   // layers.py:569-584 (a power of two for dh = 64: exact)
  for (int l = 0; l < w.L; ++l) {
    vp::f32::LayerNorm ln;
    ln.x = x; ln.ldx = D; ln.gamma1 = w.ln1_g + (size_t)l * D; ln.beta = w.ln1_b + (size_t)l * D; ln.y = n; ln.ldy = D; ln.M = M; ln.D = D;
    CK(vp::f32::layernorm(st, ln)); h->mark(st, tag[0]);
    for (int i = 0; i < 3; ++i) {   // query / key / value projections (layers.py:486-498), w [D, (N H)] = [K, N]
      vp::f32::Sgemm g;
      g.A = n; g.lda = D; g.W = w.f_wqkv + ((size_t)l * 3 + i) * D * D; g.ldw = D; g.C = qkv + (size_t)i * D; g.ldc = 3 * D;
      g.M = M; g.N = D; g.K = D; g.bias = w.f_bqkv + ((size_t)l * 3 + i) * D; g.alpha = i == 0 ? qscale : 1.0f;
      CK(vp::f32::sgemm(st, g)); h->mark(st, tag[1]);
    }
    vp::f32::Attention at;
    at.q = qkv; at.k = qkv + D; at.v = qkv + 2 * D; at.ld = 3 * D; at.out = n; at.ldo = D;
    at.num_seq = sl.num_seq; at.S = sl.S; at.group = sl.group; at.heads = H; at.dh = D / H;
    at.cap = h->cfg.atten_logit_cap; at.key_pad = sl.key_pad; at.causal = sl.causal;
    CK(vp::f32::attention(st, at)); h->mark(st, tag[2]);
    {   // post projection + residual (layers.py:483-498, :855); post.w [D_out, (N H)] = [N, K]
      vp::f32::Sgemm g;
      g.A = n; g.lda = D; g.W = w.c_wo + (size_t)l * D * D; g.ldw = D; g.w_nk = 1; g.C = x; g.ldc = D; g.M = M; g.N = D; g.K = D;
      g.bias = w.bo + (size_t)l * D; g.resid = x; g.ldr = D;
      CK(vp::f32::sgemm(st, g)); h->mark(st, tag[3]);
    }
    ln.gamma1 = w.ln2_g + (size_t)l * D; ln.beta = w.ln2_b + (size_t)l * D;
    CK(vp::f32::layernorm(st, ln)); h->mark(st, tag[0]);
    {   // ffn_layer1 + activation, zeroed on padded tokens (layers.py:391-398)
      vp::f32::Sgemm g;
      g.A = n; g.lda = D; g.W = w.f_w1 + (size_t)l * D * F; g.ldw = F; g.C = u; g.ldc = F; g.M = M; g.N = F; g.K = D;
      g.bias = w.f_b1 + (size_t)l * F; g.act = act; g.row_scale = sl.row_scale;
      CK(vp::f32::sgemm(st, g)); h->mark(st, tag[4]);
    }
    {   // ffn_layer2, zeroed on padded tokens, + residual (layers.py:402-425)
      vp::f32::Sgemm g;
      g.A = u; g.lda = F; g.W = w.c_w2 + (size_t)l * F * D; g.ldw = D; g.C = x; g.ldc = D; g.M = M; g.N = D; g.K = F;
      g.bias = w.b2 + (size_t)l * D; g.row_scale = sl.row_scale; g.resid = x; g.ldr = D;
      CK(vp::f32::sgemm(st, g)); h->mark(st, tag[5]);
    }
  }
  return VP_OK;
}

// fp32 counterpart of encoder_body: leaves LN_temporal(x) (final_ln_in_place) or the pre-LN stream in c_x.
int encoder_body_f32(vp_handle* h, const void* video, int in_dtype, int B, int T, int H, int W, const float* frame_pad, float* out_f32,
                     bf16* out_bf16, bool final_ln_in_place, float* spatial_f32, cudaStream_t st, size_t* M_out) {
  const vp_config& c = h->cfg;
  const int D = c.model_dim, P = c.patch_size;
  if (B <= 0 || T <= 0) return h->fail(VP_ERR_INVALID, "empty batch (B=%d, T=%d)", B, T);
  if (H != W) return h->fail(VP_ERR_INVALID, "H (%d) must equal W (%d) (encoders.py:435)", H, W);
  if (H % P || W % P) return h->fail(VP_ERR_INVALID, "Image height (%d) and width (%d) should be multiples of patch_size (%d)", H, W, P);
  const int gh = H / P, gw = W / P, N = gh * gw;
  const size_t M = (size_t)B * T * N;
  if (M > 0x7fffffffULL / (size_t)(3 * D > c.mlp_dim ? 3 * D : c.mlp_dim)) return h->fail(VP_ERR_INVALID, "batch too large");
  int rc;
  if ((rc = prepare_pos_tables(h, T, gh, gw, st)) != VP_OK) return rc;
  if ((rc = ensure_workspace_f32(h, M, D, c.mlp_dim)) != VP_OK) return rc;
  CK(h->c_patch.ensure(M * h->k_patch * sizeof(float)));
  float* x = static_cast<float*>(h->c_x.p);
  float* patches = static_cast<float*>(h->c_patch.p);
  const float *pad_tok = nullptr, *keep_tok = nullptr, *pad_tube = nullptr;
  if (frame_pad != nullptr) {
    CK(h->ws_misc.ensure(3 * M * sizeof(float)));
    float* base = static_cast<float*>(h->ws_misc.p);
    CK(vp::launch_pad_expand(st, frame_pad, base, base + M, base + 2 * M, B, T, N)); h->mark(st, "pad_expand");
    pad_tok = base; keep_tok = base + M; pad_tube = base + 2 * M;
  }
  h->mark(st, nullptr, false);
  CK(vp::f32::patchify(st, video, in_dtype == VP_U8, patches, B * T, H, W, P)); h->mark(st, "patchify");
  {   // patch projection + spatial position table (encoders.py:488-514)
    vp::f32::Sgemm g;
    g.A = patches; g.lda = h->k_patch; g.W = h->c_wpatch; g.ldw = D; g.C = x; g.ldc = D; g.M = (int)M; g.N = D; g.K = h->k_patch;
    g.bias = h->b_patch; g.pos_table = h->d_spatial_pos; g.pos_period = N;
    CK(vp::f32::sgemm(st, g)); h->mark(st, "patch_proj");
  }
  SeqLayout sp{B * T, N, 1, 0, pad_tok, keep_tok, 1};
  if ((rc = run_stack_f32(h, h->spatial, x, (int)M, sp, vp::ACT_GELU, st, kTagSpatial)) != VP_OK) return rc;
  {   // spatial_ln (+ temporal position table, encoders.py:528-553), in place
    vp::f32::LayerNorm ln;
    ln.x = x; ln.ldx = D; ln.gamma1 = h->sp_ln_g; ln.beta = h->sp_ln_b; ln.y = x; ln.ldy = D; ln.y2 = spatial_f32;
    ln.add_table = h->d_temporal_pos; ln.add_div = N; ln.add_mod = T; ln.M = (int)M; ln.D = D;
    CK(vp::f32::layernorm(st, ln)); h->mark(st, "spatial_ln");
  }
  SeqLayout tp{B * N, T, N, 0, pad_tube, keep_tok};
  if ((rc = run_stack_f32(h, h->temporal, x, (int)M, tp, vp::ACT_GELU, st, kTagTemporal)) != VP_OK) return rc;
  {   // temporal_ln (encoders.py:567-569)
    vp::f32::LayerNorm ln;
    ln.x = x; ln.ldx = D; ln.gamma1 = h->tp_ln_g; ln.beta = h->tp_ln_b; ln.M = (int)M; ln.D = D;
    float* tmp = static_cast<float*>(h->c_n.p);
    if (final_ln_in_place) { ln.y = x; ln.ldy = D; ln.y2 = out_f32; }
    else if (out_f32 != nullptr) { ln.y = out_f32; ln.ldy = D; }
    else { ln.y = tmp; ln.ldy = D; }
    CK(vp::f32::layernorm(st, ln)); h->mark(st, "temporal_ln");
    if (!final_ln_in_place && out_f32 == nullptr && out_bf16 != nullptr) { CK(vp::f32::cast_to_bf16(st, tmp, out_bf16, M * D)); h->mark(st, "cast_bf16"); }
  }
  if (M_out) *M_out = M;
  return VP_OK;
}

// Pooling head in fp32 (layers.py:1044-1136) on x [num_seq, S, D] -> out [num_seq, D] (+ optional l2 normalisation).
int pool_f32(vp_handle* h, const float* x, int num_seq, int S, int normalize, float* out, cudaStream_t st) {
  const vp_config& c = h->cfg;
  const int D = c.model_dim, H = c.num_heads, ph = h->pool_ph, HP = H * ph;
  const size_t n = (size_t)num_seq;
  const size_t o_scores = 0, o_xbar = o_scores + n * S * H, o_ctx = o_xbar + n * H * D, o_y = o_ctx + n * HP, o_ln = o_y + n * D, o_end = o_ln + n * D;
  CK(h->c_pool.ensure(o_end * sizeof(float)));
  float* base = static_cast<float*>(h->c_pool.p);
  float *scores = base + o_scores, *xbar = base + o_xbar, *ctx = base + o_ctx, *y = base + o_y, *yln = base + o_ln;
  {   // scores[s, h] = x[s] . wkq[:, h]  (key bias and the constant part cancel in the softmax)
    vp::f32::Sgemm g;
    g.A = x; g.lda = D; g.W = h->c_pool_wkq; g.ldw = H; g.C = scores; g.ldc = H; g.M = num_seq * S; g.N = H; g.K = D;
    CK(vp::f32::sgemm(st, g)); h->mark(st, "pooler.scores");
  }
  CK(vp::f32::pool(st, x, scores, xbar, num_seq, S, D, H)); h->mark(st, "pooler.accum");
  for (int hh = 0; hh < H; ++hh) {   // value projection of the pooled token, per head: [num_seq, D] x [D, ph]
    vp::f32::Sgemm g;
    g.A = xbar + (size_t)hh * D; g.lda = H * D; g.W = h->c_pool_wv + (size_t)hh * ph; g.ldw = HP; g.C = ctx + (size_t)hh * ph; g.ldc = HP;
    g.M = num_seq; g.N = ph; g.K = D; g.bias = h->pool_bv + (size_t)hh * ph;
    CK(vp::f32::sgemm(st, g)); h->mark(st, "pooler.value");
  }
  {   // post projection: post.w [D, (H ph)] = [N, K]
    vp::f32::Sgemm g;
    g.A = ctx; g.lda = HP; g.W = h->c_pool_wpost; g.ldw = HP; g.w_nk = 1; g.C = y; g.ldc = D; g.M = num_seq; g.N = D; g.K = HP; g.bias = h->pool_bpost;
    CK(vp::f32::sgemm(st, g)); h->mark(st, "pooler.post");
  }
  vp::f32::LayerNorm ln;
  ln.x = y; ln.ldx = D; ln.gamma1 = h->pool_ln_g; ln.beta = h->pool_ln_b; ln.y = normalize ? yln : out; ln.ldy = D; ln.M = num_seq; ln.D = D;
  CK(vp::f32::layernorm(st, ln)); h->mark(st, "pooler.ln");
  if (normalize) { CK(vp::launch_l2norm(st, yln, out, num_seq, D)); h->mark(st, "pooler.l2norm"); }
  return VP_OK;
}

// Runs `body(stream)` (a fixed sequence of launches determined by `key`) on `st`, through a cached CUDA graph when there is
// one.  First call with a key: eager (it sizes the workspace, grants shared memory, uploads tables).  Second call: captured
// on the handle's private capture stream (the caller's stream may be the legacy default stream, which cannot be captured)
// and launched into `st`.  Later calls: one cudaGraphLaunch.  Anything that moves a buffer a graph points into bumps
// ws_generation and the stale graphs are dropped.  Tracing (one event after every launch) and nested calls run eagerly.
template <typename Body>
int run_graphed(vp_handle* h, cudaStream_t st, const std::vector<uint64_t>& key, Body&& body) {
  if (!h->use_graphs || h->trace_on || h->capturing) return body(st);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return body(st); }   // the caller is capturing
  for (size_t i = 0; i < h->graphs.size();) {   // drop stale graphs
    if (h->graphs[i].gen != h->ws_generation) { cudaGraphExecDestroy(h->graphs[i].exec); h->graphs.erase(h->graphs.begin() + i); }
    else ++i;
  }
  for (GraphEntry& g : h->graphs)
    if (g.key == key) {
      g.stamp = ++h->graph_clock;
      CK(cudaGraphLaunch(g.exec, st));
      h->launches += g.launches;
      return VP_OK;
    }
  if (std::find(h->graph_bad.begin(), h->graph_bad.end(), key) != h->graph_bad.end()) return body(st);
  if (std::find(h->graph_seen.begin(), h->graph_seen.end(), key) == h->graph_seen.end()) {
    if (h->graph_seen.size() >= 64) h->graph_seen.erase(h->graph_seen.begin());
    h->graph_seen.push_back(key);
    return body(st);
  }
  if (h->s_cap == nullptr) CK(cudaStreamCreateWithFlags(&h->s_cap, cudaStreamNonBlocking));
  const uint64_t gen0 = h->ws_generation;
  const int64_t l0 = h->launches;
  if (cudaStreamBeginCapture(h->s_cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return body(st); }
  h->capturing = true;
  const int rc = body(h->s_cap);
  h->capturing = false;
  cudaGraph_t graph = nullptr;
  const cudaError_t ec = cudaStreamEndCapture(h->s_cap, &graph);
  const int64_t captured = h->launches - l0;
  h->launches = l0;
  cudaGraphExec_t exec = nullptr;
  if (rc != VP_OK || ec != cudaSuccess || graph == nullptr || gen0 != h->ws_generation ||
      cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    if (rc != VP_OK) return rc;
    h->graph_bad.push_back(key);
    return body(st);
  }
  cudaGraphDestroy(graph);
  if (h->graphs.size() >= 32) {   // least recently used goes
    size_t lru = 0;
    for (size_t i = 1; i < h->graphs.size(); ++i) if (h->graphs[i].stamp < h->graphs[lru].stamp) lru = i;
    cudaGraphExecDestroy(h->graphs[lru].exec);
    h->graphs.erase(h->graphs.begin() + lru);
  }
  GraphEntry g;
  g.key = key; g.exec = exec; g.launches = captured; g.gen = gen0; g.stamp = ++h->graph_clock;
  h->graphs.push_back(g);
  CK(cudaGraphLaunch(exec, st));
  h->launches += captured;
  return VP_OK;
}

static inline uint64_t key_ptr(const void* p) { return static_cast<uint64_t>(reinterpret_cast<uintptr_t>(p)); }

// Selects the handle's device for the duration of one C-ABI call and restores the caller's current device afterwards: an
// entry point must not change the calling thread's device as a side effect (a destructor running vp_destroy for a model
// on cuda:0 would otherwise silently move the caller off cuda:1).
struct DeviceScope {
  int prev = -1;
  bool switched = false;
  explicit DeviceScope(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (dev >= 0 && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceScope() {
    if (switched && prev >= 0) cudaSetDevice(prev);
  }
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
};

int check_ready(vp_handle* h) {   // callers hold a DeviceScope on h->device
  if (h == nullptr) return VP_ERR_INVALID;
  if (!h->finalized) return h->fail(VP_ERR_INCOMPLETE, "vp_finalize has not been called");
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev != h->device) return h->fail(VP_ERR_CUDA, "cannot select device %d", h->device);
  return VP_OK;
}

// Shared body of the encoder forward.  Leaves the final (pre-temporal_ln) residual stream in ws_x and
// writes LN outputs where requested.  Returns the token count through *M_out.
int encoder_body_eager(vp_handle* h, const void* video, int in_dtype, int B, int T, int H, int W, const float* frame_pad, float* out_f32,
                       bf16* out_bf16, bool final_ln_in_place, float* spatial_f32, cudaStream_t st, size_t* M_out) {
  if (h->check_fp32) return encoder_body_f32(h, video, in_dtype, B, T, H, W, frame_pad, out_f32, out_bf16, final_ln_in_place, spatial_f32, st, M_out);
  const vp_config& c = h->cfg;
  const int D = c.model_dim, P = c.patch_size;
  if (B <= 0 || T <= 0) return h->fail(VP_ERR_INVALID, "empty batch (B=%d, T=%d)", B, T);
  if (H != W) return h->fail(VP_ERR_INVALID, "H (%d) must equal W (%d) (encoders.py:435)", H, W);
  if (H % P || W % P) return h->fail(VP_ERR_INVALID, "Image height (%d) and width (%d) should be multiples of patch_size (%d)", H, W, P);
  const int gh = H / P, gw = W / P, N = gh * gw;
  const size_t M = (size_t)B * T * N;
  if (M > 0x7fffffffULL / (size_t)(3 * D > c.mlp_dim ? 3 * D : c.mlp_dim)) return h->fail(VP_ERR_INVALID, "batch too large");
  int rc;
  if ((rc = prepare_pos_tables(h, T, gh, gw, st)) != VP_OK) return rc;
  if ((rc = ensure_workspace(h, M, D, c.mlp_dim)) != VP_OK) return rc;
  CK(h->ws_patch.ensure(M * h->k_patch_pad * sizeof(bf16), /*zero=*/true, st));   // the K padding columns stay zero
  bf16* x = static_cast<bf16*>(h->ws_x.p);
  bf16* patches = static_cast<bf16*>(h->ws_patch.p);

  const float *pad_tok = nullptr, *keep_tok = nullptr, *pad_tube = nullptr;
  if (frame_pad != nullptr) {
    CK(h->ws_misc.ensure(3 * M * sizeof(float)));
    float* base = static_cast<float*>(h->ws_misc.p);
    CK(vp::launch_pad_expand(st, frame_pad, base, base + M, base + 2 * M, B, T, N)); h->mark(st, "pad_expand");
    pad_tok = base; keep_tok = base + M; pad_tube = base + 2 * M;
  }

  h->mark(st, nullptr, false);   // trace: start of this forward
  // patchify + cast (encoders.py:436-439), patch projection + spatial pos-emb (:488-514)
  if (in_dtype == VP_U8) CK(vp::launch_patchify_u8(st, static_cast<const uint8_t*>(video), patches, h->k_patch_pad, B * T, H, W, P));
  else CK(vp::launch_patchify(st, static_cast<const float*>(video), patches, h->k_patch_pad, B * T, H, W, P));
  h->mark(st, "patchify");
  float* stats_a = h->fuse_ln ? static_cast<float*>(h->ws_stats.p) : nullptr;
  float* stats_b = h->fuse_ln ? stats_a + h->stats_stride : nullptr;
  vp::GemmEpilogue ep;
  ep.bias = h->b_patch;
  if (N % 32 == 0) {   // tiles of 32 rows never straddle a frame: the position rows of a tile are one TMA box of the bf16 table
    ep.resid = h->d_spatial_pos_bf16; ep.ldr = D; ep.resid_period = N;
  } else {
    ep.pos_table = h->d_spatial_pos; ep.pos_period = N;
  }
  ep.stats_out = stats_a;
  CK(vp::launch_gemm(st, patches, h->k_patch_pad, h->w_patch, h->k_patch_pad, x, D, (int)M, D, h->k_patch_pad, ep)); h->mark(st, "patch_proj");

  // spatial stack: sequences = frames (N contiguous tokens)
  SeqLayout sp{B * T, N, 1, 0, pad_tok, keep_tok, /*pad_whole_seq=*/1};   // a frame is padded as a whole (encoders.py:440-447)
  if ((rc = run_stack(h, h->spatial, x, (int)M, sp, vp::ACT_GELU, st, stats_a, vp::gemm_stats_slots(D), stats_b, kTagSpatial)) != VP_OK) return rc;

  // spatial_ln (+ temporal pos-emb add, encoders.py:528-553); in place on the residual stream
  vp::LnArgs ln;
  ln.x = x; ln.ldx = D; ln.gamma1 = h->sp_ln_g; ln.beta = h->sp_ln_b; ln.y_bf16 = x; ln.y_f32 = spatial_f32;
  ln.add_table = h->d_temporal_pos; ln.add_div = N; ln.add_mod = T; ln.M = (int)M; ln.D = D;
  ln.stats_out = stats_a;   // statistics of LN(x) + Et: the input rows of temporal block 0
  CK(vp::launch_layernorm(st, ln)); h->mark(st, "spatial_ln");

  // temporal stack: sequences = tubes (T tokens, N rows apart)
  SeqLayout tp{B * N, T, N, 0, pad_tube, keep_tok};
  if ((rc = run_stack(h, h->temporal, x, (int)M, tp, vp::ACT_GELU, st, stats_a, 1, stats_b, kTagTemporal)) != VP_OK) return rc;

  // temporal_ln (:567-569); '(bn)td->b(tn)d' (:570-572) is the identity in this layout
  vp::LnArgs lo;
  lo.x = x; lo.ldx = D; lo.gamma1 = h->tp_ln_g; lo.beta = h->tp_ln_b; lo.y_bf16 = final_ln_in_place ? x : out_bf16; lo.y_f32 = out_f32;
  lo.M = (int)M; lo.D = D;
  if (final_ln_in_place) lo.stats_out = stats_a;   // the auxiliary encoder continues on LN(x)
  CK(vp::launch_layernorm(st, lo)); h->mark(st, "temporal_ln");
  if (M_out) *M_out = M;
  return VP_OK;
}

int encoder_body(vp_handle* h, const void* video, int in_dtype, int B, int T, int H, int W, const float* frame_pad, float* out_f32,
                 bf16* out_bf16, bool final_ln_in_place, float* spatial_f32, cudaStream_t st, size_t* M_out) {
  const int P = h->cfg.patch_size;
  if (B <= 0 || T <= 0 || P <= 0 || H <= 0 || W <= 0 || H % P || W % P || H != W)   // the eager body reports the precise reason
    return encoder_body_eager(h, video, in_dtype, B, T, H, W, frame_pad, out_f32, out_bf16, final_ln_in_place, spatial_f32, st, M_out);
  if (M_out) *M_out = (size_t)B * T * (H / P) * (W / P);
  const std::vector<uint64_t> key = {1, (uint64_t)in_dtype, (uint64_t)B, (uint64_t)T, (uint64_t)H, (uint64_t)W, key_ptr(video), key_ptr(frame_pad),
                                     key_ptr(out_f32), key_ptr(out_bf16), (uint64_t)final_ln_in_place, key_ptr(spatial_f32), (uint64_t)h->fuse_ln};
  return run_graphed(h, st, key, [&](cudaStream_t s) {
    return encoder_body_eager(h, video, in_dtype, B, T, H, W, frame_pad, out_f32, out_bf16, final_ln_in_place, spatial_f32, s, nullptr);
  });
}

}  // namespace

// ===================================================================== C ABI
extern "C" {

int vp_create(const vp_config* cfg, vp_handle** out) { return vp_create_on_device(cfg, -1, out); }

int vp_create_on_device(const vp_config* cfg, int device, vp_handle** out) {
  // VP_CHECK_FP32=1 turns every handle of the process into a check-mode handle (the switch the parity tests use)
  const char* ev = getenv("VP_CHECK_FP32");
  return vp_create_ex(cfg, device, (ev && atoi(ev) != 0) ? VP_FLAG_CHECK_FP32 : 0, out);
}

int vp_handle_flags(const vp_handle* h) { return h ? (h->check_fp32 ? VP_FLAG_CHECK_FP32 : 0) : 0; }

int vp_handle_device(const vp_handle* h) { return h ? h->device : -1; }

int vp_create_ex(const vp_config* cfg, int device, unsigned flags, vp_handle** out) {
  if (out == nullptr || cfg == nullptr) { g_create_error = "null argument"; return VP_ERR_INVALID; }
  if (flags & ~static_cast<unsigned>(VP_FLAG_CHECK_FP32)) { g_create_error = "unknown flag"; return VP_ERR_INVALID; }
  if ((flags & VP_FLAG_CHECK_FP32) && cfg->text_norm_policy != 0) { g_create_error = "fp32 check mode supports norm_policy 'pre' only"; return VP_ERR_UNSUPPORTED; }
  *out = nullptr;
  if (cfg->model_dim <= 0 || cfg->num_heads <= 0 || cfg->model_dim % cfg->num_heads || cfg->patch_size <= 0 ||
      (cfg->patch_size % 2) || cfg->mlp_dim <= 0 || (cfg->model_dim % 8) || (cfg->mlp_dim % 8)) {
    g_create_error = "invalid config (model_dim/num_heads/patch_size/mlp_dim)";
    return VP_ERR_INVALID;
  }
  if (cfg->kind != VP_KIND_ENCODER && cfg->kind != VP_KIND_CLIP && cfg->kind != VP_KIND_CLASSIFIER) { g_create_error = "unknown model kind"; return VP_ERR_INVALID; }
  if (cfg->text_norm_policy != 0 && cfg->text_norm_policy != 1) { g_create_error = "text_norm_policy must be 0 ('pre') or 1 ('primer_hybrid')"; return VP_ERR_UNSUPPORTED; }
  if (cfg->kind == VP_KIND_CLASSIFIER && cfg->num_classes <= 0) { g_create_error = "num_classes must be positive for a classifier"; return VP_ERR_INVALID; }
  const int dh = cfg->model_dim / cfg->num_heads;
  if (dh % 8 || dh > 128) { g_create_error = "dim_per_head must be a multiple of 8, at most 128"; return VP_ERR_UNSUPPORTED; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
    return VP_ERR_CUDA;
  }
  if (device >= ndev) { g_create_error = "device ordinal " + std::to_string(device) + " out of range (" + std::to_string(ndev) + " devices)"; return VP_ERR_INVALID; }
  if (device < 0 && (e = cudaGetDevice(&device)) != cudaSuccess) {
    g_create_error = std::string("cannot query the current device: ") + cudaGetErrorString(e);
    return VP_ERR_CUDA;
  }
  DeviceScope device_scope(device);   // cudaSetDevice also creates the primary context if this is the device's first use
  vp_handle* h = new vp_handle();
  h->cfg = *cfg;
  h->check_fp32 = (flags & VP_FLAG_CHECK_FP32) != 0;
  if (const char* ev = getenv("VP_HOST_CHUNK_CLIPS")) h->host_chunk_clips = atoi(ev);   // tuning knob of the host pipeline
  if (const char* ev = getenv("VP_FUSE_LN")) h->fuse_ln = atoi(ev) != 0;
  if (const char* ev = getenv("VP_GRAPHS")) h->use_graphs = atoi(ev) != 0;
  {
    DevBuf* bufs[] = {&h->ws_x, &h->ws_n, &h->ws_qkv, &h->ws_u, &h->ws_patch, &h->ws_misc, &h->ws_io_in, &h->ws_io_out, &h->ws_pool, &h->ws_stats,
                      &h->c_x, &h->c_n, &h->c_qkv, &h->c_u, &h->c_patch, &h->c_pool};
    for (DevBuf* b : bufs) b->gen = &h->ws_generation;
  }
  if (cudaGetDevice(&h->device) != cudaSuccess || h->device != device) {
    g_create_error = "cannot select device " + std::to_string(device);
    delete h;
    return VP_ERR_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, h->device);
  if (prop.major != 10) {
    g_create_error = "device is not sm_100 (B200): compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor);
    delete h;
    return VP_ERR_UNSUPPORTED;
  }
  e = add_encoder(h, cfg->kind == VP_KIND_CLIP ? "params/vision_encoder" : cfg->kind == VP_KIND_CLASSIFIER ? "params/encoder" : "params");
  if (e == cudaSuccess && cfg->kind == VP_KIND_CLIP) e = add_clip_extras(h);
  if (e == cudaSuccess && cfg->kind == VP_KIND_CLASSIFIER) e = add_classifier_extras(h);
  if (e != cudaSuccess) {
    g_create_error = std::string("allocation failed: ") + cudaGetErrorString(e);
    vp_destroy(h);
    return VP_ERR_CUDA;
  }
  *out = h;
  return VP_OK;
}

void vp_destroy(vp_handle* h) {
  if (h == nullptr) return;
  DeviceScope device_scope(h->device);
  for (void* p : h->owned) cudaFree(p);
  for (auto& t : h->trace) cudaEventDestroy(t.second);
  for (GraphEntry& g : h->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (h->s_cap) cudaStreamDestroy(h->s_cap);

  if (h->d_spatial_pos) cudaFree(h->d_spatial_pos);
  if (h->d_spatial_pos_bf16) cudaFree(h->d_spatial_pos_bf16);
  if (h->d_temporal_pos) cudaFree(h->d_temporal_pos);
  if (h->d_pe) cudaFree(h->d_pe);
  if (h->pipe_init) {
    cudaStreamSynchronize(h->s_in); cudaStreamSynchronize(h->s_out);
    cudaStreamDestroy(h->s_in); cudaStreamDestroy(h->s_out);
    for (int i = 0; i < 2; ++i) { cudaEventDestroy(h->ev_in[i]); cudaEventDestroy(h->ev_comp[i]); cudaEventDestroy(h->ev_out[i]); }
    for (int i = 0; i < vp_handle::kTickets; ++i) cudaEventDestroy(h->ev_done[i]);
  }
  DevBuf* bufs[] = {&h->staging, &h->ws_x, &h->ws_n, &h->ws_qkv, &h->ws_u, &h->ws_patch, &h->ws_misc, &h->ws_io_in, &h->ws_io_out, &h->ws_pool, &h->ws_stats,
                    &h->c_x, &h->c_n, &h->c_qkv, &h->c_u, &h->c_patch, &h->c_pool};
  for (DevBuf* b : bufs) b->release();
  delete h;
}

const char* vp_last_error(const vp_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int vp_num_weights(const vp_handle* h) { return h ? (int)h->specs.size() : 0; }
const char* vp_weight_key(const vp_handle* h, int i) { return (h && i >= 0 && i < (int)h->specs.size()) ? h->specs[i].key.c_str() : nullptr; }
int vp_weight_ndim(const vp_handle* h, int i) { return (h && i >= 0 && i < (int)h->specs.size()) ? (int)h->specs[i].shape.size() : -1; }
int64_t vp_weight_dim(const vp_handle* h, int i, int axis) {
  if (!h || i < 0 || i >= (int)h->specs.size() || axis < 0 || axis >= (int)h->specs[i].shape.size()) return -1;
  return h->specs[i].shape[axis];
}

int vp_set_weight(vp_handle* h, const char* key, const void* data, const int64_t* shape, int ndim) {
  if (h == nullptr || key == nullptr || data == nullptr || (ndim > 0 && shape == nullptr)) return VP_ERR_INVALID;
  DeviceScope device_scope(h->device);
  auto it = h->spec_index.find(key);
  if (it == h->spec_index.end()) return h->fail(VP_ERR_KEY, "unknown parameter key '%s'", key);
  ParamSpec& sp = h->specs[it->second];
  bool same = (int)sp.shape.size() == ndim;
  size_t count = 1;
  for (int i = 0; same && i < ndim; ++i) same = sp.shape[i] == shape[i];
  if (!same) return h->fail(VP_ERR_KEY, "parameter '%s' has the wrong shape", key);
  for (int64_t d : sp.shape) count *= (size_t)d;
  CK(h->staging.ensure(count * sizeof(float)));
  CK(cudaMemcpy(h->staging.p, data, count * sizeof(float), cudaMemcpyDefault));
  CK(sp.repack(static_cast<const float*>(h->staging.p), 0));
  CK(cudaStreamSynchronize(0));
  sp.set = true;
  h->finalized = false;
  return VP_OK;
}

static int finalize_stack(vp_handle* h, StackWeights* w) {
  const int L = w->L, D = w->D, F = w->F;
  const float qscale = 1.0f / sqrtf(static_cast<float>(D / w->H));
  for (int l = 0; l < L; ++l) {
    for (int i = 0; i < 3; ++i) {
      CK(vp::launch_fold_ln_weight(0, w->f_wqkv + ((size_t)l * 3 + i) * D * D, w->ln1_g + (size_t)l * D, w->ln1_b + (size_t)l * D,
                                   w->f_bqkv + ((size_t)l * 3 + i) * D, w->wqkv_ln + (size_t)l * 3 * D * D + (size_t)i * D * D,
                                   w->c1_qkv + (size_t)l * 3 * D + (size_t)i * D, w->c2_qkv + (size_t)l * 3 * D + (size_t)i * D, D, D, D,
                                   i == 0 ? qscale : 1.0f));
    }
    CK(vp::launch_fold_ln_weight(0, w->f_w1 + (size_t)l * D * F, w->ln2_g + (size_t)l * D, w->ln2_b + (size_t)l * D, w->f_b1 + (size_t)l * F,
                                 w->w1_ln + (size_t)l * F * D, w->c1_ffn1 + (size_t)l * F, w->c2_ffn1 + (size_t)l * F, D, F, D, 1.0f));
  }
  CK(cudaStreamSynchronize(0));
  return VP_OK;
}

int vp_finalize(vp_handle* h) {
  if (h == nullptr) return VP_ERR_INVALID;
  DeviceScope device_scope(h->device);
  for (const ParamSpec& s : h->specs)
    if (!s.set) return h->fail(VP_ERR_INCOMPLETE, "parameter '%s' was never set", s.key.c_str());
  for (StackWeights* w : h->stacks) {
    int rc = finalize_stack(h, w);
    if (rc != VP_OK) return rc;
  }
  if (h->cfg.kind == VP_KIND_CLIP || h->cfg.kind == VP_KIND_CLASSIFIER) {
    int rc = finalize_pooler(h);
    if (rc != VP_OK) return rc;
  }
  h->staging.release();
  h->finalized = true;
  return VP_OK;
}

static int encoder_forward_dev(vp_handle* h, const void* video, int in_dtype, int B, int T, int H, int W, const float* frame_paddings,
                               void* out_features, void* spatial_features, int out_dtype, void* stream) {
  DeviceScope device_scope(h ? h->device : -1);
  int rc = check_ready(h);
  if (rc != VP_OK) return rc;
  if (video == nullptr || out_features == nullptr) return h->fail(VP_ERR_INVALID, "null video / output pointer");
  if (out_dtype != VP_F32 && out_dtype != VP_BF16) return h->fail(VP_ERR_INVALID, "out_dtype must be VP_F32 or VP_BF16");
  if (spatial_features != nullptr && out_dtype != VP_F32) return h->fail(VP_ERR_UNSUPPORTED, "spatial_features requires VP_F32 outputs");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Clips are independent (encoders.py:434-436 only folds B into the leading axis), so a batch larger than one pass can
  // hold (32-bit element indices, bounded workspace: at most 2^18 tokens = 64 clips of 16 x 256) runs as consecutive
  // passes on the same stream; the results are bitwise those of a single pass.
  const int P = h->cfg.patch_size;
  if (B <= 0 || T <= 0 || P <= 0 || H % P || W % P || H != W)   // encoder_body reports the precise reason
    return encoder_body(h, video, in_dtype, B, T, H, W, frame_paddings, nullptr, nullptr, false, nullptr, st, nullptr);
  const size_t tokens_per_clip = (size_t)T * (H / P) * (W / P);
  const size_t D = h->cfg.model_dim;
  const int max_clips = (int)std::max<size_t>(1, ((size_t)1 << 18) / tokens_per_clip);
  const size_t in_elem = in_dtype == VP_U8 ? 1 : sizeof(float), out_elem = out_dtype == VP_F32 ? sizeof(float) : sizeof(bf16);
  for (int b0 = 0; b0 < B; b0 += max_clips) {
    const int bc = std::min(max_clips, B - b0);
    const char* vin = static_cast<const char*>(video) + (size_t)b0 * T * H * W * 3 * in_elem;
    char* o = static_cast<char*>(out_features) + (size_t)b0 * tokens_per_clip * D * out_elem;
    float* sp = spatial_features ? static_cast<float*>(spatial_features) + (size_t)b0 * tokens_per_clip * D : nullptr;
    rc = encoder_body(h, vin, in_dtype, bc, T, H, W, frame_paddings ? frame_paddings + (size_t)b0 * T : nullptr,
                      out_dtype == VP_F32 ? reinterpret_cast<float*>(o) : nullptr, out_dtype == VP_BF16 ? reinterpret_cast<bf16*>(o) : nullptr,
                      false, sp, st, nullptr);
    if (rc != VP_OK) return rc;
  }
  return VP_OK;
}

int vp_encoder_forward(vp_handle* h, const float* video, int B, int T, int H, int W, const float* frame_paddings,
                       void* out_features, void* spatial_features, int out_dtype, void* stream) {
  return encoder_forward_dev(h, video, VP_F32, B, T, H, W, frame_paddings, out_features, spatial_features, out_dtype, stream);
}

int vp_encoder_forward_u8(vp_handle* h, const uint8_t* video, int B, int T, int H, int W, const float* frame_paddings,
                          void* out_features, void* spatial_features, int out_dtype, void* stream) {
  return encoder_forward_dev(h, video, VP_U8, B, T, H, W, frame_paddings, out_features, spatial_features, out_dtype, stream);
}

// ---------------------------------------------------------------------------------------------- host-buffer pipeline
// Host-buffer entry points are software-pipelined over clip chunks: H2D of chunk i+1 (copy-in stream), forward of chunk i
// (caller's stream) and D2H of chunk i-1 (copy-out stream) overlap; device staging is double buffered and ordered with
// events.  The slot sequence runs over ALL calls of the handle, and the *_async entry points return once everything is
// enqueued, so the H2D of call k+1 also overlaps the forward of call k and the D2H of call k overlaps the forward of
// call k+1 (vp_wait(ticket) blocks until a call's results are in the caller's buffers).  Pinned host buffers give true
// DMA overlap; pageable ones still work (the runtime stages them and the enqueue blocks).
static int pipe_setup(vp_handle* h) {
  if (h->pipe_init) return VP_OK;
  CK(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CK(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < vp_handle::kTickets; ++i) CK(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
  h->pipe_init = true;
  return VP_OK;
}

// Chunk schedule.  BLOCKING call (`ramp`): only the first chunk's H2D and the last chunk's D2H are exposed (everything
// else overlaps the forward of a neighbouring chunk), so both ends are small (2 clips); every chunk costs fixed launch
// overhead and small chunks run the GEMMs less efficiently, so the middle runs 8-clip chunks with 6-clip ramps:
// 32 clips -> 2 6 8 8 6 2.  ASYNCHRONOUS call: the ends overlap the neighbouring CALLS' forwards, so there is nothing to
// ramp: equal chunks of at most 8 clips (32 -> 8 8 8 8, 8 -> 8, 12 -> 6 6).
static std::vector<int> chunk_schedule(const vp_handle* h, int B, bool ramp) {
  std::vector<int> sizes;
  if (h->host_chunk_clips > 0) {
    for (int c0 = 0; c0 < B; c0 += h->host_chunk_clips) sizes.push_back(B - c0 < h->host_chunk_clips ? B - c0 : h->host_chunk_clips);
  } else if (!ramp && B > 4) {
    const int n = (B + 7) / 8;
    for (int i = 0; i < n; ++i) sizes.push_back(B / n + (i < B % n ? 1 : 0));
  } else if (B <= 4) {
    for (int c0 = 0; c0 < B; c0 += 2) sizes.push_back(B - c0 < 2 ? B - c0 : 2);
  } else {
    const int R = B - 4;   // between the two 2-clip end chunks
    sizes.push_back(2);
    if (R <= 8) {
      sizes.push_back(R);
    } else {
      const int n8 = R >= 12 ? (R - 12) / 8 : 0;
      const int rem = R - 8 * n8;
      sizes.push_back(rem / 2);
      for (int i = 0; i < n8; ++i) sizes.push_back(8);
      sizes.push_back(rem - rem / 2);
    }
    sizes.push_back(2);
  }
  return sizes;
}

static int clip_video_forward_dev(vp_handle* h, const void* video, int in_dtype, int B, int T, int H, int W, const float* frame_paddings,
                                  int normalize, float* video_emb, float* spatial_features, float* spatiotemporal_features,
                                  float* frame_embeddings, cudaStream_t st);

// mode 0: encoder features (out_dtype VP_F32 / VP_BF16, optional fp32 spatial_features); mode 1: video-text model, pooled
// video embeddings [B, D] fp32 (`normalize`).  Enqueues everything and records the call's completion event; no host wait.
static int host_pipeline(vp_handle* h, int mode, const void* video_v, int in_dtype, int B, int T, int H, int W, const float* frame_paddings,
                         void* out_v, float* spatial_features, int out_dtype, int normalize, cudaStream_t st, uint64_t* ticket,
                         bool ramp) {
  const char* video = static_cast<const char*>(video_v);
  const size_t esz = in_dtype == VP_U8 ? 1 : sizeof(float);
  int rc = check_ready(h);
  if (rc != VP_OK) return rc;
  if (mode == 1 && h->cfg.kind != VP_KIND_CLIP) return h->fail(VP_ERR_INVALID, "handle is not a video-text (CLIP) model");
  if (video == nullptr || out_v == nullptr) return h->fail(VP_ERR_INVALID, "null video / output pointer");
  if (in_dtype != VP_F32 && in_dtype != VP_U8) return h->fail(VP_ERR_INVALID, "in_dtype must be VP_F32 or VP_U8");
  if (out_dtype != VP_F32 && out_dtype != VP_BF16) return h->fail(VP_ERR_INVALID, "out_dtype must be VP_F32 or VP_BF16");
  if (spatial_features != nullptr && out_dtype != VP_F32) return h->fail(VP_ERR_UNSUPPORTED, "spatial_features requires VP_F32 outputs");
  if (B <= 0 || T <= 0 || H <= 0 || W <= 0 || H % h->cfg.patch_size || W % h->cfg.patch_size) return h->fail(VP_ERR_INVALID, "bad clip shape");
  if ((rc = pipe_setup(h)) != VP_OK) return rc;
  const size_t clip_in = (size_t)T * H * W * 3;
  const size_t N = (size_t)(H / h->cfg.patch_size) * (W / h->cfg.patch_size);
  const size_t D = h->cfg.model_dim;
  const size_t clip_out = (size_t)T * N * D;
  const size_t osz = out_dtype == VP_BF16 ? sizeof(bf16) : sizeof(float);
  const std::vector<int> sizes = chunk_schedule(h, B, ramp);
  const int nchunks = (int)sizes.size();
  int chunk = 0;
  for (int c : sizes) chunk = c > chunk ? c : chunk;
  // The two staging slots of a buffer are `stride` bytes apart, and calls of DIFFERENT batch sizes may be in flight together
  // (asynchronous entry points), so the stride must not depend on the call: it only ever grows, and when it does (a larger
  // chunk than any before) everything in flight is drained first, because the slots move.
  const size_t in_need = ((size_t)chunk * clip_in * esz + 255) / 256 * 256;
  const size_t out_need = mode == 0 ? ((size_t)chunk * clip_out * sizeof(float) + 255) / 256 * 256 : 0;
  if (in_need > h->pipe_in_stride || out_need > h->pipe_out_stride || mode != h->pipe_last_mode) {
    CK(cudaDeviceSynchronize());
    h->pipe_last_mode = mode;
    h->pipe_in_stride = std::max(h->pipe_in_stride, in_need);
    h->pipe_out_stride = std::max(h->pipe_out_stride, out_need);
  }
  const size_t in_stride = h->pipe_in_stride;
  const size_t out_stride = h->pipe_out_stride;
  const size_t pad_bytes = frame_paddings ? ((size_t)B * T * sizeof(float) + 255) / 256 * 256 : 0;
  CK(h->ws_io_in.ensure(2 * in_stride + 256 + pad_bytes));
  CK(h->ws_io_out.ensure(mode == 0 ? 2 * out_stride * (spatial_features ? 2 : 1) : (size_t)B * D * sizeof(float)));
  char* in_base = static_cast<char*>(h->ws_io_in.p);
  char* out_base = static_cast<char*>(h->ws_io_out.p);
  float* d_pad = nullptr;
  if (frame_paddings) {   // on the compute stream: ordered after the previous call's kernels that read this region
    d_pad = reinterpret_cast<float*>(in_base + 2 * in_stride);
    CK(cudaMemcpyAsync(d_pad, frame_paddings, (size_t)B * T * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  int c0 = 0;
  for (int i = 0; i < nchunks; c0 += sizes[i], ++i) {
    const int b = (int)(h->chunk_seq & 1);
    const int bc = sizes[i];
    void* d_in = in_base + b * in_stride;
    char* d_out = out_base + b * out_stride;
    float* d_sp = spatial_features ? reinterpret_cast<float*>(out_base + (2 + b) * out_stride) : nullptr;
    if (h->comp_rec[b]) CK(cudaStreamWaitEvent(h->s_in, h->ev_comp[b], 0));   // the forward two chunks ago no longer reads this input slot
    CK(cudaMemcpyAsync(d_in, video + (size_t)c0 * clip_in * esz, (size_t)bc * clip_in * esz, cudaMemcpyHostToDevice, h->s_in));
    CK(cudaEventRecord(h->ev_in[b], h->s_in));
    CK(cudaStreamWaitEvent(st, h->ev_in[b], 0));
    if (mode == 0 && h->out_rec[b]) CK(cudaStreamWaitEvent(st, h->ev_out[b], 0));   // the D2H two chunks ago has drained this output slot
    const float* padc = d_pad ? d_pad + (size_t)c0 * T : nullptr;
    if (mode == 0) {
      rc = encoder_body(h, d_in, in_dtype, bc, T, H, W, padc, out_dtype == VP_F32 ? reinterpret_cast<float*>(d_out) : nullptr,
                        out_dtype == VP_BF16 ? reinterpret_cast<bf16*>(d_out) : nullptr, false, d_sp, st, nullptr);
    } else {
      rc = clip_video_forward_dev(h, d_in, in_dtype, bc, T, H, W, padc, normalize, reinterpret_cast<float*>(out_base) + (size_t)c0 * D,
                                  nullptr, nullptr, nullptr, st);
    }
    if (rc != VP_OK) { cudaDeviceSynchronize(); return rc; }
    CK(cudaEventRecord(h->ev_comp[b], st));
    h->comp_rec[b] = true;
    if (mode == 0) {
      CK(cudaStreamWaitEvent(h->s_out, h->ev_comp[b], 0));
      CK(cudaMemcpyAsync(static_cast<char*>(out_v) + (size_t)c0 * clip_out * osz, d_out, (size_t)bc * clip_out * osz, cudaMemcpyDeviceToHost, h->s_out));
      if (spatial_features)
        CK(cudaMemcpyAsync(spatial_features + (size_t)c0 * clip_out, d_sp, (size_t)bc * clip_out * sizeof(float), cudaMemcpyDeviceToHost, h->s_out));
      CK(cudaEventRecord(h->ev_out[b], h->s_out));
      h->out_rec[b] = true;
    }
    h->chunk_seq++;
  }
  if (mode == 1) {   // the pooled embeddings of the whole call, once (B x D floats)
    CK(cudaStreamWaitEvent(h->s_out, h->ev_comp[(h->chunk_seq - 1) & 1], 0));
    CK(cudaMemcpyAsync(out_v, out_base, (size_t)B * D * sizeof(float), cudaMemcpyDeviceToHost, h->s_out));
    // the next call's first forward must not overwrite the embedding buffer before this copy has read it
    CK(cudaEventRecord(h->ev_out[0], h->s_out));
    CK(cudaStreamWaitEvent(st, h->ev_out[0], 0));
  }
  const uint64_t tk = h->next_ticket++;
  CK(cudaEventRecord(h->ev_done[tk % vp_handle::kTickets], h->s_out));
  if (ticket) *ticket = tk;
  return VP_OK;
}

int vp_wait(vp_handle* h, uint64_t ticket) {
  if (h == nullptr) return VP_ERR_INVALID;
  DeviceScope device_scope(h->device);
  if (!h->pipe_init || ticket == 0 || ticket >= h->next_ticket) return h->fail(VP_ERR_INVALID, "unknown ticket");
  if (h->next_ticket - ticket > (uint64_t)vp_handle::kTickets) {   // its event slot has been reused: everything that old is long done
    CK(cudaStreamSynchronize(h->s_out));
    return VP_OK;
  }
  CK(cudaEventSynchronize(h->ev_done[ticket % vp_handle::kTickets]));
  return VP_OK;
}

int vp_encoder_forward_host_async(vp_handle* h, const void* video, int in_dtype, int B, int T, int H, int W, const float* frame_paddings,
                                  void* out_features, float* spatial_features, int out_dtype, void* stream, uint64_t* ticket) {
  DeviceScope device_scope(h ? h->device : -1);
  if (h == nullptr) return VP_ERR_INVALID;
  return host_pipeline(h, 0, video, in_dtype, B, T, H, W, frame_paddings, out_features, spatial_features, out_dtype, 0,
                       static_cast<cudaStream_t>(stream), ticket, /*ramp=*/false);
}

int vp_clip_video_forward_host_async(vp_handle* h, const void* video, int in_dtype, int B, int T, int H, int W, const float* frame_paddings,
                                     int normalize, float* video_emb, void* stream, uint64_t* ticket) {
  DeviceScope device_scope(h ? h->device : -1);
  if (h == nullptr) return VP_ERR_INVALID;
  return host_pipeline(h, 1, video, in_dtype, B, T, H, W, frame_paddings, video_emb, nullptr, VP_F32, normalize,
                       static_cast<cudaStream_t>(stream), ticket, /*ramp=*/false);
}

static int host_sync_call(vp_handle* h, int mode, const void* video, int in_dtype, int B, int T, int H, int W, const float* frame_paddings,
                          void* out, float* spatial_features, int normalize, void* stream) {
  DeviceScope device_scope(h ? h->device : -1);
  if (h == nullptr) return VP_ERR_INVALID;
  uint64_t tk = 0;
  int rc = host_pipeline(h, mode, video, in_dtype, B, T, H, W, frame_paddings, out, spatial_features, VP_F32, normalize,
                         static_cast<cudaStream_t>(stream), &tk, /*ramp=*/true);
  if (rc != VP_OK) return rc;
  CK(cudaEventSynchronize(h->ev_done[tk % vp_handle::kTickets]));
  CK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return VP_OK;
}

int vp_encoder_forward_host(vp_handle* h, const float* video, int B, int T, int H, int W, const float* frame_paddings,
                            float* out_features, float* spatial_features, void* stream) {
  return host_sync_call(h, 0, video, VP_F32, B, T, H, W, frame_paddings, out_features, spatial_features, 0, stream);
}

int vp_encoder_forward_host_u8(vp_handle* h, const uint8_t* video, int B, int T, int H, int W, const float* frame_paddings,
                               float* out_features, float* spatial_features, void* stream) {
  return host_sync_call(h, 0, video, VP_U8, B, T, H, W, frame_paddings, out_features, spatial_features, 0, stream);
}

int vp_clip_video_forward(vp_handle* h, const float* video, int B, int T, int H, int W, const float* frame_paddings,
                          int normalize, float* video_emb, float* spatial_features, float* spatiotemporal_features,
                          float* frame_embeddings, void* stream) {
  DeviceScope device_scope(h ? h->device : -1);
  int rc = check_ready(h);
  if (rc != VP_OK) return rc;
  return clip_video_forward_dev(h, video, VP_F32, B, T, H, W, frame_paddings, normalize, video_emb, spatial_features,
                                spatiotemporal_features, frame_embeddings, static_cast<cudaStream_t>(stream));
}

int vp_clip_video_forward_u8(vp_handle* h, const uint8_t* video, int B, int T, int H, int W, const float* frame_paddings,
                             int normalize, float* video_emb, float* spatial_features, float* spatiotemporal_features,
                             float* frame_embeddings, void* stream) {
  DeviceScope device_scope(h ? h->device : -1);
  int rc = check_ready(h);
  if (rc != VP_OK) return rc;
  return clip_video_forward_dev(h, video, VP_U8, B, T, H, W, frame_paddings, normalize, video_emb, spatial_features,
                                spatiotemporal_features, frame_embeddings, static_cast<cudaStream_t>(stream));
}

static int clip_video_forward_eager(vp_handle* h, const void* video, int in_dtype, int B, int T, int H, int W, const float* frame_paddings,
                                    int normalize, float* video_emb, float* spatial_features, float* spatiotemporal_features,
                                    float* frame_embeddings, cudaStream_t st);

// callers hold a DeviceScope and have passed check_ready
static int clip_video_forward_dev(vp_handle* h, const void* video, int in_dtype, int B, int T, int H, int W, const float* frame_paddings,
                                  int normalize, float* video_emb, float* spatial_features, float* spatiotemporal_features,
                                  float* frame_embeddings, cudaStream_t st) {
  const std::vector<uint64_t> key = {2, (uint64_t)in_dtype, (uint64_t)B, (uint64_t)T, (uint64_t)H, (uint64_t)W, key_ptr(video), key_ptr(frame_paddings),
                                     (uint64_t)normalize, key_ptr(video_emb), key_ptr(spatial_features), key_ptr(spatiotemporal_features),
                                     key_ptr(frame_embeddings), (uint64_t)h->fuse_ln};
  return run_graphed(h, st, key, [&](cudaStream_t s) {
    return clip_video_forward_eager(h, video, in_dtype, B, T, H, W, frame_paddings, normalize, video_emb, spatial_features,
                                    spatiotemporal_features, frame_embeddings, s);
  });
}

static int clip_video_forward_eager(vp_handle* h, const void* video, int in_dtype, int B, int T, int H, int W, const float* frame_paddings,
                                    int normalize, float* video_emb, float* spatial_features, float* spatiotemporal_features,
                                    float* frame_embeddings, cudaStream_t st) {
  int rc;
  if (h->cfg.kind != VP_KIND_CLIP) return h->fail(VP_ERR_INVALID, "handle is not a video-text (CLIP) model");
  if (video == nullptr || video_emb == nullptr) return h->fail(VP_ERR_INVALID, "null video / output pointer");
  const vp_config& c = h->cfg;
  const int D = c.model_dim;
  size_t M = 0;
  // vision_encoder (encoders.py:822-841).  Its output (after temporal_ln) becomes the residual
  // stream of the auxiliary encoder, so LN writes bf16 back into ws_x.
  rc = encoder_body(h, video, in_dtype, B, T, H, W, frame_paddings, spatiotemporal_features, nullptr, true, spatial_features, st, &M);
  if (rc != VP_OK) return rc;
  if (h->check_fp32) {   // the same three stages in float32
    float* xf = static_cast<float*>(h->c_x.p);
    const int Nf = (int)(M / ((size_t)B * T));
    if (c.num_auxiliary_layers > 0) {
      SeqLayout ax{B, T * Nf, 1, 0, nullptr, nullptr};
      if ((rc = run_stack_f32(h, h->aux, xf, (int)M, ax, vp::ACT_GELU, st, kTagAux)) != VP_OK) return rc;
    }
    if ((rc = pool_f32(h, xf, B, T * Nf, normalize, video_emb, st)) != VP_OK) return rc;
    if (frame_embeddings && (rc = pool_f32(h, xf, B * T, Nf, normalize, frame_embeddings, st)) != VP_OK) return rc;
    return VP_OK;
  }
  bf16* x = static_cast<bf16*>(h->ws_x.p);
  const int N = (int)(M / ((size_t)B * T));
  if (c.num_auxiliary_layers > 0) {  // auxiliary_encoder: full attention over all T*N tokens of a clip (:846-857)
    SeqLayout ax{B, T * N, 1, 0, nullptr, nullptr};
    float* stats_a = h->fuse_ln ? static_cast<float*>(h->ws_stats.p) : nullptr;
    float* stats_b = h->fuse_ln ? stats_a + h->stats_stride : nullptr;
    if ((rc = run_stack(h, h->aux, x, (int)M, ax, vp::ACT_GELU, st, stats_a, 1, stats_b, kTagAux)) != VP_OK) return rc;
  }
  const int ph = h->pool_ph;
  size_t need = vp::pool_scratch_floats(B, T * N, D, c.num_heads, ph);
  size_t need_f = frame_embeddings ? vp::pool_scratch_floats(B * T, N, D, c.num_heads, ph) : 0;
  CK(h->ws_pool.ensure((need > need_f ? need : need_f) * sizeof(float)));
  CK(vp::launch_pool(st, x, B, T * N, D, c.num_heads, ph, h->pool_wkq, h->pool_wv, h->pool_bv, h->pool_wpost, h->pool_bpost,
                     h->pool_ln_g, h->pool_ln_b, normalize, static_cast<float*>(h->ws_pool.p), video_emb, &h->launches));
  h->mark(st, "pooler", false);
  if (frame_embeddings) {  // same pooler on the per-frame token sets (:874-885)
    CK(vp::launch_pool(st, x, B * T, N, D, c.num_heads, ph, h->pool_wkq, h->pool_wv, h->pool_bv, h->pool_wpost, h->pool_bpost,
                       h->pool_ln_g, h->pool_ln_b, normalize, static_cast<float*>(h->ws_pool.p), frame_embeddings, &h->launches));
    h->mark(st, "pooler.frames", false);
  }
  return VP_OK;
}

int vp_classifier_forward(vp_handle* h, const float* video, int B, int T, int H, int W, const float* frame_paddings, float* logits,
                          float* global_embeddings, float* spatial_features, float* spatiotemporal_features, void* stream) {
  DeviceScope device_scope(h ? h->device : -1);
  int rc = check_ready(h);
  if (rc != VP_OK) return rc;
  if (h->cfg.kind != VP_KIND_CLASSIFIER) return h->fail(VP_ERR_INVALID, "handle is not a video classifier");
  if (video == nullptr || logits == nullptr) return h->fail(VP_ERR_INVALID, "null video / output pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const vp_config& c = h->cfg;
  const int D = c.model_dim;
  size_t M = 0;
  // encoder (encoders.py:616-627); the pooler reads LN(x) in bf16 from the residual buffer
  rc = encoder_body(h, video, VP_F32, B, T, H, W, frame_paddings, spatiotemporal_features, nullptr, true, spatial_features, st, &M);
  if (rc != VP_OK) return rc;
  if (h->check_fp32) {
    const int Nf = (int)(M / ((size_t)B * T));
    CK(h->ws_pool.ensure((size_t)B * D * sizeof(float)));
    float* embf = global_embeddings ? global_embeddings : static_cast<float*>(h->ws_pool.p);
    if ((rc = pool_f32(h, static_cast<const float*>(h->c_x.p), B, T * Nf, 0, embf, st)) != VP_OK) return rc;
    CK(vp::launch_dense_f32(st, embf, h->cls_w, h->cls_b, logits, B, D, c.num_classes)); h->mark(st, "classifier_projection");
    return VP_OK;
  }
  const bf16* x = static_cast<const bf16*>(h->ws_x.p);
  const int N = (int)(M / ((size_t)B * T));
  const int ph = h->pool_ph;
  // atten_pooler (:631-639): paddings=None, LayerNorm inside, no l2 normalisation; then projection (:643-650)
  CK(h->ws_pool.ensure((vp::pool_scratch_floats(B, T * N, D, c.num_heads, ph) + (size_t)B * D) * sizeof(float)));
  float* scratch = static_cast<float*>(h->ws_pool.p);
  float* emb = global_embeddings ? global_embeddings : scratch + vp::pool_scratch_floats(B, T * N, D, c.num_heads, ph);
  CK(vp::launch_pool(st, x, B, T * N, D, c.num_heads, ph, h->pool_wkq, h->pool_wv, h->pool_bv, h->pool_wpost, h->pool_bpost,
                     h->pool_ln_g, h->pool_ln_b, 0, scratch, emb, &h->launches));
  h->mark(st, "pooler", false);
  CK(vp::launch_dense_f32(st, emb, h->cls_w, h->cls_b, logits, B, D, c.num_classes)); h->mark(st, "classifier_projection");
  return VP_OK;
}

static int clip_text_forward_eager(vp_handle* h, const int32_t* ids, const float* paddings, int Q, int L, int normalize, float* text_emb,
                                   cudaStream_t st);

int vp_clip_text_forward(vp_handle* h, const int32_t* ids, const float* paddings, int Q, int L, int normalize, float* text_emb,
                         void* stream) {
  DeviceScope device_scope(h ? h->device : -1);
  int rc = check_ready(h);
  if (rc != VP_OK) return rc;
  if (h->cfg.kind != VP_KIND_CLIP) return h->fail(VP_ERR_INVALID, "handle is not a video-text (CLIP) model");
  if (ids == nullptr || text_emb == nullptr) return h->fail(VP_ERR_INVALID, "null ids / output pointer");
  if (paddings == nullptr) return h->fail(VP_ERR_INVALID, "Text paddings are required. (encoders.py:888)");
  if (Q <= 0 || L <= 0) return h->fail(VP_ERR_INVALID, "empty text batch");
  const std::vector<uint64_t> key = {3, (uint64_t)Q, (uint64_t)L, key_ptr(ids), key_ptr(paddings), (uint64_t)normalize, key_ptr(text_emb), (uint64_t)h->fuse_ln};
  return run_graphed(h, static_cast<cudaStream_t>(stream), key, [&](cudaStream_t s) {
    return clip_text_forward_eager(h, ids, paddings, Q, L, normalize, text_emb, s);
  });
}

static int clip_text_forward_eager(vp_handle* h, const int32_t* ids, const float* paddings, int Q, int L, int normalize, float* text_emb,
                                   cudaStream_t st) {
  int rc;
  const vp_config& c = h->cfg;
  const int D = c.model_dim, S = L + 1;
  const size_t M = (size_t)Q * S;
  if ((rc = prepare_pe(h, L, st)) != VP_OK) return rc;
  const size_t Mp = (M + 63) & ~static_cast<size_t>(63);  // keeps the sub-buffers 16-byte aligned (float4 stores)
  CK(h->ws_misc.ensure((2 * Mp + (size_t)Q * D) * sizeof(float)));
  float* keep = static_cast<float*>(h->ws_misc.p);
  float* pad_ext = keep + Mp;
  if (h->check_fp32) {   // the text tower in float32 (encoders.py:656-759)
    if ((rc = ensure_workspace_f32(h, M, D, 4 * D)) != VP_OK) return rc;
    float* xf = static_cast<float*>(h->c_x.p);
    h->mark(st, nullptr, false);
    CK(vp::f32::text_embed(st, ids, paddings, h->tok_emb, h->d_pe, h->cls_emb, xf, keep, pad_ext, Q, L, D, c.vocabulary_size)); h->mark(st, "text.embed");
    SeqLayout tlf{Q, S, 1, 1, pad_ext, keep};
    if ((rc = run_stack_f32(h, h->text, xf, (int)M, tlf, vp::ACT_RELU, st, kTagText)) != VP_OK) return rc;
    float* tmpf = static_cast<float*>(h->ws_misc.p) + 2 * Mp;
    vp::f32::LayerNorm lnf;   // unimodal_ln on the class token only (features[:, -1], encoders.py:756-758,:906)
    lnf.x = xf + (size_t)L * D; lnf.ldx = S * D; lnf.gamma1 = h->uni_ln_g; lnf.beta = h->uni_ln_b; lnf.y = normalize ? tmpf : text_emb; lnf.ldy = D;
    lnf.M = Q; lnf.D = D;
    CK(vp::f32::layernorm(st, lnf)); h->mark(st, "text.unimodal_ln");
    if (normalize) { CK(vp::launch_l2norm(st, tmpf, text_emb, Q, D)); h->mark(st, "text.l2norm"); }
    return VP_OK;
  }
  if ((rc = ensure_workspace(h, M, D, 4 * D)) != VP_OK) return rc;
  bf16* x = static_cast<bf16*>(h->ws_x.p);
  h->mark(st, nullptr, false);   // trace: start of this forward
  CK(vp::launch_text_embed(st, ids, paddings, h->tok_emb, h->d_pe, h->cls_emb, x, keep, pad_ext, Q, L, D, c.vocabulary_size)); h->mark(st, "text.embed");
  SeqLayout tl{Q, S, 1, 1, pad_ext, keep};
  float* stats_a = h->fuse_ln ? static_cast<float*>(h->ws_stats.p) : nullptr;
  float* stats_b = h->fuse_ln ? stats_a + h->stats_stride : nullptr;
  if (stats_a) { CK(vp::launch_row_stats(st, x, D, stats_a, (int)M, D)); h->mark(st, "text.row_stats"); }
  if ((rc = run_stack(h, h->text, x, (int)M, tl, vp::ACT_RELU, st, stats_a, 1, stats_b, kTagText)) != VP_OK) return rc;
  // unimodal_ln on the class token only (features[:, -1], encoders.py:756-758,:906), then l2 normalise
  float* tmp = static_cast<float*>(h->ws_misc.p) + 2 * Mp;
  vp::LnArgs ln;
  ln.x = x + (size_t)L * D; ln.ldx = S * D; ln.gamma1 = h->uni_ln_g; ln.beta = h->uni_ln_b;
  ln.y_f32 = normalize ? tmp : text_emb; ln.M = Q; ln.D = D;
  CK(vp::launch_layernorm(st, ln)); h->mark(st, "text.unimodal_ln");
  if (normalize) { CK(vp::launch_l2norm(st, tmp, text_emb, Q, D)); h->mark(st, "text.l2norm"); }
  return VP_OK;
}

int vp_clip_video_forward_host(vp_handle* h, const float* video, int B, int T, int H, int W, int normalize, float* video_emb,
                               void* stream) {
  return host_sync_call(h, 1, video, VP_F32, B, T, H, W, nullptr, video_emb, nullptr, normalize, stream);
}

int vp_clip_text_forward_host(vp_handle* h, const int32_t* ids, const float* paddings, int Q, int L, int normalize,
                              float* text_emb, void* stream) {
  DeviceScope device_scope(h ? h->device : -1);
  int rc = check_ready(h);
  if (rc != VP_OK) return rc;
  if (ids == nullptr || paddings == nullptr || text_emb == nullptr || Q <= 0 || L <= 0) return h->fail(VP_ERR_INVALID, "bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = (size_t)Q * L;
  CK(h->ws_io_in.ensure(n * 8 + 256));
  CK(h->ws_io_out.ensure((size_t)Q * h->cfg.model_dim * sizeof(float)));
  int32_t* d_ids = static_cast<int32_t*>(h->ws_io_in.p);
  float* d_pad = reinterpret_cast<float*>(d_ids + n);
  CK(cudaMemcpyAsync(d_ids, ids, n * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_pad, paddings, n * 4, cudaMemcpyHostToDevice, st));
  rc = vp_clip_text_forward(h, d_ids, d_pad, Q, L, normalize, static_cast<float*>(h->ws_io_out.p), stream);
  if (rc != VP_OK) return rc;
  CK(cudaMemcpyAsync(text_emb, h->ws_io_out.p, (size_t)Q * h->cfg.model_dim * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return VP_OK;
}

int vp_similarity(const float* v, const float* t, float* sim, int Nv, int Nt, int D, void* stream) {
  if (!v || !t || !sim || Nv <= 0 || Nt <= 0 || D <= 0) return VP_ERR_INVALID;
  return vp::launch_similarity(static_cast<cudaStream_t>(stream), v, t, sim, Nv, Nt, D) == cudaSuccess ? VP_OK : VP_ERR_CUDA;
}

size_t vp_workspace_bytes(const vp_handle* h, int B, int T, int H, int W) {
  if (!h || B <= 0 || T <= 0 || H <= 0 || W <= 0) return 0;
  const vp_config& c = h->cfg;
  const size_t N = (size_t)(H / c.patch_size) * (W / c.patch_size);
  if (N == 0) return 0;
  const size_t tokens_per_clip = (size_t)T * N;
  size_t clips = (size_t)B;
  if (c.kind == VP_KIND_ENCODER) {   // the encoder entry points run at most 2^18 tokens per pass (encoder_forward_dev)
    const size_t max_clips = std::max<size_t>(1, ((size_t)1 << 18) / tokens_per_clip);
    clips = std::min(clips, max_clips);
  }
  const size_t M = clips * tokens_per_clip;
  const size_t D = c.model_dim, F = c.mlp_dim;
  if (h->check_fp32) {   // fp32 check mode: x, n, qkv (3D), u and the patch matrix in float32 (+ the pooling head's scratch)
    size_t bytes32 = M * (5 * D + F + h->k_patch) * sizeof(float);
    if (c.kind != VP_KIND_ENCODER) bytes32 += (clips * tokens_per_clip * c.num_heads + clips * (size_t)c.num_heads * D + clips * ((size_t)c.num_heads * h->pool_ph + 2 * D)) * sizeof(float);
    return bytes32;
  }
  size_t bytes = M * (5 * D + F + h->k_patch_pad) * sizeof(bf16);                 // x, n, qkv (3D), u, patches
  const size_t slots = std::max(vp::gemm_stats_slots((int)D), 16);
  bytes += 2 * (M * 2 * slots + 64) * sizeof(float);                              // LayerNorm row statistics (two buffers)
  if (c.kind != VP_KIND_ENCODER)                                                  // pooling-head scratch
    bytes += vp::pool_scratch_floats((int)clips, (int)tokens_per_clip, (int)D, c.num_heads, h->pool_ph) * sizeof(float);
  return bytes;
}

int vp_release_workspace(vp_handle* h) {
  if (h == nullptr) return VP_ERR_INVALID;
  DeviceScope device_scope(h->device);
  // cudaFree waits for the device: forwards still in flight finish first.  The buffers grow again on the next call.
  DevBuf* bufs[] = {&h->ws_x, &h->ws_n, &h->ws_qkv, &h->ws_u, &h->ws_patch, &h->ws_misc, &h->ws_io_in, &h->ws_io_out, &h->ws_pool, &h->ws_stats,
                    &h->c_x, &h->c_n, &h->c_qkv, &h->c_u, &h->c_patch, &h->c_pool};
  for (DevBuf* b : bufs) b->release();
  h->stats_stride = 0;
  return VP_OK;
}

int64_t vp_kernel_launches(const vp_handle* h) { return h ? h->launches : 0; }

// In-situ timeline: with tracing on, every launch is followed by a CUDA event on the launch stream; the report
// aggregates the event-to-event times by label ("label count total_ms" per line, then "TOTAL n ms").
int vp_trace(vp_handle* h, int enable) {
  if (h == nullptr) return VP_ERR_INVALID;
  DeviceScope device_scope(h->device);
  for (auto& t : h->trace) cudaEventDestroy(t.second);
  h->trace.clear();
  h->trace_on = enable != 0;
  return VP_OK;
}

int vp_trace_report(vp_handle* h, char* buf, int cap) {
  if (h == nullptr || buf == nullptr || cap <= 0) return VP_ERR_INVALID;
  DeviceScope device_scope(h->device);
  CK(cudaDeviceSynchronize());
  std::vector<std::string> order;
  std::map<std::string, std::pair<int, double>> agg;
  double total = 0.0;
  int n = 0;
  for (size_t i = 1; i < h->trace.size(); ++i) {
    if (h->trace[i].first == nullptr) continue;   // start marker of the next forward: the gap before it is not a kernel
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->trace[i - 1].second, h->trace[i].second) != cudaSuccess) continue;
    auto it = agg.find(h->trace[i].first);
    if (it == agg.end()) { order.push_back(h->trace[i].first); it = agg.emplace(h->trace[i].first, std::make_pair(0, 0.0)).first; }
    it->second.first++; it->second.second += ms;
    total += ms; n++;
  }
  std::string out;
  char line[160];
  for (const std::string& k : order) {
    snprintf(line, sizeof(line), "%s %d %.6f\n", k.c_str(), agg[k].first, agg[k].second);
    out += line;
  }
  snprintf(line, sizeof(line), "TOTAL %d %.6f\n", n, total);
  out += line;
  if ((int)out.size() + 1 > cap) return h->fail(VP_ERR_INVALID, "trace report needs %d bytes", (int)out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return VP_OK;
}

int vp_device_sm_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return -1;
  return vp::num_sms();
}

// ------------------------------------------------------------ kernel-level entry points
static int ck(cudaError_t e) { return e == cudaSuccess ? VP_OK : (e == cudaErrorInvalidValue ? VP_ERR_INVALID : VP_ERR_CUDA); }

int vp_gemm_bf16(const void* A, int lda, const void* Wt, int ldb, void* C, int ldc, int M, int N, int K, const float* bias,
                 int act, const void* resid, int ldr, const float* row_scale, const float* pos_table, int pos_period,
                 int out_f32, void* stream) {
  vp::GemmEpilogue e;
  e.bias = bias; e.act = act; e.resid = static_cast<const bf16*>(resid); e.ldr = ldr; e.row_scale = row_scale;
  e.pos_table = pos_table; e.pos_period = pos_period; e.out_f32 = out_f32;
  return ck(vp::launch_gemm(static_cast<cudaStream_t>(stream), static_cast<const bf16*>(A), lda, static_cast<const bf16*>(Wt), ldb, C,
                            ldc, M, N, K, e));
}

int vp_gemm_bf16_ln(const void* A, int lda, const void* Wt, int ldb, void* C, int ldc, int M, int N, int K, const float* bias, int act,
                    const void* resid, int ldr, const float* ln_stats_in, int ln_slots, const float* ln_colsum, int ln_dim,
                    float* stats_out, void* stream) {
  vp::GemmEpilogue e;
  e.bias = bias; e.act = act; e.resid = static_cast<const bf16*>(resid); e.ldr = ldr;
  e.ln_stats_in = ln_stats_in; e.ln_slots = ln_slots; e.ln_colsum = ln_colsum; e.ln_dim = ln_dim; e.stats_out = stats_out;
  return ck(vp::launch_gemm(static_cast<cudaStream_t>(stream), static_cast<const bf16*>(A), lda, static_cast<const bf16*>(Wt), ldb, C,
                            ldc, M, N, K, e));
}

int vp_gemm_stats_slots(int N) { return vp::gemm_stats_slots(N); }

int vp_row_stats(const void* x, int ldx, float* stats, int M, int D, void* stream) {
  return ck(vp::launch_row_stats(static_cast<cudaStream_t>(stream), static_cast<const bf16*>(x), ldx, stats, M, D));
}

int vp_fold_ln_weight(const float* src, const float* gamma1, const float* beta, const float* bias_in, void* dst, float* colsum,
                      float* bias_out, int K, int N, int ldk, float scale, void* stream) {
  return ck(vp::launch_fold_ln_weight(static_cast<cudaStream_t>(stream), src, gamma1, beta, bias_in, static_cast<bf16*>(dst), colsum,
                                      bias_out, K, N, ldk, scale));
}

int vp_layernorm(const void* x, int ldx, const float* gamma1, const float* beta, void* y_bf16, float* y_f32,
                 const float* add_table, int add_div, int add_mod, int M, int D, void* stream) {
  vp::LnArgs a;
  a.x = static_cast<const bf16*>(x); a.ldx = ldx; a.gamma1 = gamma1; a.beta = beta; a.y_bf16 = static_cast<bf16*>(y_bf16);
  a.y_f32 = y_f32; a.add_table = add_table; a.add_div = add_div > 0 ? add_div : 1; a.add_mod = add_mod > 0 ? add_mod : 1;
  a.M = M; a.D = D;
  return ck(vp::launch_layernorm(static_cast<cudaStream_t>(stream), a));
}

int vp_patchify(const float* video, void* out, int ldo, int BT, int H, int W, int p, void* stream) {
  return ck(vp::launch_patchify(static_cast<cudaStream_t>(stream), video, static_cast<bf16*>(out), ldo, BT, H, W, p));
}

int vp_resize_frames_u8(const uint8_t* frames, int T, int H, int W, uint8_t* out, int target_size, int resize_mode, void* stream) {
  if (frames == nullptr || out == nullptr) return VP_ERR_INVALID;
  return ck(vp::launch_resize_frames_u8(static_cast<cudaStream_t>(stream), frames, T, H, W, out, target_size, resize_mode));
}

int vp_attention(const void* q, const void* k, const void* v, int ld, void* out, int ldo, int num_seq, int S, int group,
                 int heads, int dh, float cap, const float* key_pad, int causal, void* stream) {
  vp::AttnArgs a;
  a.q = static_cast<const bf16*>(q); a.k = static_cast<const bf16*>(k); a.v = static_cast<const bf16*>(v); a.ld = ld;
  a.out = static_cast<bf16*>(out); a.ldo = ldo; a.num_seq = num_seq; a.S = S; a.group = group; a.heads = heads; a.dh = dh;
  a.cap = cap; a.key_pad = key_pad; a.causal = causal & 1; a.force_mma_sync = (causal >> 1) & 1;
  return ck(vp::launch_attention(static_cast<cudaStream_t>(stream), a));
}

}  // extern "C"
