// Internal C++ interface between the engine (engine.cu), the C-ABI (capi.cu) and the
// kernel translation units.  Not part of the public ABI (see include/videoprism_b200.h).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vp {

typedef __nv_bfloat16 bf16;

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the (kernel, device) pair: a process that drives several
// GPUs must set it on each of them.  `cache` is a per-call-site array, one slot per device ordinal, holding the largest
// size already granted on that device; concurrent callers may both set the attribute, which is harmless.
constexpr int kMaxDevices = 64;
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, int bytes, int (&cache)[kMaxDevices]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const bool cached = dev >= 0 && dev < kMaxDevices;
  if (cached && cache[dev] >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && cached) cache[dev] = bytes;
  return e;
}

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

// Epilogue of C = A * W^T:  v = acc + bias[n]; v = act(v); v *= row_scale[m];
// v += pos_table[(m % pos_period), n]; v += resid[m, n]; store (bf16 or fp32).
struct GemmEpilogue {
  const float* bias = nullptr;       // [N]
  int act = ACT_NONE;
  const float* row_scale = nullptr;  // [M]  (1 - padding), layers.py:397-398,:410-411
  const float* pos_table = nullptr;  // [pos_period, N] fp32 (spatial pos-emb, encoders.py:514)
  int pos_period = 0;
  const bf16* resid = nullptr;       // [M, ldr] bf16 residual stream (layers.py:855, :425); may alias C
  int ldr = 0;
  int resid_period = 0;              // > 0: resid is a [resid_period, ldr] table indexed by (m % resid_period), resid_period % 32 == 0
                                     // (the spatial position table in bf16: the patch projection's `+ emb_var`, encoders.py:514)
  int out_f32 = 0;                   // C is float* when set
  // LayerNorm folded into the GEMM (layers.py:237-270 applied to the A rows):  A holds the RAW rows x, Wt holds
  // (gamma1 (.) W)^T, and   v = rstd[m] * acc - rstd[m] * mean[m] * ln_colsum[n] + bias[n]   with bias = beta.W + b.
  // ln_stats_in [M][ln_slots][2] = partial (sum, sum of squares) of each A row over its ln_dim features; the
  // partials are added in slot order (deterministic).
  const float* ln_stats_in = nullptr;
  int ln_slots = 1;
  const float* ln_colsum = nullptr;  // [N] column sums of the bf16-rounded (gamma1 (.) W)
  int ln_dim = 0;
  // Emit partial (sum, sum of squares) of every stored bf16 output row into stats_out [M][gemm_stats_slots(N)][2]:
  // one slot per (column tile, epilogue half), each written exactly once (no atomics, no zeroing).  The statistics
  // the NEXT LayerNorm-folded GEMM needs.  Requires N == the full feature dimension.
  float* stats_out = nullptr;
};
int gemm_stats_slots(int N);   // slots per row a GEMM with N output columns writes into stats_out

// C[M,N] = A[M,K] (bf16, row-major, lda) * Wt[N,K]^T (bf16, row-major, ldb) with epilogue.
// Requirements: K % 8 == 0, N % 8 == 0, lda % 8 == 0, ldb % 8 == 0, 16-byte aligned bases.
cudaError_t launch_gemm(cudaStream_t s, const bf16* A, int lda, const bf16* Wt, int ldb, void* C, int ldc,
                        int M, int N, int K, const GemmEpilogue& epi);

// LayerNorm over the last dim (layers.py:237-270): fp32 statistics, (1 + scale), + bias.
// y = LN(x) [+ add_table[(m / add_div) % add_mod]]; optional second fp32 copy of LN(x) (before the add).
struct LnArgs {
  const bf16* x = nullptr;  int ldx = 0;
  const float* gamma1 = nullptr;  // 1 + scale, [D]
  const float* beta = nullptr;    // [D]
  bf16* y_bf16 = nullptr;         // [M, D] or null
  float* y_f32 = nullptr;         // [M, D] or null (LN(x) without the table add)
  float* stats_out = nullptr;     // [M][1][2] (sum, sum of squares) of the stored bf16 rows y_bf16 (one slot), or null
  const float* add_table = nullptr; int add_div = 1; int add_mod = 1;   // temporal pos-emb, encoders.py:553
  const bf16* resid = nullptr; int ldr = 0;   // y_bf16 = LN(x) + resid (norm_policy 'primer_hybrid', layers.py:846-855); may alias y_bf16
  int M = 0, D = 0;
};
cudaError_t launch_layernorm(cudaStream_t s, const LnArgs& a);

// Patchify + cast (encoders.py:70-104): video [BT, H, W, 3] fp32 -> patches [BT*(H/p)*(W/p), ldo] bf16,
// column (py*p + px)*3 + c.  Columns >= p*p*3 are left untouched (zeroed once by the engine).
cudaError_t launch_patchify(cudaStream_t s, const float* video, bf16* out, int ldo, int BT, int H, int W, int p);
// same for uint8 frames: value / 255.0f in fp32 first, exactly as video_utils.load_video does (video_utils.py:88-93)
cudaError_t launch_patchify_u8(cudaStream_t s, const uint8_t* video, bf16* out, int ldo, int BT, int H, int W, int p);

// Frame ingest (video_utils.py:74-84, :97-127): uint8 RGB [T, H, W, 3] -> uint8 [T, target, target, 3], bit-exact with
// cv2.resize (INTER_LINEAR) + the reference's centre crop.  mode 0 = "center_crop", 1 = "resize".
cudaError_t launch_resize_frames_u8(cudaStream_t s, const uint8_t* src, int T, int H, int W, uint8_t* dst, int target, int mode);

// Attention over sequences embedded in a packed qkv buffer [rows, ld] (q at col q_off + h*dh, ...).
// Sequence `sid` token j lives at row (sid / group) * (group * S) + (sid % group) + j * group.
//   spatial stack : group = 1           (tokens of a frame are contiguous)
//   temporal stack: group = N patches   (tokens of a tube are N rows apart)
// scores = q . k (q is pre-scaled); cap * tanh(scores / cap) if cap > 0 (layers.py:586-594);
// masks (layers.py:51-179): key_pad [num_seq, S] (1 = padded) or null; causal => the reference's
// merged 2-D mask (query padded OR key padded OR key > query).  Fully masked rows are uniform.
struct AttnArgs {
  const bf16* q = nullptr; const bf16* k = nullptr; const bf16* v = nullptr; int ld = 0;
  bf16* out = nullptr; int ldo = 0;
  int num_seq = 0, S = 0, group = 1, heads = 0, dh = 0;
  float cap = 0.f;
  const float* key_pad = nullptr;
  int causal = 0;
  int force_mma_sync = 0;   // tests: bypass the tcgen05 kernel
  int pad_whole_seq = 0;    // key_pad is constant within every sequence (frame paddings in the spatial stack)
  int* launched = nullptr;  // optional: number of kernels launch_attention enqueued
};
cudaError_t launch_attention(cudaStream_t s, const AttnArgs& a);

// Weight repack: src fp32 [K, N] (row-major, Flax Dense kernel / [D,(N H)] projection)
// -> dst bf16 [N, ldk] = src^T * scale, columns K..ldk-1 zero.
cudaError_t launch_transpose_cast(cudaStream_t s, const float* src, bf16* dst, int K, int N, int ldk, float scale);
// dst bf16 [rows, cols] = src fp32 [rows, cols] * scale
cudaError_t launch_cast_bf16(cudaStream_t s, const float* src, bf16* dst, size_t n, float scale);
// dst fp32 = a * src + b
cudaError_t launch_affine_f32(cudaStream_t s, const float* src, float* dst, size_t n, float a, float b);

// (sum, sum of squares) of each bf16 row: stats [M][2]
cudaError_t launch_row_stats(cudaStream_t s, const bf16* x, int ldx, float* stats, int M, int D);
// LayerNorm folding of a projection weight (load time):  src fp32 [K, N] (Flax [in, out]) ->
//   dst bf16 [N, ldk] = (gamma1[k] * src[k, n] * scale)^T ;  colsum[n] = sum_k float(dst[n, k]) ;
//   bias_out[n] = sum_k beta[k] * src[k, n] * scale + bias_in[n] * scale
cudaError_t launch_fold_ln_weight(cudaStream_t s, const float* src, const float* gamma1, const float* beta, const float* bias_in,
                                  bf16* dst, float* colsum, float* bias_out, int K, int N, int ldk, float scale);

// y[rows, C] = x[rows, D] . w[D, C] + b[C], fp32 (classifier projection, encoders.py:643-650)
cudaError_t launch_dense_f32(cudaStream_t s, const float* x, const float* w, const float* b, float* y, int rows, int D, int C);

// L2 normalise rows in fp32 (encoders.py:50-67): y = x / sqrt(sum(x^2) + 1e-12)
cudaError_t launch_l2norm(cudaStream_t s, const float* x, float* y, int rows, int D);

// Text embedding (encoders.py:708-740): x[q, j] = emb[ids[q,j]] * sqrt(D) + pe[j] for j < L,
// x[q, L] = cls * sqrt(D); bf16 out [Q*(L+1), D]; also writes keep[q*(L+1)+j] = 1 - pad.
cudaError_t launch_text_embed(cudaStream_t s, const int32_t* ids, const float* pad, const float* emb, const float* pe,
                              const float* cls, bf16* x, float* keep, float* pad_ext, int Q, int L, int D, int vocab);

}  // namespace vp
