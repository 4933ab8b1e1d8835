// fp32 CHECK MODE kernels: the whole forward in float32 on the CUDA cores (fp32 residual stream, fp32 LayerNorm output,
// fp32 GEMM operands and accumulators, fp32 softmax and P), evaluated exactly as the reference writes it
// (get_model(name) computes in float32 by default, layers.py:182-205).  This is the second, independent opinion on the
// GPU: the bf16 tensor-core path is held to a cosine bound, this path to the reference's own Flax-vs-MLX envelope
// (max-abs <= 1e-3 on features, <= 1e-5 on normalised embeddings; FLAX_TO_MLX_CONVERSION_GUIDE.md:321-358,
// verify_clip_models.py:92-95).  Throughput is not a goal here (SIMT FFMA, ~20-30 TFLOP/s); accuracy is: libm-grade
// expf / tanhf / erff, no approximations, no reduced-precision operands.
#include <math_constants.h>
#include <stdint.h>

#include "check_fp32.h"

namespace vp {
namespace f32 {

namespace {

// ------------------------------------------------------------------------------------------------ SGEMM
// C[M,N] = epilogue(A[M,K] . W), A row-major fp32.  W_NK = false: W is [K, N] (Flax Dense kernels, q/k/v projections);
// W_NK = true: W is [N, K] (the attention output projection 'post/w' [D_out, N, H], layers.py:483).
// 128 x 128 tile per block of 256 threads, 8 x 8 outputs per thread (two 4-wide groups 64 apart in each dimension, so
// the shared-memory reads are contiguous across the lanes), K in slabs of 16.
constexpr int BM = 128, BN = 128, BK = 16;

template <bool W_NK>
__global__ void __launch_bounds__(256) sgemm_kernel(const Sgemm a) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.K; k0 += BK) {
    {  // A tile: thread -> row tid / 2, 8 consecutive k
      const int r = tid >> 1, kk = (tid & 1) * 8;
      const int m = m0 + r;
      const float* src = a.A + static_cast<size_t>(m) * a.lda + k0 + kk;
#pragma unroll
      for (int j = 0; j < 8; ++j) As[kk + j][r] = (m < a.M && k0 + kk + j < a.K) ? src[j] : 0.f;
    }
    if (W_NK) {  // W[n][k]: thread -> column n = tid / 2, 8 consecutive k
      const int c = tid >> 1, kk = (tid & 1) * 8;
      const int n = n0 + c;
      const float* src = a.W + static_cast<size_t>(n) * a.ldw + k0 + kk;
#pragma unroll
      for (int j = 0; j < 8; ++j) Bs[kk + j][c] = (n < a.N && k0 + kk + j < a.K) ? src[j] : 0.f;
    } else {     // W[k][n]: thread -> k = tid / 16, 8 consecutive n
      const int kk = tid >> 4, c = (tid & 15) * 8;
      const float* src = a.W + static_cast<size_t>(k0 + kk) * a.ldw + n0 + c;
#pragma unroll
      for (int j = 0; j < 8; ++j) Bs[kk][c + j] = (k0 + kk < a.K && n0 + c + j < a.N) ? src[j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // epilogue, in the reference's order: (x.W + b) [* alpha: the query scale] -> activation -> * (1 - padding) -> + table -> + residual
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= a.M) continue;
    const float rs = a.row_scale ? a.row_scale[m] : 1.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (n >= a.N) continue;
      float v = acc[i][j];
      if (a.bias) v += a.bias[n];
      v *= a.alpha;
      if (a.act == 1) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));   // jax.nn.gelu(approximate=False), layers.py:31
      else if (a.act == 2) v = fmaxf(v, 0.f);
      if (a.row_scale) v *= rs;
      if (a.pos_table) v += a.pos_table[static_cast<size_t>(m % a.pos_period) * a.N + n];
      if (a.resid) v += a.resid[static_cast<size_t>(m) * a.ldr + n];
      a.C[static_cast<size_t>(m) * a.ldc + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// layers.py:237-270: mean, biased variance, rsqrt(var + 1e-6), * (1 + scale), + bias; one warp per row, two passes.
__global__ void __launch_bounds__(256) layernorm_kernel(const LayerNorm a) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= a.M) return;
  const float* x = a.x + static_cast<size_t>(row) * a.ldx;
  float s = 0.f;
  for (int d = lane; d < a.D; d += 32) s += x[d];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / a.D;
  float q = 0.f;
  for (int d = lane; d < a.D; d += 32) { const float t = x[d] - mean; q += t * t; }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / a.D + 1e-6f);
  const float* tab = a.add_table ? a.add_table + static_cast<size_t>((row / a.add_div) % a.add_mod) * a.D : nullptr;
  for (int d = lane; d < a.D; d += 32) {
    const float v = (x[d] - mean) * rstd * a.gamma1[d] + a.beta[d];
    if (a.y2) a.y2[static_cast<size_t>(row) * a.D + d] = v;
    if (a.y) a.y[static_cast<size_t>(row) * a.ldy + d] = tab ? v + tab[d] : v;
  }
}

// ------------------------------------------------------------------------------------------------ attention
// layers.py:601-661 in fp32: logits = q . k (q pre-scaled), cap * tanh(logits / cap), masks (layers.py:51-179: padded keys,
// causal, padded queries; a fully masked row is uniform because the mask value is finite), softmax, P . v.
// One thread per query, keys / values staged in shared memory in tiles of KT keys, online softmax per tile.
constexpr int KT = 16;
constexpr float kMasked = -1.0e30f;

template <int DH>
__global__ void __launch_bounds__(64) attention_kernel(const Attention a) {
  __shared__ __align__(16) float sK[KT][DH];
  __shared__ __align__(16) float sV[KT][DH];
  __shared__ float sPad[KT];
  const int sid = blockIdx.y, h = blockIdx.z;
  const int qi = blockIdx.x * 64 + threadIdx.x;
  const size_t row0 = static_cast<size_t>(sid / a.group) * (static_cast<size_t>(a.group) * a.S) + (sid % a.group);
  const bool active = qi < a.S;
  float q[DH], o[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) { q[d] = 0.f; o[d] = 0.f; }
  if (active) {
    const float* qp = a.q + (row0 + static_cast<size_t>(qi) * a.group) * a.ld + h * a.dh;
#pragma unroll
    for (int d = 0; d < DH; ++d) q[d] = d < a.dh ? qp[d] : 0.f;
  }
  const bool qpad = active && a.causal && a.key_pad != nullptr && a.key_pad[static_cast<size_t>(sid) * a.S + qi] > 0.5f;
  float m = -CUDART_INF_F, l = 0.f;
  for (int j0 = 0; j0 < a.S; j0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * DH; i += 64) {
      const int j = i / DH, d = i % DH;
      const bool ok = j0 + j < a.S && d < a.dh;
      const size_t r = (row0 + static_cast<size_t>(j0 + j) * a.group) * a.ld + h * a.dh + d;
      sK[j][d] = ok ? a.k[r] : 0.f;
      sV[j][d] = ok ? a.v[r] : 0.f;
    }
    if (threadIdx.x < KT) sPad[threadIdx.x] = (a.key_pad && j0 + threadIdx.x < a.S) ? a.key_pad[static_cast<size_t>(sid) * a.S + j0 + threadIdx.x] : 0.f;
    __syncthreads();
    if (!active) continue;
    float s[KT];
    float tmax = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) acc = fmaf(q[d], sK[j][d], acc);
      if (a.cap > 0.f) acc = a.cap * tanhf(acc / a.cap);
      bool masked = sPad[j] > 0.5f;
      if (a.causal) masked = masked || (j0 + j > qi) || qpad;
      acc = masked ? kMasked : acc;
      s[j] = (j0 + j < a.S) ? acc : -CUDART_INF_F;
      tmax = fmaxf(tmax, s[j]);
    }
    const float mn = fmaxf(m, tmax);
    const float corr = expf(m - mn);   // exp(-inf) = 0 on the first tile
    l *= corr;
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] *= corr;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float p = expf(s[j] - mn);
      l += p;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = fmaf(p, sV[j][d], o[d]);
    }
    m = mn;
  }
  if (active) {
    const float inv = 1.0f / l;
    float* op = a.out + (row0 + static_cast<size_t>(qi) * a.group) * a.ldo + h * a.dh;
#pragma unroll
    for (int d = 0; d < DH; ++d)
      if (d < a.dh) op[d] = o[d] * inv;
  }
}

// ------------------------------------------------------------------------------------------------ glue
// patches (p q c) per token (encoders.py:95-103); uint8 frames are divided by 255 first (video_utils.py:91)
template <typename TIn>
__global__ void patchify_kernel(const TIn* __restrict__ video, float* __restrict__ out, int H, int W, int p, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int kp = p * p * 3;
  const int col = static_cast<int>(i % kp);
  const size_t tok = i / kp;
  const int gw = W / p, gh = H / p;
  const int gx = static_cast<int>(tok % gw), gy = static_cast<int>((tok / gw) % gh);
  const size_t bt = tok / (static_cast<size_t>(gw) * gh);
  const int c = col % 3, px = (col / 3) % p, py = col / (3 * p);
  const size_t src = ((bt * H + gy * p + py) * W + gx * p + px) * 3 + c;
  out[i] = sizeof(TIn) == 1 ? __fdiv_rn(static_cast<float>(video[src]), 255.0f) : static_cast<float>(video[src]);
}

// encoders.py:708-740: emb[id] * sqrt(D) + PE on the L real positions, cls * sqrt(D) appended
__global__ void text_embed_kernel(const int32_t* __restrict__ ids, const float* __restrict__ pad, const float* __restrict__ emb,
                                  const float* __restrict__ pe, const float* __restrict__ cls, float* __restrict__ x,
                                  float* __restrict__ keep, float* __restrict__ pad_ext, int L, int D, int vocab) {
  const int row = blockIdx.x;
  const int q = row / (L + 1), j = row % (L + 1);
  const float sq = sqrtf(static_cast<float>(D));
  float* xr = x + static_cast<size_t>(row) * D;
  if (j < L) {
    int id = ids[q * L + j];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float* er = emb + static_cast<size_t>(id) * D;
    const float* pr = pe + static_cast<size_t>(j) * D;
    for (int c = threadIdx.x; c < D; c += blockDim.x) xr[c] = er[c] * sq + pr[c];
    if (threadIdx.x == 0) { const float pv = pad[q * L + j]; keep[row] = 1.0f - pv; pad_ext[row] = pv; }
  } else {
    for (int c = threadIdx.x; c < D; c += blockDim.x) xr[c] = cls[c] * sq;
    if (threadIdx.x == 0) { keep[row] = 1.0f; pad_ext[row] = 0.0f; }
  }
}

// Pooling head, one block per (sequence, head): p = softmax_s(scores[s, h]) (layers.py:1093-1121 with the single learned
// query folded into the key projection), xbar[h, :] = sum_s p_s x[s, :].
__global__ void __launch_bounds__(256) pool_kernel(const float* __restrict__ x, const float* __restrict__ scores, float* __restrict__ xbar,
                                                   int S, int D, int H) {
  extern __shared__ float sp[];   // [S]
  __shared__ float red[8];
  const int seq = blockIdx.x, h = blockIdx.y;
  const float* sc = scores + static_cast<size_t>(seq) * S * H + h;
  auto block_reduce = [&](float v, bool is_max) {
#pragma unroll
    for (int o = 16; o; o >>= 1) { const float t = __shfl_xor_sync(0xffffffffu, v, o); v = is_max ? fmaxf(v, t) : v + t; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < 8; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
    return r;
  };
  float mx = -CUDART_INF_F;
  for (int s = threadIdx.x; s < S; s += 256) mx = fmaxf(mx, sc[static_cast<size_t>(s) * H]);
  mx = block_reduce(mx, true);
  float sum = 0.f;
  for (int s = threadIdx.x; s < S; s += 256) { const float e = expf(sc[static_cast<size_t>(s) * H] - mx); sp[s] = e; sum += e; }
  sum = block_reduce(sum, false);
  const float inv = 1.0f / sum;
  for (int d = threadIdx.x; d < D; d += 256) {
    const float* xp = x + static_cast<size_t>(seq) * S * D + d;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc = fmaf(sp[s], xp[static_cast<size_t>(s) * D], acc);
    xbar[(static_cast<size_t>(seq) * H + h) * D + d] = acc * inv;
  }
}

__global__ void cast_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(src[i]);
}

}  // namespace

cudaError_t sgemm(cudaStream_t s, const Sgemm& a) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return cudaSuccess;
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM);
  if (a.w_nk) sgemm_kernel<true><<<grid, 256, 0, s>>>(a);
  else sgemm_kernel<false><<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t layernorm(cudaStream_t s, const LayerNorm& a) {
  if (a.M <= 0) return cudaSuccess;
  layernorm_kernel<<<(a.M + 7) / 8, 256, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t attention(cudaStream_t s, const Attention& a) {
  if (a.num_seq <= 0 || a.S <= 0) return cudaSuccess;
  if (a.dh <= 0 || a.dh > 128) return cudaErrorInvalidValue;
  dim3 grid((a.S + 63) / 64, a.num_seq, a.heads);
  if (a.dh <= 32) attention_kernel<32><<<grid, 64, 0, s>>>(a);
  else if (a.dh <= 64) attention_kernel<64><<<grid, 64, 0, s>>>(a);
  else if (a.dh <= 96) attention_kernel<96><<<grid, 64, 0, s>>>(a);
  else attention_kernel<128><<<grid, 64, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t patchify(cudaStream_t s, const void* video, int is_u8, float* out, int BT, int H, int W, int p) {
  const size_t total = static_cast<size_t>(BT) * (H / p) * (W / p) * p * p * 3;
  if (total == 0) return cudaSuccess;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (is_u8) patchify_kernel<uint8_t><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(video), out, H, W, p, total);
  else patchify_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(video), out, H, W, p, total);
  return cudaGetLastError();
}

cudaError_t text_embed(cudaStream_t s, const int32_t* ids, const float* pad, const float* emb, const float* pe, const float* cls,
                       float* x, float* keep, float* pad_ext, int Q, int L, int D, int vocab) {
  if (Q <= 0) return cudaSuccess;
  text_embed_kernel<<<Q * (L + 1), 128, 0, s>>>(ids, pad, emb, pe, cls, x, keep, pad_ext, L, D, vocab);
  return cudaGetLastError();
}

cudaError_t pool(cudaStream_t s, const float* x, const float* scores, float* xbar, int num_seq, int S, int D, int H) {
  if (num_seq <= 0) return cudaSuccess;
  const size_t smem = static_cast<size_t>(S) * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static int granted[kMaxDevices] = {};
  const cudaError_t e = ensure_dynamic_smem(pool_kernel, static_cast<int>(smem), granted);
  if (e != cudaSuccess) return e;
  pool_kernel<<<dim3(num_seq, H), 256, smem, s>>>(x, scores, xbar, S, D, H);
  return cudaGetLastError();
}

cudaError_t cast_to_bf16(cudaStream_t s, const float* src, void* dst, size_t n) {
  if (n == 0) return cudaSuccess;
  cast_to_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(src, static_cast<__nv_bfloat16*>(dst), n);
  return cudaGetLastError();
}

}  // namespace f32
}  // namespace vp
