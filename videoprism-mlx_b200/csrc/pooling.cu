// Attention-pooling head (AttenTokenPoolingLayer, layers.py:1044-1136) in its single-query form,
// frame-padding expansion, and the retrieval similarity matrix.
//
// The pooler has ONE learned query, so with qh[h,:] the projected + PerDimScale'd query (a constant,
// computed at load time in engine.cu:finalize_pooler):
//   scores[s,h] = k[s,h,:] . qh[h,:] = x[s] . wkq[h,:] + const(h)          (const cancels in softmax)
//   ctx[h,:]    = sum_s p[s,h] v[s,h,:] = (sum_s p[s,h] x[s]) . Wv[:,h,:] + bv[h,:]
//   out         = LN( ctx . post.w + post.b ) ; optional l2 normalise (encoders.py:50-67)
// which replaces the reference's [S,D]x[D,4D] key/value projections (38.7 GF per base clip) by two
// passes over x (0.15 GF); everything here is bandwidth-bound fp32 CUDA-core work.
#include <math_constants.h>

#include "kernels.h"
#include "ptx.cuh"

namespace vp {

namespace {

constexpr int kMaxHeads = 16;
constexpr int kChunk = 256;  // tokens per accumulation block

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// scores[row, h] = x[row, :] . wkq[h, :]      (one warp per row, wkq staged in smem)
__global__ void __launch_bounds__(256) pool_scores_kernel(const bf16* __restrict__ x, const float* __restrict__ wkq,
                                                          float* __restrict__ scores, int rows, int D, int H) {
  extern __shared__ float s_w[];  // [H][D]
  for (int i = threadIdx.x; i < H * D; i += blockDim.x) s_w[i] = wkq[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + warp; row < rows; row += gridDim.x * wpb) {
    float acc[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) acc[h] = 0.f;
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * D);
    for (int vi = lane; vi < D / 8; vi += 32) {
      const uint4 u = xr[vi];
      float xv[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h) {
        if (h < H) {
          const float4 w0 = *reinterpret_cast<const float4*>(s_w + h * D + vi * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(s_w + h * D + vi * 8 + 4);
          acc[h] += xv[0] * w0.x + xv[1] * w0.y + xv[2] * w0.z + xv[3] * w0.w + xv[4] * w1.x + xv[5] * w1.y + xv[6] * w1.z + xv[7] * w1.w;
        }
      }
    }
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) {
      if (h < H) {
        const float v = warp_sum(acc[h]);
        if (lane == 0) scores[static_cast<size_t>(row) * H + h] = v;
      }
    }
  }
}

// stats[seq, h] = (max_s scores, sum_s exp(scores - max))    (one block per sequence, one warp per head in turn)
__global__ void __launch_bounds__(256) pool_stats_kernel(const float* __restrict__ scores, float* __restrict__ stats, int S, int H) {
  const int seq = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* sc = scores + static_cast<size_t>(seq) * S * H;
  for (int h = warp; h < H; h += nw) {
    float m = -CUDART_INF_F;
    for (int s = lane; s < S; s += 32) m = fmaxf(m, sc[static_cast<size_t>(s) * H + h]);
    m = warp_max(m);
    float sum = 0.f;
    for (int s = lane; s < S; s += 32) sum += __expf(sc[static_cast<size_t>(s) * H + h] - m);
    sum = warp_sum(sum);
    if (lane == 0) {
      stats[(static_cast<size_t>(seq) * H + h) * 2 + 0] = m;
      stats[(static_cast<size_t>(seq) * H + h) * 2 + 1] = sum;
    }
  }
}

// partial[seq, chunk, h, d] = sum_{s in chunk} p[s,h] x[s,d]
__global__ void __launch_bounds__(256) pool_accum_kernel(const bf16* __restrict__ x, const float* __restrict__ scores,
                                                         const float* __restrict__ stats, float* __restrict__ partial, int S, int D,
                                                         int H, int nchunk) {
  __shared__ float s_p[kChunk][kMaxHeads];
  const int chunk = blockIdx.x, seq = blockIdx.y;
  const int s0 = chunk * kChunk;
  const int ns = min(kChunk, S - s0);
  for (int i = threadIdx.x; i < ns * H; i += blockDim.x) {
    const int s = i / H, h = i % H;
    const float m = stats[(static_cast<size_t>(seq) * H + h) * 2 + 0];
    const float sum = stats[(static_cast<size_t>(seq) * H + h) * 2 + 1];
    s_p[s][h] = __expf(scores[(static_cast<size_t>(seq) * S + s0 + s) * H + h] - m) / sum;
  }
  __syncthreads();
  for (int d0 = 0; d0 < D; d0 += blockDim.x) {
    const int d = d0 + threadIdx.x;
    if (d >= D) break;
    float acc[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) acc[h] = 0.f;
    const bf16* xp = x + (static_cast<size_t>(seq) * S + s0) * D + d;
    for (int s = 0; s < ns; ++s) {
      const float xv = __bfloat162float(xp[static_cast<size_t>(s) * D]);
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h) acc[h] += s_p[s][h] * xv;   // rows of s_p beyond H hold stale data but are never stored
    }
    float* pp = partial + (static_cast<size_t>(seq) * nchunk + chunk) * H * D;
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h)
      if (h < H) pp[static_cast<size_t>(h) * D + d] = acc[h];
  }
}

// ctx[seq, h*dh + j] = (sum_chunks partial[seq, :, h, :]) . Wv[:, h, j] + bv[h, j]     grid (H, num_seq), block dh
__global__ void pool_ctx_kernel(const float* __restrict__ partial, const bf16* __restrict__ wv /*[D, H*dh]*/,
                                const float* __restrict__ bv, float* __restrict__ ctx, int D, int H, int dh, int nchunk) {
  extern __shared__ float s_xbar[];  // [D]
  const int h = blockIdx.x, seq = blockIdx.y;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a = 0.f;
    for (int c = 0; c < nchunk; ++c) a += partial[((static_cast<size_t>(seq) * nchunk + c) * H + h) * D + d];
    s_xbar[d] = a;
  }
  __syncthreads();
  const int j = threadIdx.x;
  if (j < dh) {
    float a = 0.f;
    const bf16* wp = wv + static_cast<size_t>(h) * dh + j;
    for (int d = 0; d < D; ++d) a += s_xbar[d] * __bfloat162float(wp[static_cast<size_t>(d) * H * dh]);
    ctx[static_cast<size_t>(seq) * H * dh + h * dh + j] = a + bv[h * dh + j];
  }
}

// out[seq, :] = [l2norm] LN( ctx[seq, :] . wpost[d, :] + bpost[d] )       one block per sequence
__global__ void __launch_bounds__(256) pool_out_kernel(const float* __restrict__ ctx, const bf16* __restrict__ wpost /*[D, HD]*/,
                                                       const float* __restrict__ bpost, const float* __restrict__ g1,
                                                       const float* __restrict__ beta, float* __restrict__ out, int D, int HD,
                                                       int normalize) {
  extern __shared__ float sm[];  // ctx [HD] | y [D] | red [32]
  float* s_ctx = sm;
  float* s_y = sm + HD;
  float* s_red = s_y + D;
  const int seq = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < HD; i += blockDim.x) s_ctx[i] = ctx[static_cast<size_t>(seq) * HD + i];
  __syncthreads();
  for (int d = warp; d < D; d += nw) {
    const bf16* wr = wpost + static_cast<size_t>(d) * HD;
    float a = 0.f;
    for (int i = lane; i < HD; i += 32) a += s_ctx[i] * __bfloat162float(wr[i]);
    a = warp_sum(a);
    if (lane == 0) s_y[d] = a + bpost[d];
  }
  __syncthreads();
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < nw) ? s_red[threadIdx.x] : 0.f;
    if (warp == 0) {
      t = warp_sum(t);
      if (lane == 0) s_red[0] = t;
    }
    __syncthreads();
    const float r = s_red[0];
    __syncthreads();
    return r;
  };
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) s += s_y[d];
  const float mean = block_sum(s) / D;
  float sq = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) { const float t = s_y[d] - mean; sq += t * t; }
  const float rstd = rsqrtf(block_sum(sq) / D + 1e-6f);
  float nn = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float y = (s_y[d] - mean) * rstd * g1[d] + beta[d];
    s_y[d] = y;
    nn += y * y;
  }
  const float tot = block_sum(nn);
  const float inv = normalize ? 1.0f / sqrtf(tot + 1e-12f) : 1.0f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) out[static_cast<size_t>(seq) * D + d] = s_y[d] * inv;
}

__global__ void pad_expand_kernel(const float* __restrict__ fp, float* __restrict__ pad_tok, float* __restrict__ keep_tok,
                                  float* __restrict__ pad_tube, int B, int T, int N) {
  const size_t m = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t M = static_cast<size_t>(B) * T * N;
  if (m >= M) return;
  const int n = static_cast<int>(m % N);
  const int t = static_cast<int>((m / N) % T);
  const int b = static_cast<int>(m / (static_cast<size_t>(N) * T));
  const float p = fp[b * T + t];
  pad_tok[m] = p;
  keep_tok[m] = 1.0f - p;
  pad_tube[(static_cast<size_t>(b) * N + n) * T + t] = p;
}

__global__ void similarity_kernel(const float* __restrict__ v, const float* __restrict__ t, float* __restrict__ sim, int Nv, int Nt,
                                  int D) {
  __shared__ float sv[16][17], st[16][17];
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  float acc = 0.f;
  for (int d0 = 0; d0 < D; d0 += 16) {
    const int vi = blockIdx.y * 16 + threadIdx.y, tj = blockIdx.x * 16 + threadIdx.y;
    sv[threadIdx.y][threadIdx.x] = (vi < Nv && d0 + threadIdx.x < D) ? v[static_cast<size_t>(vi) * D + d0 + threadIdx.x] : 0.f;
    st[threadIdx.y][threadIdx.x] = (tj < Nt && d0 + threadIdx.x < D) ? t[static_cast<size_t>(tj) * D + d0 + threadIdx.x] : 0.f;
    __syncthreads();
#pragma unroll
    for (int d = 0; d < 16; ++d) acc += sv[threadIdx.y][d] * st[threadIdx.x][d];
    __syncthreads();
  }
  if (i < Nv && j < Nt) sim[static_cast<size_t>(i) * Nt + j] = acc;
}

// y[r, c] = x[r, :] . w[:, c] + b[c]  (fp32; the classifier's `projection`, encoders.py:643-650).  w is [D, C] as Flax stores
// Dense kernels, so threads over c read it coalesced; x[r] is staged in shared memory.  Tiny (B x D x C).
__global__ void __launch_bounds__(256) dense_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, float* __restrict__ y, int D, int Cn) {
  extern __shared__ float s_x[];  // [D]
  const int r = blockIdx.y;
  for (int d = threadIdx.x; d < D; d += blockDim.x) s_x[d] = x[static_cast<size_t>(r) * D + d];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cn) return;
  float a0 = 0.f, a1 = 0.f;
  int d = 0;
  for (; d + 1 < D; d += 2) {
    a0 = fmaf(s_x[d], w[static_cast<size_t>(d) * Cn + c], a0);
    a1 = fmaf(s_x[d + 1], w[static_cast<size_t>(d + 1) * Cn + c], a1);
  }
  if (d < D) a0 = fmaf(s_x[d], w[static_cast<size_t>(d) * Cn + c], a0);
  y[static_cast<size_t>(r) * Cn + c] = a0 + a1 + b[c];
}

}  // namespace

size_t pool_scratch_floats(int num_seq, int S, int D, int H, int dh) {
  const size_t nchunk = (S + kChunk - 1) / kChunk;
  return static_cast<size_t>(num_seq) * S * H + static_cast<size_t>(num_seq) * H * 2 + static_cast<size_t>(num_seq) * nchunk * H * D +
         static_cast<size_t>(num_seq) * H * dh + 64;
}

cudaError_t launch_pool(cudaStream_t s, const bf16* x, int num_seq, int S, int D, int H, int dh, const float* wkq, const bf16* wv,
                        const float* bv, const bf16* wpost, const float* bpost, const float* ln_g1, const float* ln_b, int normalize,
                        float* scratch, float* out, int64_t* launches) {
  if (H > kMaxHeads || (D % 8) || dh > 1024) return cudaErrorInvalidValue;
  const int nchunk = (S + kChunk - 1) / kChunk;
  float* scores = scratch;
  float* stats = scores + static_cast<size_t>(num_seq) * S * H;
  float* partial = stats + static_cast<size_t>(num_seq) * H * 2;
  float* ctx = partial + static_cast<size_t>(num_seq) * nchunk * H * D;
  const int rows = num_seq * S;
  cudaError_t e;
  {
    const size_t smem = static_cast<size_t>(H) * D * sizeof(float);
    static int granted[kMaxDevices] = {};
    if (smem > 48 * 1024 && (e = ensure_dynamic_smem(pool_scores_kernel, static_cast<int>(smem), granted)) != cudaSuccess) return e;
    int grid = (rows + 7) / 8;
    if (grid > 148 * 8) grid = 148 * 8;
    pool_scores_kernel<<<grid, 256, smem, s>>>(x, wkq, scores, rows, D, H);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
  }
  pool_stats_kernel<<<num_seq, 256, 0, s>>>(scores, stats, S, H);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  pool_accum_kernel<<<dim3(nchunk, num_seq), 256, 0, s>>>(x, scores, stats, partial, S, D, H, nchunk);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  const int ctx_threads = ((dh + 31) / 32) * 32;
  pool_ctx_kernel<<<dim3(H, num_seq), ctx_threads, D * sizeof(float), s>>>(partial, wv, bv, ctx, D, H, dh, nchunk);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  pool_out_kernel<<<num_seq, 256, (static_cast<size_t>(H) * dh + D + 32) * sizeof(float), s>>>(ctx, wpost, bpost, ln_g1, ln_b, out, D,
                                                                                                H * dh, normalize);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (launches) *launches += 5;
  return cudaSuccess;
}

cudaError_t launch_pad_expand(cudaStream_t s, const float* frame_pad, float* pad_tok, float* keep_tok, float* pad_tube, int B, int T,
                              int N) {
  const size_t M = static_cast<size_t>(B) * T * N;
  pad_expand_kernel<<<static_cast<unsigned>((M + 255) / 256), 256, 0, s>>>(frame_pad, pad_tok, keep_tok, pad_tube, B, T, N);
  return cudaGetLastError();
}

cudaError_t launch_similarity(cudaStream_t s, const float* v, const float* t, float* sim, int Nv, int Nt, int D) {
  dim3 grid((Nt + 15) / 16, (Nv + 15) / 16), block(16, 16);
  similarity_kernel<<<grid, block, 0, s>>>(v, t, sim, Nv, Nt, D);
  return cudaGetLastError();
}

cudaError_t launch_dense_f32(cudaStream_t s, const float* x, const float* w, const float* b, float* y, int rows, int D, int C) {
  if (rows <= 0 || D <= 0 || C <= 0 || static_cast<size_t>(D) * sizeof(float) > 48 * 1024) return cudaErrorInvalidValue;
  dense_f32_kernel<<<dim3((C + 255) / 256, rows), 256, D * sizeof(float), s>>>(x, w, b, y, D, C);
  return cudaGetLastError();
}

}  // namespace vp
