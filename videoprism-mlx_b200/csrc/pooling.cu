// Attention-pooling head (AttenTokenPoolingLayer, layers.py:1044-1136) in its single-query form,
// frame-padding expansion, and the retrieval similarity matrix.
//
// The pooler has ONE learned query, so with qh[h,:] the projected + PerDimScale'd query (a constant,
// computed at load time in engine.cu:finalize_pooler):
//   scores[s,h] = k[s,h,:] . qh[h,:] = x[s] . wkq[h,:] + const(h)          (const cancels in softmax)
//   ctx[h,:]    = sum_s p[s,h] v[s,h,:] = (sum_s p[s,h] x[s]) . Wv[:,h,:] + bv[h,:]
//   out         = LN( ctx . post.w + post.b ) ; optional l2 normalise (encoders.py:50-67)
// which replaces the reference's [S,D]x[D,4D] key/value projections (38.7 GF per base clip) by two
// passes over x (0.15 GF): the score pass runs on the tensor core (launch_gemm), the rest is bandwidth- / latency-bound
// fp32 CUDA-core work laid out for parallelism across sequences (a retrieval batch pools 32+ sequences at once).
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

// scores[row, 0..31] = x[row, :] . wkq_hl[0..31, :] comes from the tcgen05 GEMM (fp32 accumulators written as they are):
// columns h and H + h hold the products with the high and the low bf16 part of the folded weight; their sum is the score.
constexpr int kScoreLd = 32;

// stats[seq, h] = (max_s scores, sum_s exp(scores - max))    one block of 32 warps per sequence.  A warp reads whole score
// rows (32 floats = one 128-byte line: lane h and lane H + h hold the two parts of head h's score) and keeps a running
// (max, sum) per lane; the 32 warps are merged in a fixed order.
__global__ void __launch_bounds__(1024) pool_stats_kernel(const float* __restrict__ scores, float* __restrict__ stats, int S, int H) {
  __shared__ float s_m[32][kMaxHeads], s_l[32][kMaxHeads];
  const int seq = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* sc = scores + static_cast<size_t>(seq) * S * kScoreLd;
  float m = -CUDART_INF_F, l = 0.f;
  for (int s = warp; s < S; s += 32) {
    const float r = sc[static_cast<size_t>(s) * kScoreLd + lane];
    const float v = r + __shfl_sync(0xffffffffu, r, (lane + H) & 31);   // meaningful for lane < H
    const float mn = fmaxf(m, v);
    l = l * __expf(m - mn) + __expf(v - mn);
    m = mn;
  }
  if (lane < kMaxHeads) { s_m[warp][lane] = m; s_l[warp][lane] = l; }
  __syncthreads();
  if (threadIdx.x < H) {
    const int h = threadIdx.x;
    float mm = s_m[0][h];
    for (int w = 1; w < 32; ++w) mm = fmaxf(mm, s_m[w][h]);
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += s_l[w][h] * __expf(s_m[w][h] - mm);   // warps without rows: l = 0, exp(-inf) = 0
    stats[(static_cast<size_t>(seq) * H + h) * 2 + 0] = mm;
    stats[(static_cast<size_t>(seq) * H + h) * 2 + 1] = t;
  }
}

// partial[seq, chunk, h, d] = sum_{s in chunk} p[s,h] x[s,d]     grid (nchunk, num_seq); each thread owns 4 consecutive
// feature columns (one 8-byte load per token) and all heads: the probabilities of the chunk sit in shared memory.
template <int HG>   // head groups of 4 actually accumulated: ceil(H / 4)
__global__ void __launch_bounds__(256) pool_accum_kernel(const bf16* __restrict__ x, const float* __restrict__ scores,
                                                         const float* __restrict__ stats, float* __restrict__ partial, int S, int D,
                                                         int H, int nchunk) {
  __shared__ __align__(16) float s_p[kChunk][kMaxHeads];
  const int chunk = blockIdx.x, seq = blockIdx.y;
  const int s0 = chunk * kChunk;
  const int ns = min(kChunk, S - s0);
  for (int i = threadIdx.x; i < ns * kMaxHeads; i += blockDim.x) {
    const int s = i / kMaxHeads, h = i % kMaxHeads;
    float p = 0.f;
    if (h < H) {
      const float m = stats[(static_cast<size_t>(seq) * H + h) * 2 + 0];
      const float sum = stats[(static_cast<size_t>(seq) * H + h) * 2 + 1];
      const float* sr = scores + (static_cast<size_t>(seq) * S + s0 + s) * kScoreLd;
      p = __expf(sr[h] + sr[H + h] - m) / sum;
    }
    s_p[s][h] = p;
  }
  __syncthreads();
  for (int d = threadIdx.x * 4; d < D; d += blockDim.x * 4) {   // D % 8 == 0
    float acc[HG * 4][4];
#pragma unroll
    for (int h = 0; h < HG * 4; ++h) acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.f;
    const bf16* xp = x + (static_cast<size_t>(seq) * S + s0) * D + d;
#pragma unroll 4
    for (int s = 0; s < ns; ++s) {
      const uint2 u = *reinterpret_cast<const uint2*>(xp + static_cast<size_t>(s) * D);
      const float x0 = bf16_lo(u.x), x1 = bf16_hi(u.x), x2 = bf16_lo(u.y), x3 = bf16_hi(u.y);
#pragma unroll
      for (int h4 = 0; h4 < HG; ++h4) {
        const float4 p = *reinterpret_cast<const float4*>(&s_p[s][h4 * 4]);
        const float pv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float* a = acc[h4 * 4 + k];
          a[0] = fmaf(pv[k], x0, a[0]); a[1] = fmaf(pv[k], x1, a[1]); a[2] = fmaf(pv[k], x2, a[2]); a[3] = fmaf(pv[k], x3, a[3]);
        }
      }
    }
    float* pp = partial + (static_cast<size_t>(seq) * nchunk + chunk) * H * D + d;
#pragma unroll
    for (int h = 0; h < HG * 4; ++h)
      if (h < H) *reinterpret_cast<float4*>(pp + static_cast<size_t>(h) * D) = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
  }
}

// xbar[seq, h, d] = sum_chunks partial[seq, chunk, h, d]   (chunk order: deterministic)
__global__ void __launch_bounds__(256) pool_reduce_kernel(const float* __restrict__ partial, float* __restrict__ xbar, int HD_in,
                                                          int nchunk, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // over num_seq * H * D
  if (i >= total) return;
  const size_t seq = i / HD_in, r = i % HD_in;
  const float* pp = partial + seq * nchunk * HD_in + r;
  float a = 0.f;
  for (int c = 0; c < nchunk; ++c) a += pp[static_cast<size_t>(c) * HD_in];
  xbar[i] = a;
}

// ctx[seq, h*dh + j] = xbar[seq, h, :] . Wv[:, h, j] + bv[h, j]
// grid (H * dh / blockDim, ceil(num_seq / kSeqTile)); a block owns blockDim columns of ONE head (blockDim = min(dh, 128))
// and kSeqTile sequences, so every weight element is read once per tile of sequences instead of once per sequence.
constexpr int kSeqTile = 8;
constexpr int kDTile = 64;
__global__ void __launch_bounds__(128) pool_ctx_kernel(const float* __restrict__ xbar, const bf16* __restrict__ wv /*[D, H*dh]*/,
                                                       const float* __restrict__ bv, float* __restrict__ ctx, int D, int H, int dh,
                                                       int num_seq) {
  __shared__ __align__(16) float s_x[kDTile][kSeqTile];
  const int col = blockIdx.x * blockDim.x + threadIdx.x;   // = h * dh + j
  const int h = (blockIdx.x * blockDim.x) / dh;
  const int seq0 = blockIdx.y * kSeqTile;
  const int HD = H * dh;
  float acc[kSeqTile];
#pragma unroll
  for (int q = 0; q < kSeqTile; ++q) acc[q] = 0.f;
  for (int d0 = 0; d0 < D; d0 += kDTile) {
    __syncthreads();
    for (int i = threadIdx.x; i < kDTile * kSeqTile; i += blockDim.x) {
      const int q = i / kDTile, dd = i % kDTile;   // consecutive threads read consecutive d: coalesced
      const int seq = seq0 + q, d = d0 + dd;
      s_x[dd][q] = (seq < num_seq && d < D) ? xbar[(static_cast<size_t>(seq) * H + h) * D + d] : 0.f;
    }
    __syncthreads();
    const int dmax = min(kDTile, D - d0);
    const bf16* wp = wv + static_cast<size_t>(d0) * HD + col;
#pragma unroll 16
    for (int dd = 0; dd < dmax; ++dd) {
      const float w = __bfloat162float(wp[static_cast<size_t>(dd) * HD]);
      const float4 a = *reinterpret_cast<const float4*>(&s_x[dd][0]);
      const float4 b = *reinterpret_cast<const float4*>(&s_x[dd][4]);
      acc[0] = fmaf(a.x, w, acc[0]); acc[1] = fmaf(a.y, w, acc[1]); acc[2] = fmaf(a.z, w, acc[2]); acc[3] = fmaf(a.w, w, acc[3]);
      acc[4] = fmaf(b.x, w, acc[4]); acc[5] = fmaf(b.y, w, acc[5]); acc[6] = fmaf(b.z, w, acc[6]); acc[7] = fmaf(b.w, w, acc[7]);
    }
  }
  const float bias = bv[col];
#pragma unroll
  for (int q = 0; q < kSeqTile; ++q)
    if (seq0 + q < num_seq) ctx[static_cast<size_t>(seq0 + q) * HD + col] = acc[q] + bias;
}

// y[seq, d] = ctx[seq, :] . wpost[d, :] + bpost[d]     grid (ceil(D / 8), ceil(num_seq / kSeqTile)), one warp per output
// feature d and kSeqTile sequences: the weight row is read once per tile, the ctx rows come from L1 / L2.
__global__ void __launch_bounds__(256) pool_post_kernel(const float* __restrict__ ctx, const bf16* __restrict__ wpost /*[D, HD]*/,
                                                        const float* __restrict__ bpost, float* __restrict__ y, int D, int HD,
                                                        int num_seq) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = blockIdx.x * 8 + warp;
  const int seq0 = blockIdx.y * kSeqTile;
  if (d >= D) return;
  float acc[kSeqTile];
#pragma unroll
  for (int q = 0; q < kSeqTile; ++q) acc[q] = 0.f;
  const bf16* wr = wpost + static_cast<size_t>(d) * HD;
  for (int i = lane * 8; i < HD; i += 256) {   // HD % 8 == 0
    const uint4 u = *reinterpret_cast<const uint4*>(wr + i);
    const float w[8] = {bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y), bf16_lo(u.z), bf16_hi(u.z), bf16_lo(u.w), bf16_hi(u.w)};
#pragma unroll
    for (int q = 0; q < kSeqTile; ++q) {
      if (seq0 + q < num_seq) {
        const float* cp = ctx + static_cast<size_t>(seq0 + q) * HD + i;
        const float4 c0 = *reinterpret_cast<const float4*>(cp);
        const float4 c1 = *reinterpret_cast<const float4*>(cp + 4);
        acc[q] += c0.x * w[0] + c0.y * w[1] + c0.z * w[2] + c0.w * w[3] + c1.x * w[4] + c1.y * w[5] + c1.z * w[6] + c1.w * w[7];
      }
    }
  }
  const float bias = bpost[d];
#pragma unroll
  for (int q = 0; q < kSeqTile; ++q) {
    const float v = warp_sum(acc[q]);
    if (lane == 0 && seq0 + q < num_seq) y[static_cast<size_t>(seq0 + q) * D + d] = v + bias;
  }
}

// out[seq, :] = [l2norm] LN( y[seq, :] )       one block per sequence
__global__ void __launch_bounds__(256) pool_norm_kernel(const float* __restrict__ y, const float* __restrict__ g1,
                                                        const float* __restrict__ beta, float* __restrict__ out, int D, int normalize) {
  extern __shared__ float sm[];  // y [D] | red [32]
  float* s_y = sm;
  float* s_red = s_y + D;
  const int seq = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int d = threadIdx.x; d < D; d += blockDim.x) s_y[d] = y[static_cast<size_t>(seq) * D + d];
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
    const float v = (s_y[d] - mean) * rstd * g1[d] + beta[d];
    s_y[d] = v;
    nn += v * v;
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

namespace {
inline size_t align4(size_t n) { return (n + 3) & ~static_cast<size_t>(3); }   // sub-buffers start on 16-byte boundaries
struct PoolScratch {
  size_t scores, stats, partial, xbar, ctx, y, end;
  PoolScratch(int num_seq, int S, int D, int H, int dh) {
    const size_t n = static_cast<size_t>(num_seq), nchunk = (S + kChunk - 1) / kChunk;
    scores = 0;
    stats = align4(scores + n * S * kScoreLd);
    partial = align4(stats + n * H * 2);
    xbar = align4(partial + n * nchunk * H * D);
    ctx = align4(xbar + n * H * D);
    y = align4(ctx + n * H * dh);
    end = align4(y + n * D);
  }
};
}  // namespace

size_t pool_scratch_floats(int num_seq, int S, int D, int H, int dh) { return PoolScratch(num_seq, S, D, H, dh).end + 64; }

cudaError_t launch_pool(cudaStream_t s, const bf16* x, int num_seq, int S, int D, int H, int dh, const bf16* wkq, const bf16* wv,
                        const float* bv, const bf16* wpost, const float* bpost, const float* ln_g1, const float* ln_b, int normalize,
                        float* scratch, float* out, int64_t* launches) {
  const int HD = H * dh;
  if (H > kMaxHeads || 2 * H > kScoreLd || (D % 8) || (dh % 8) || dh <= 0 || dh > 1024 || num_seq <= 0 || S <= 0) return cudaErrorInvalidValue;
  // pool_ctx_kernel: a block owns ctx_threads columns of ONE head, so ctx_threads must divide dh: the largest multiple of 8
  // that does, at most 128 (128 for dh = 256, 64 for 64, 88 for the giant configuration's 4 * 1408 / 16 = 352)
  int ctx_threads = (dh < 128 ? dh : 128) & ~7;
  while (ctx_threads > 8 && (dh % ctx_threads)) ctx_threads -= 8;
  if (reinterpret_cast<uintptr_t>(scratch) & 15) return cudaErrorInvalidValue;
  const int nchunk = (S + kChunk - 1) / kChunk;
  const size_t n = static_cast<size_t>(num_seq);
  const PoolScratch off(num_seq, S, D, H, dh);
  float* scores = scratch + off.scores;
  float* stats = scratch + off.stats;
  float* partial = scratch + off.partial;
  float* xbar = scratch + off.xbar;
  float* ctx = scratch + off.ctx;
  float* y = scratch + off.y;
  cudaError_t e;
  {
    GemmEpilogue ep;
    ep.out_f32 = 1;
    if ((e = launch_gemm(s, x, D, wkq, D, scores, kScoreLd, num_seq * S, kScoreLd, D, ep)) != cudaSuccess) return e;
  }
  pool_stats_kernel<<<num_seq, 1024, 0, s>>>(scores, stats, S, H);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  const int accum_threads = D / 4 >= 256 ? 256 : ((D / 4 + 31) / 32) * 32;
  switch ((H + 3) / 4) {
    case 1: pool_accum_kernel<1><<<dim3(nchunk, num_seq), accum_threads, 0, s>>>(x, scores, stats, partial, S, D, H, nchunk); break;
    case 2: pool_accum_kernel<2><<<dim3(nchunk, num_seq), accum_threads, 0, s>>>(x, scores, stats, partial, S, D, H, nchunk); break;
    case 3: pool_accum_kernel<3><<<dim3(nchunk, num_seq), accum_threads, 0, s>>>(x, scores, stats, partial, S, D, H, nchunk); break;
    default: pool_accum_kernel<4><<<dim3(nchunk, num_seq), accum_threads, 0, s>>>(x, scores, stats, partial, S, D, H, nchunk); break;
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  const size_t total = n * H * D;
  pool_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(partial, xbar, H * D, nchunk, total);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  const int seq_tiles = (num_seq + kSeqTile - 1) / kSeqTile;
  pool_ctx_kernel<<<dim3(HD / ctx_threads, seq_tiles), ctx_threads, 0, s>>>(xbar, wv, bv, ctx, D, H, dh, num_seq);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  pool_post_kernel<<<dim3((D + 7) / 8, seq_tiles), 256, 0, s>>>(ctx, wpost, bpost, y, D, HD, num_seq);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  pool_norm_kernel<<<num_seq, 256, (static_cast<size_t>(D) + 32) * sizeof(float), s>>>(y, ln_g1, ln_b, out, D, normalize);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (launches) *launches += 7;
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
