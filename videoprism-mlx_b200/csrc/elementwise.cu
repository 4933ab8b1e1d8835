// Bandwidth-bound kernels: LayerNorm, patchify+cast, weight repack, l2-normalise, text embedding.
// All use 128-bit accesses where the layout allows and warp-shuffle reductions.
#include "kernels.h"
#include "ptx.cuh"

namespace vp {

int num_sms();

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ LayerNorm
// Persistent warps: each warp walks rows with a grid stride, holding one row (D bf16, D % 8 == 0,
// D <= 8 * 32 * VPL) in registers as VPL 16-byte vectors per lane, and prefetches its next row before reducing the
// current one; gamma / beta stay in registers.  Two-pass statistics in fp32 (mean, then biased variance of the centred values) as
// layers.py:240-242 does.  Streaming loads / stores bypass L1.
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <int VPL>
__global__ void __launch_bounds__(256) layernorm_kernel(const LnArgs a) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = a.D >> 3;
  const int stride = gridDim.x * warps_per_block;
  int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (row >= a.M) return;
  const float inv_d = 1.0f / static_cast<float>(a.D);
  // gamma / beta of this lane's columns stay in registers for all rows
  float4 g0[VPL], g1[VPL], b0[VPL], b1[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const int c = vi * 8;
      g0[i] = __ldg(reinterpret_cast<const float4*>(a.gamma1 + c)); g1[i] = __ldg(reinterpret_cast<const float4*>(a.gamma1 + c + 4));
      b0[i] = __ldg(reinterpret_cast<const float4*>(a.beta + c)); b1[i] = __ldg(reinterpret_cast<const float4*>(a.beta + c + 4));
    }
  }
  uint4 nxt[VPL];
  {
    const uint4* xr = reinterpret_cast<const uint4*>(a.x + static_cast<size_t>(row) * a.ldx);
#pragma unroll
    for (int i = 0; i < VPL; ++i)
      if (lane + i * 32 < nvec) nxt[i] = ld_stream(xr + lane + i * 32);
  }
  for (; row < a.M; row += stride) {
    float v[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + i * 32 < nvec) {
        const uint4 u = nxt[i];
        v[i][0] = bf16_lo(u.x); v[i][1] = bf16_hi(u.x); v[i][2] = bf16_lo(u.y); v[i][3] = bf16_hi(u.y);
        v[i][4] = bf16_lo(u.z); v[i][5] = bf16_hi(u.z); v[i][6] = bf16_lo(u.w); v[i][7] = bf16_hi(u.w);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
      }
    }
    // prefetch the next row of this warp while this one is reduced and written
    if (row + stride < a.M) {
      const uint4* xr = reinterpret_cast<const uint4*>(a.x + static_cast<size_t>(row + stride) * a.ldx);
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (lane + i * 32 < nvec) nxt[i] = ld_stream(xr + lane + i * 32);
    }
    const float mean = warp_sum(s) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (lane + i * 32 < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[i][j] - mean;
          sq += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + 1e-6f);
    const float* addrow = nullptr;
    if (a.add_table != nullptr) addrow = a.add_table + static_cast<size_t>((row / a.add_div) % a.add_mod) * a.D;
    float os = 0.f, oq = 0.f;   // statistics of the stored bf16 row (for the next LayerNorm-folded GEMM)
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const int c = vi * 8;
        float y[8];
        y[0] = (v[i][0] - mean) * rstd * g0[i].x + b0[i].x; y[1] = (v[i][1] - mean) * rstd * g0[i].y + b0[i].y;
        y[2] = (v[i][2] - mean) * rstd * g0[i].z + b0[i].z; y[3] = (v[i][3] - mean) * rstd * g0[i].w + b0[i].w;
        y[4] = (v[i][4] - mean) * rstd * g1[i].x + b1[i].x; y[5] = (v[i][5] - mean) * rstd * g1[i].y + b1[i].y;
        y[6] = (v[i][6] - mean) * rstd * g1[i].z + b1[i].z; y[7] = (v[i][7] - mean) * rstd * g1[i].w + b1[i].w;
        if (a.y_f32 != nullptr) {
          float* yp = a.y_f32 + static_cast<size_t>(row) * a.D + c;
          *reinterpret_cast<float4*>(yp) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(yp + 4) = make_float4(y[4], y[5], y[6], y[7]);
        }
        if (a.y_bf16 != nullptr) {
          if (addrow != nullptr) {
            const float4 t0 = __ldg(reinterpret_cast<const float4*>(addrow + c));
            const float4 t1 = __ldg(reinterpret_cast<const float4*>(addrow + c + 4));
            y[0] += t0.x; y[1] += t0.y; y[2] += t0.z; y[3] += t0.w;
            y[4] += t1.x; y[5] += t1.y; y[6] += t1.z; y[7] += t1.w;
          }
          if (a.resid != nullptr) {
            const uint4 r = *reinterpret_cast<const uint4*>(a.resid + static_cast<size_t>(row) * a.ldr + c);
            y[0] += bf16_lo(r.x); y[1] += bf16_hi(r.x); y[2] += bf16_lo(r.y); y[3] += bf16_hi(r.y);
            y[4] += bf16_lo(r.z); y[5] += bf16_hi(r.z); y[6] += bf16_lo(r.w); y[7] += bf16_hi(r.w);
          }
          uint4 o;
          o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
          o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
          *reinterpret_cast<uint4*>(a.y_bf16 + static_cast<size_t>(row) * a.D + c) = o;
          if (a.stats_out != nullptr) {
            const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float lo = bf16_lo(ow[j]), hi = bf16_hi(ow[j]);
              os += lo + hi;
              oq = fmaf(lo, lo, fmaf(hi, hi, oq));
            }
          }
        }
      }
    }
    if (a.stats_out != nullptr) {
      os = warp_sum(os); oq = warp_sum(oq);
      if (lane == 0) { a.stats_out[2 * static_cast<size_t>(row)] = os; a.stats_out[2 * static_cast<size_t>(row) + 1] = oq; }
    }
  }
}

// ------------------------------------------------------------------- patchify
// One block per row of patches (bt, gy): p image rows of W*3 contiguous values each (fully coalesced pair loads);
// value pair e of image row py lands in patch gx = e / (3p) at column py*3p + (e mod 3p): a 3p-element (108-byte)
// contiguous run of the output row.  (p even => pairs never straddle patches.)
template <typename TIn>
__global__ void __launch_bounds__(448) patchify_kernel(const TIn* __restrict__ video, bf16* __restrict__ out, int ldo,
                                                       int H, int W, int p) {
  const int gh = H / p, gw = W / p;
  const int gy = blockIdx.x % gh;
  const size_t bt = blockIdx.x / gh;
  const int run = 3 * p;                 // values per (patch, py)
  const int row_vals = W * 3;
  const TIn* src0 = video + (bt * H + static_cast<size_t>(gy) * p) * row_vals;
  bf16* dst0 = out + ((bt * gh + gy) * gw) * static_cast<size_t>(ldo);
  for (int t = threadIdx.x; 2 * t < row_vals; t += blockDim.x) {
    const int e = 2 * t;
    const int gx = e / run, j = e - gx * run;
    bf16* dst = dst0 + static_cast<size_t>(gx) * ldo + j;
#pragma unroll 6
    for (int py = 0; py < p; ++py) {
      float2 v;
      if (sizeof(TIn) == 1) {
        const uchar2 u = *reinterpret_cast<const uchar2*>(src0 + static_cast<size_t>(py) * row_vals + e);
        v = make_float2(__fdiv_rn(static_cast<float>(u.x), 255.0f), __fdiv_rn(static_cast<float>(u.y), 255.0f));
      } else {
        v = *reinterpret_cast<const float2*>(src0 + static_cast<size_t>(py) * row_vals + e);
      }
      *reinterpret_cast<uint32_t*>(dst + py * run) = pack_bf16x2(v.x, v.y);
    }
  }
}

// -------------------------------------------------------------- weight repack
__global__ void transpose_cast_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int K, int N, int ldk, float scale) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? src[static_cast<size_t>(k) * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < ldk) dst[static_cast<size_t>(n) * ldk + k] = __float2bfloat16(k < K ? tile[threadIdx.x][i] * scale : 0.f);
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n, float scale) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16(src[i] * scale);
}

__global__ void affine_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n, float a, float b) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = a * src[i] + b;
}

// ------------------------------------------------------------------ row stats
__global__ void __launch_bounds__(256) row_stats_kernel(const bf16* __restrict__ x, int ldx, float* __restrict__ stats, int M, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * ldx);
  float s = 0.f, q = 0.f;
  for (int vi = lane; vi < (D >> 3); vi += 32) {
    const uint4 u = xr[vi];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lo = bf16_lo(w[j]), hi = bf16_hi(w[j]);
      s += lo + hi;
      q = fmaf(lo, lo, fmaf(hi, hi, q));
    }
  }
  s = warp_sum(s); q = warp_sum(q);
  if (lane == 0) { stats[2 * static_cast<size_t>(row)] = s; stats[2 * static_cast<size_t>(row) + 1] = q; }
}

// ------------------------------------------------------- LayerNorm weight folding (load time)
// One block per output feature n: dst[n, k] = bf16(gamma1[k] * src[k, n] * scale); colsum / bias reductions in fp32.
__global__ void __launch_bounds__(256) fold_ln_weight_kernel(const float* __restrict__ src, const float* __restrict__ gamma1,
                                                             const float* __restrict__ beta, const float* __restrict__ bias_in,
                                                             bf16* __restrict__ dst, float* __restrict__ colsum,
                                                             float* __restrict__ bias_out, int K, int N, int ldk, float scale) {
  __shared__ float red[2][8];
  const int n = blockIdx.x;
  float cs = 0.f, bs = 0.f;
  for (int k = threadIdx.x; k < ldk; k += blockDim.x) {
    float wq = 0.f;
    if (k < K) {
      const float w = src[static_cast<size_t>(k) * N + n] * scale;
      const bf16 wb = __float2bfloat16(w * gamma1[k]);
      wq = __bfloat162float(wb);
      bs += beta[k] * w;
      dst[static_cast<size_t>(n) * ldk + k] = wb;
    } else {
      dst[static_cast<size_t>(n) * ldk + k] = __float2bfloat16(0.f);
    }
    cs += wq;
  }
  cs = warp_sum(cs); bs = warp_sum(bs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = cs; red[1][warp] = bs; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float c = 0.f, b = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) { c += red[0][w]; b += red[1][w]; }
    colsum[n] = c;
    bias_out[n] = b + (bias_in != nullptr ? bias_in[n] * scale : 0.f);
  }
}

// ------------------------------------------------------------------- l2 norm
__global__ void l2norm_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int D) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<size_t>(row) * D;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += xr[c] * xr[c];
  const float inv = 1.0f / sqrtf(warp_sum(s) + 1e-12f);
  for (int c = lane; c < D; c += 32) y[static_cast<size_t>(row) * D + c] = xr[c] * inv;
}

// ------------------------------------------------------------- text embedding
__global__ void text_embed_kernel(const int32_t* __restrict__ ids, const float* __restrict__ pad, const float* __restrict__ emb,
                                  const float* __restrict__ pe, const float* __restrict__ cls, bf16* __restrict__ x,
                                  float* __restrict__ keep, float* __restrict__ pad_ext, int Q, int L, int D, int vocab) {
  const int row = blockIdx.x;  // q * (L + 1) + j
  const int q = row / (L + 1), j = row % (L + 1);
  const float sq = sqrtf(static_cast<float>(D));
  bf16* xr = x + static_cast<size_t>(row) * D;
  if (j < L) {
    int id = ids[q * L + j];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float* er = emb + static_cast<size_t>(id) * D;
    const float* pr = pe + static_cast<size_t>(j) * D;
    for (int c = threadIdx.x; c < D; c += blockDim.x) xr[c] = __float2bfloat16(er[c] * sq + pr[c]);
    if (threadIdx.x == 0) {
      const float pv = pad[q * L + j];
      keep[row] = 1.0f - pv;
      pad_ext[row] = pv;
    }
  } else {
    for (int c = threadIdx.x; c < D; c += blockDim.x) xr[c] = __float2bfloat16(cls[c] * sq);
    if (threadIdx.x == 0) {
      keep[row] = 1.0f;
      pad_ext[row] = 0.0f;
    }
  }
}

}  // namespace

cudaError_t launch_layernorm(cudaStream_t s, const LnArgs& a) {
  if (a.M <= 0) return cudaSuccess;
  if ((a.D % 8) || (a.ldx % 8) || a.D > 8 * 32 * 6 || (a.resid != nullptr && (a.ldr % 8))) return cudaErrorInvalidValue;
  const int warps = 8;
  const int full = (a.M + warps - 1) / warps;
  const int cap = num_sms() * 8;                       // persistent: up to 8 blocks of 8 warps per SM
  const int grid = full < cap ? full : cap;
  const int nvec = a.D / 8;
  const int vpl = (nvec + 31) / 32;
  switch (vpl) {
    case 1: layernorm_kernel<1><<<grid, warps * 32, 0, s>>>(a); break;
    case 2: layernorm_kernel<2><<<grid, warps * 32, 0, s>>>(a); break;
    case 3: layernorm_kernel<3><<<grid, warps * 32, 0, s>>>(a); break;
    case 4: layernorm_kernel<4><<<grid, warps * 32, 0, s>>>(a); break;
    default: layernorm_kernel<6><<<grid, warps * 32, 0, s>>>(a); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_patchify(cudaStream_t s, const float* video, bf16* out, int ldo, int BT, int H, int W, int p) {
  if ((p % 2) || (H % p) || (W % p) || (ldo % 2)) return cudaErrorInvalidValue;
  if (BT <= 0) return cudaSuccess;
  const int block = (W * 3 / 2 >= 448) ? 448 : ((W * 3 / 2 + 31) / 32) * 32;
  patchify_kernel<float><<<static_cast<unsigned>(BT) * (H / p), block, 0, s>>>(video, out, ldo, H, W, p);
  return cudaGetLastError();
}

cudaError_t launch_patchify_u8(cudaStream_t s, const uint8_t* video, bf16* out, int ldo, int BT, int H, int W, int p) {
  if ((p % 2) || (H % p) || (W % p) || (ldo % 2)) return cudaErrorInvalidValue;
  if (BT <= 0) return cudaSuccess;
  const int block = (W * 3 / 2 >= 448) ? 448 : ((W * 3 / 2 + 31) / 32) * 32;
  patchify_kernel<uint8_t><<<static_cast<unsigned>(BT) * (H / p), block, 0, s>>>(video, out, ldo, H, W, p);
  return cudaGetLastError();
}

cudaError_t launch_transpose_cast(cudaStream_t s, const float* src, bf16* dst, int K, int N, int ldk, float scale) {
  dim3 grid((ldk + 31) / 32, (N + 31) / 32), block(32, 8);
  transpose_cast_kernel<<<grid, block, 0, s>>>(src, dst, K, N, ldk, scale);
  return cudaGetLastError();
}

cudaError_t launch_cast_bf16(cudaStream_t s, const float* src, bf16* dst, size_t n, float scale) {
  if (n == 0) return cudaSuccess;
  cast_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(src, dst, n, scale);
  return cudaGetLastError();
}

cudaError_t launch_affine_f32(cudaStream_t s, const float* src, float* dst, size_t n, float a, float b) {
  if (n == 0) return cudaSuccess;
  affine_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(src, dst, n, a, b);
  return cudaGetLastError();
}

cudaError_t launch_row_stats(cudaStream_t s, const bf16* x, int ldx, float* stats, int M, int D) {
  if (M <= 0) return cudaSuccess;
  if ((D % 8) || (ldx % 8)) return cudaErrorInvalidValue;
  row_stats_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, ldx, stats, M, D);
  return cudaGetLastError();
}

cudaError_t launch_fold_ln_weight(cudaStream_t s, const float* src, const float* gamma1, const float* beta, const float* bias_in,
                                  bf16* dst, float* colsum, float* bias_out, int K, int N, int ldk, float scale) {
  fold_ln_weight_kernel<<<N, 256, 0, s>>>(src, gamma1, beta, bias_in, dst, colsum, bias_out, K, N, ldk, scale);
  return cudaGetLastError();
}

cudaError_t launch_l2norm(cudaStream_t s, const float* x, float* y, int rows, int D) {
  if (rows <= 0) return cudaSuccess;
  l2norm_kernel<<<(rows + 3) / 4, 128, 0, s>>>(x, y, rows, D);
  return cudaGetLastError();
}

cudaError_t launch_text_embed(cudaStream_t s, const int32_t* ids, const float* pad, const float* emb, const float* pe,
                              const float* cls, bf16* x, float* keep, float* pad_ext, int Q, int L, int D, int vocab) {
  if (Q <= 0) return cudaSuccess;
  text_embed_kernel<<<Q * (L + 1), 128, 0, s>>>(ids, pad, emb, pe, cls, x, keep, pad_ext, Q, L, D, vocab);
  return cudaGetLastError();
}

}  // namespace vp
