// Bandwidth-bound kernels: LayerNorm, patchify+cast, weight repack, l2-normalise, text embedding.
// All use 128-bit accesses where the layout allows and warp-shuffle reductions.
#include "kernels.h"
#include "ptx.cuh"

namespace vp {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ LayerNorm
// One warp per row; the row (D bf16, D % 8 == 0, D <= 8 * 32 * VPL) is held in registers as
// VPL 16-byte vectors per lane.  Two-pass statistics in fp32 (mean, then biased variance of the
// centred values) as layers.py:240-242 does.
template <int VPL>
__global__ void __launch_bounds__(256) layernorm_kernel(const LnArgs a) {
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= a.M) return;
  const int nvec = a.D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(a.x + static_cast<size_t>(row) * a.ldx);
  float v[VPL][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint4 u = xr[vi];
      v[i][0] = bf16_lo(u.x); v[i][1] = bf16_hi(u.x); v[i][2] = bf16_lo(u.y); v[i][3] = bf16_hi(u.y);
      v[i][4] = bf16_lo(u.z); v[i][5] = bf16_hi(u.z); v[i][6] = bf16_lo(u.w); v[i][7] = bf16_hi(u.w);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
  const float inv_d = 1.0f / static_cast<float>(a.D);
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
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const int c = vi * 8;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma1 + c));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma1 + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.beta + c + 4));
      float y[8];
      y[0] = (v[i][0] - mean) * rstd * g0.x + b0.x; y[1] = (v[i][1] - mean) * rstd * g0.y + b0.y;
      y[2] = (v[i][2] - mean) * rstd * g0.z + b0.z; y[3] = (v[i][3] - mean) * rstd * g0.w + b0.w;
      y[4] = (v[i][4] - mean) * rstd * g1.x + b1.x; y[5] = (v[i][5] - mean) * rstd * g1.y + b1.y;
      y[6] = (v[i][6] - mean) * rstd * g1.z + b1.z; y[7] = (v[i][7] - mean) * rstd * g1.w + b1.w;
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
        uint4 o;
        o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
        o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
        *reinterpret_cast<uint4*>(a.y_bf16 + static_cast<size_t>(row) * a.D + c) = o;
      }
    }
  }
}

// ------------------------------------------------------------------- patchify
// One thread per pair of adjacent output elements: the (px, c) run of a patch row is 3p
// contiguous floats in the source frame (p even => float2 / bf16x2 aligned).
template <typename TIn>
__global__ void __launch_bounds__(256) patchify_kernel(const TIn* __restrict__ video, bf16* __restrict__ out, int ldo,
                                                       int BT, int H, int W, int p, long long total_pairs) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total_pairs) return;
  const int run = 3 * p / 2;  // float2 pairs per (patch, py)
  const int gw = W / p, gh = H / p;
  const int j = static_cast<int>(idx % run);
  long long t = idx / run;
  const int py = static_cast<int>(t % p);  t /= p;
  const int gx = static_cast<int>(t % gw); t /= gw;
  const int gy = static_cast<int>(t % gh); t /= gh;
  const long long bt = t;
  const size_t src = ((static_cast<size_t>(bt) * H + (gy * p + py)) * W + gx * p) * 3 + 2 * j;
  float2 v;
  if (sizeof(TIn) == 1) {
    const uchar2 u = *reinterpret_cast<const uchar2*>(video + src);
    v = make_float2(__fdiv_rn(static_cast<float>(u.x), 255.0f), __fdiv_rn(static_cast<float>(u.y), 255.0f));
  } else {
    v = *reinterpret_cast<const float2*>(video + src);
  }
  const size_t row = (static_cast<size_t>(bt) * gh + gy) * gw + gx;
  const size_t dst = row * ldo + static_cast<size_t>(py) * 3 * p + 2 * j;
  *reinterpret_cast<uint32_t*>(out + dst) = pack_bf16x2(v.x, v.y);
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
  if ((a.D % 8) || (a.ldx % 8) || a.D > 8 * 32 * 6) return cudaErrorInvalidValue;
  const int warps = 8;
  const int grid = (a.M + warps - 1) / warps;
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
  const long long total = static_cast<long long>(BT) * (H / p) * (W / p) * p * (3 * p / 2);
  if (total == 0) return cudaSuccess;
  const int block = 256;
  const long long grid = (total + block - 1) / block;
  patchify_kernel<float><<<static_cast<unsigned>(grid), block, 0, s>>>(video, out, ldo, BT, H, W, p, total);
  return cudaGetLastError();
}

cudaError_t launch_patchify_u8(cudaStream_t s, const uint8_t* video, bf16* out, int ldo, int BT, int H, int W, int p) {
  if ((p % 2) || (H % p) || (W % p) || (ldo % 2)) return cudaErrorInvalidValue;
  const long long total = static_cast<long long>(BT) * (H / p) * (W / p) * p * (3 * p / 2);
  if (total == 0) return cudaSuccess;
  const int block = 256;
  const long long grid = (total + block - 1) / block;
  patchify_kernel<uint8_t><<<static_cast<unsigned>(grid), block, 0, s>>>(video, out, ldo, BT, H, W, p, total);
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
