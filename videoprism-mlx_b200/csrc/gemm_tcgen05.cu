// bf16 GEMM for sm_100a: C[M,N] = A[M,K] * Wt[N,K]^T with a fused epilogue.
//
// Persistent, warp-specialised:
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B-swizzled K-major tiles)
//   warp 1      : MMA issuer    (one elected thread, tcgen05.mma kind::f16, M=128 x N=BN x K=16)
//   warp 2      : TMEM allocator
//   warps 4..11 : epilogue      (tcgen05.ld -> bias / GELU / row-scale / pos-emb / residual -> global)
// The fp32 accumulator lives in TMEM and is double buffered (2 x BN columns), so the epilogue
// of tile i overlaps the main loop of tile i+1.  The smem ring has kStages slots of
// (128 x 64 A, BN x 64 B) bf16.
//
// Replaces the reference's nn.Dense / einsum projections (layers.py:304-312, :486-488, :483-498).
#include <cuda.h>
#include <stdio.h>

#include "kernels.h"
#include "ptx.cuh"

namespace vp {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kNumThreads = 384;
constexpr int kFirstEpiWarp = 4;
constexpr int kNumEpiWarps = 8;

template <int BN>
struct Cfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // 512 or 256: power of two
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct KParams {
  int M, N, K;
  void* C;
  int ldc;
  const float* bias;
  const float* row_scale;
  const float* pos_table;
  int pos_period;
  const bf16* resid;
  int ldr;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN, int ACT, bool OUT_F32>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const KParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::kStages * C::kStageBytes;
  // barrier layout (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem ptr (4 B)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * C::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m_tiles = (p.M + BM - 1) / BM;
  const int num_n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kNumEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_addr, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n_tiles) * BM;
        const int n0 = (tile % num_n_tiles) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint32_t sb = sa + C::kABytes;
          mbar_expect_tx(full_bar(stage), C::kStageBytes);
          tma_load_2d(sa, &tmA, full_bar(stage), kb * BK, m0);
          tma_load_2d(sb, &tmB, full_bar(stage), kb * BK, n0);
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint32_t sb = sa + C::kABytes;
          const uint64_t da = umma_desc_kmajor_sw128(sa);
          const uint64_t db = umma_desc_kmajor_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row (start address is in 16 B units)
            umma_bf16_ss(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ---------------------------------------------------------------- epilogue
    const int e = warp - kFirstEpiWarp;
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int half = e >> 2;         // which half of the BN columns
    constexpr int kColsPerWarp = BN / 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / num_n_tiles) * BM;
      const int n0 = (tile % num_n_tiles) * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < p.M;
      const float rscale = (p.row_scale != nullptr && row_ok) ? __ldg(p.row_scale + m) : 1.0f;
      const float* pos_row = nullptr;
      if (p.pos_table != nullptr) pos_row = p.pos_table + static_cast<size_t>(m % p.pos_period) * p.N;
      const bf16* res_row = (p.resid != nullptr) ? p.resid + static_cast<size_t>(m) * p.ldr : nullptr;
#pragma unroll 1
      for (int ch = 0; ch < kColsPerWarp / 32; ++ch) {
        const int col = half * kColsPerWarp + ch * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + acc * BN + col + (static_cast<uint32_t>(q * 32) << 16), r);
        tmem_ld_wait();
        if (ch == kColsPerWarp / 32 - 1) {
          // all TMEM reads of this warp for this tile are done: release the accumulator early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
        const int n = n0 + col;
        if (row_ok && n < p.N) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int ng = n + g * 8;
            if (ng < p.N) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
              if (p.bias != nullptr) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ng));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + ng + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              if (ACT == ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
              } else if (ACT == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
              }
              if (p.row_scale != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] *= rscale;
              }
              if (pos_row != nullptr) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(pos_row + ng));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(pos_row + ng + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              if (res_row != nullptr) {
                const uint4 rr = *reinterpret_cast<const uint4*>(res_row + ng);
                v[0] += bf16_lo(rr.x); v[1] += bf16_hi(rr.x); v[2] += bf16_lo(rr.y); v[3] += bf16_hi(rr.y);
                v[4] += bf16_lo(rr.z); v[5] += bf16_hi(rr.z); v[6] += bf16_lo(rr.w); v[7] += bf16_hi(rr.w);
              }
              if (OUT_F32) {
                float* cp = reinterpret_cast<float*>(p.C) + static_cast<size_t>(m) * p.ldc + ng;
                *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(cp + 4) = make_float4(v[4], v[5], v[6], v[7]);
              } else {
                bf16* cp = reinterpret_cast<bf16*>(p.C) + static_cast<size_t>(m) * p.ldc + ng;
                uint4 o;
                o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
                o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
                *reinterpret_cast<uint4*>(cp) = o;
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
  }
  return fn;
}

}  // namespace

// 2-D bf16 tensor map: inner dim `cols` (contiguous), outer dim `rows` with row pitch `ld` elements,
// box = box_cols x box_rows, 128B swizzle, zero fill out of bounds.
bool make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                       uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(bf16)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int ACT, bool OUT_F32>
static cudaError_t launch_gemm_t(cudaStream_t s, const CUtensorMap& ta, const CUtensorMap& tb, const KParams& kp, int grid) {
  auto kern = gemm_bf16_kernel<BN, ACT, OUT_F32>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  kern<<<grid, kNumThreads, Cfg<BN>::kSmemBytes, s>>>(ta, tb, kp);
  return cudaGetLastError();
}

template <int BN>
static cudaError_t launch_gemm_bn(cudaStream_t s, const CUtensorMap& ta, const CUtensorMap& tb, const KParams& kp, int grid,
                                  int act, int out_f32) {
  if (out_f32) {
    if (act == ACT_NONE) return launch_gemm_t<BN, ACT_NONE, true>(s, ta, tb, kp, grid);
    return cudaErrorInvalidValue;
  }
  switch (act) {
    case ACT_NONE: return launch_gemm_t<BN, ACT_NONE, false>(s, ta, tb, kp, grid);
    case ACT_GELU: return launch_gemm_t<BN, ACT_GELU, false>(s, ta, tb, kp, grid);
    case ACT_RELU: return launch_gemm_t<BN, ACT_RELU, false>(s, ta, tb, kp, grid);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_gemm(cudaStream_t s, const bf16* A, int lda, const bf16* Wt, int ldb, void* Cout, int ldc, int M, int N,
                        int K, const GemmEpilogue& epi) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaErrorInvalidValue;
  if ((K % 8) || (N % 8) || (lda % 8) || (ldb % 8) || (ldc % 8)) return cudaErrorInvalidValue;
  if (epi.resid != nullptr && (epi.ldr % 8)) return cudaErrorInvalidValue;
  const int BN = (N % 256 == 0) ? 256 : 128;
  CUtensorMap ta, tb;
  if (!make_tmap_2d_bf16(&ta, A, M, K, lda, BM, BK)) return cudaErrorUnknown;
  if (!make_tmap_2d_bf16(&tb, Wt, N, K, ldb, BN, BK)) return cudaErrorUnknown;
  KParams kp;
  kp.M = M; kp.N = N; kp.K = K;
  kp.C = Cout; kp.ldc = ldc;
  kp.bias = epi.bias;
  kp.row_scale = epi.row_scale;
  kp.pos_table = epi.pos_table;
  kp.pos_period = epi.pos_period > 0 ? epi.pos_period : 1;
  kp.resid = epi.resid; kp.ldr = epi.ldr;
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (BN == 256) return launch_gemm_bn<256>(s, ta, tb, kp, grid, epi.act, epi.out_f32);
  return launch_gemm_bn<128>(s, ta, tb, kp, grid, epi.act, epi.out_f32);
}

}  // namespace vp
