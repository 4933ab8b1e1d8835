// tcgen05 / TMEM fused attention for LONG unmasked sequences (the auxiliary encoder of the video-text models:
// S = T*N = 4096 tokens per clip, dh = 64, capped logits; encoders.py:846-857, layers.py:601-661).
//
// Flash-style key loop WITHOUT the online-softmax rescale: the logit cap cap*tanh(s/cap) (layers.py:586-594) bounds
// the base-2 exponent (72 for cap = 50), so exp2 cannot overflow without a running row maximum, softmax is
// shift-invariant, and the unnormalised O = sum_j P_j V_j and l = sum_j rowsum(P_j) simply ACCUMULATE over the key
// blocks (O in TMEM through the MMA's accumulate flag, l in registers).
//
// One persistent CTA per SM walks (sequence, head, 256-row query block) problems; per problem it loops over key
// blocks of 128 keys.  Two query tiles (A, B: 128 rows each) alternate, so the tensor core works for one while the
// softmax warps exponentiate the other:
//   warp 0             : TMA producer (Q per problem, double buffered; K / V blocks through a 4-stage ring)
//   warp 1             : MMA issuer, static order  S_A(g+1) | PV_A(g) | S_B(g+1) | PV_B(g)
//   warps 2, 3, 20, 21 : drain warps (one per TMEM lane quarter): O / l -> bf16 -> (dead) Q buffer -> TMA store
//   warps 4..19        : softmax; warp (q, c) owns rows [32q, 32q+32) x 32 of a block's 128 key columns of both tiles
// TMEM per query tile (256 columns): S block [0,128) fp32 | P [128,192) bf16 pairs | O [192,256) fp32.  P has its
// own columns, so S(g+1) only waits for the softmax warps to have READ S(g), not for PV(g).
#include <cuda.h>
#include <math_constants.h>
#include <stdlib.h>

#include "kernels.h"
#include "ptx.cuh"

namespace vp {

bool make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                       uint32_t box_cols, int swizzle_bytes);
int num_sms();
bool pdl_enabled();

namespace {

constexpr int kQBytes = 256 * 64 * 2;           // Q of a problem (both query tiles): 32 KB
constexpr int kKVBlockBytes = 128 * 64 * 2;     // one K or V block: 16 KB
constexpr int kKVStageBytes = 2 * kKVBlockBytes;
constexpr int kKVStages = 4;
constexpr int kSoftmaxWarps = 16;
constexpr int kThreads = 32 * (4 + kSoftmaxWarps + 2);   // 704
constexpr int kXsumBytes = 2 * 4 * 128 * 4;     // row-sum partials [tile][slice][row]
constexpr int kSmemBytes = 2 * kQBytes + kKVStages * kKVStageBytes + kXsumBytes + 1024 /*align slack*/ + 512 /*barriers*/;
constexpr float kLog2e = 1.4426950408889634f;

struct LongParams {
  int num_problems, heads, D, S;   // problems = num_seq * heads * (S / 256)
  float b0, b1, b2;                // cap*log2e*tanh(s/cap) ~= s*(b0 + b1 s^2 + b2 s^4) for |s| <= range
  float range;
  float cap_l2, inv_cap;
};

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
attn_long_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                         const __grid_constant__ CUtensorMap tmO, const LongParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t kv_base = smem_base + 2 * kQBytes;
  const uint32_t xsum_base = kv_base + kKVStages * kKVStageBytes;
  const uint32_t bar_base = xsum_base + kXsumBytes;
  auto q_full = [&](int b) { return bar_base + 8u * b; };
  auto q_empty = [&](int b) { return bar_base + 8u * (2 + b); };
  auto kv_full = [&](int s) { return bar_base + 8u * (4 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (4 + kKVStages + s); };
  constexpr int kB0 = 4 + 2 * kKVStages;
  auto s_full = [&](int t) { return bar_base + 8u * (kB0 + t); };
  auto s_free = [&](int t) { return bar_base + 8u * (kB0 + 2 + t); };
  auto p_full = [&](int t) { return bar_base + 8u * (kB0 + 4 + t); };
  auto p_free = [&](int t) { return bar_base + 8u * (kB0 + 6 + t); };
  auto o_full = [&](int t) { return bar_base + 8u * (kB0 + 8 + t); };
  auto o_free = [&](int t) { return bar_base + 8u * (kB0 + 10 + t); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (kB0 + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_it = (p.num_problems - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int nkb = p.S / 128;              // key blocks per problem
  const int qblocks = p.S / 256;          // query blocks per (sequence, head)
  const int G = n_it * nkb;               // key blocks this CTA walks in total

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(q_full(b), 1);
      mbar_init(q_empty(b), 2);   // one arrive per stored query tile
    }
    for (int s = 0; s < kKVStages; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(s_full(t), 1);
      mbar_init(s_free(t), kSoftmaxWarps);
      mbar_init(p_full(t), kSoftmaxWarps);
      mbar_init(p_free(t), 1);
      mbar_init(o_full(t), 1);
      mbar_init(o_free(t), 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_addr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  pdl_launch_dependents();
  pdl_wait();

  // problem index -> (row of the sequence's first token, query block, head)
  auto decode = [&](int it, int& row0, int& qb, int& h) {
    const int pr = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
    h = pr % p.heads;
    const int r = pr / p.heads;
    qb = r % qblocks;
    row0 = (r / qblocks) * p.S;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int g = 0;
      for (int it = 0; it < n_it; ++it) {
        int row0, qb, h;
        decode(it, row0, qb, h);
        const int qbuf = it & 1;
        mbar_wait(q_empty(qbuf), ((it >> 1) & 1u) ^ 1u);
        mbar_expect_tx(q_full(qbuf), kQBytes);
        tma_load_2d(smem_base + qbuf * kQBytes, &tmQ, q_full(qbuf), h * 64, row0 + qb * 256);
        for (int j = 0; j < nkb; ++j, ++g) {
          const int stage = g % kKVStages;
          mbar_wait(kv_empty(stage), ((g / kKVStages) & 1u) ^ 1u);
          const uint32_t sk = kv_base + stage * kKVStageBytes;
          mbar_expect_tx(kv_full(stage), kKVStageBytes);
          tma_load_2d(sk, &tmKV, kv_full(stage), p.D + h * 64, row0 + j * 128);
          tma_load_2d(sk + kKVBlockBytes, &tmKV, kv_full(stage), 2 * p.D + h * 64, row0 + j * 128);
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer (static order, blocking waits)
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);   // B = V is MN-major (dh contiguous per key)
      auto issue_s = [&](int t, int g) {
        const int it = g / nkb, j = g - it * nkb;
        const int stage = g % kKVStages;
        if (t == 0) {
          if (j == 0) mbar_wait(q_full(it & 1), (it >> 1) & 1u);
          mbar_wait(kv_full(stage), (g / kKVStages) & 1u);
        }
        if (g > 0) mbar_wait(s_free(t), (g - 1) & 1u);   // the softmax warps have read S(g-1)
        tc_fence_after();
        const uint32_t T = tmem_base + t * 256;
        const uint64_t dq = umma_desc_kmajor_sw128(smem_base + (it & 1) * kQBytes + t * (kQBytes / 2));
        const uint64_t dk = umma_desc_kmajor_sw128(kv_base + stage * kKVStageBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(T, dq + 2u * k, dk + 2u * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full(t));
      };
      auto issue_pv = [&](int t, int g) {
        const int it = g / nkb, j = g - it * nkb;
        const int stage = g % kKVStages;
        mbar_wait(p_full(t), g & 1u);
        if (j == 0 && it > 0) mbar_wait(o_free(t), (it - 1) & 1u);   // the previous problem's O has been read out
        tc_fence_after();
        const uint32_t T = tmem_base + t * 256;
        // V block: key k at byte k*128 (64 dh values), 8-key swizzle atoms of 1024 B; one K=16 step = 2 atoms
        const uint64_t dv = umma_desc_mnmajor_sw128(kv_base + stage * kKVStageBytes + kKVBlockBytes, 1024, 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_bf16_ts(T + 192, T + 128 + 8 * k, dv + static_cast<uint64_t>(k) * (2048 >> 4), idesc_pv, (j | k) != 0 ? 1u : 0u);
        umma_commit(p_free(t));
        if (j == nkb - 1) umma_commit(o_full(t));
      };
      if (G > 0) {
        issue_s(0, 0);
        issue_s(1, 0);
      }
      for (int g = 0; g < G; ++g) {
        if (g + 1 < G) issue_s(0, g + 1);
        issue_pv(0, g);
        if (g + 1 < G) issue_s(1, g + 1);
        issue_pv(1, g);
        umma_commit(kv_empty(g % kKVStages));   // every MMA that reads this K / V block has been issued
      }
    }
  } else if (warp == 2 || warp == 3 || warp >= 4 + kSoftmaxWarps) {
    // -------------------------------------------------------------- drain warps, one per TMEM lane quarter
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool leader = (warp == 3) && elect_one();
    for (int it = 0; it < n_it; ++it) {
      int row0, qb, h;
      decode(it, row0, qb, h);
      const int qbuf = it & 1;
#pragma unroll 1
      for (int tile = 0; tile < 2; ++tile) {
        const uint32_t T = tmem_base + tile * 256 + lane_off;
        const uint32_t so = smem_base + qbuf * kQBytes + tile * (kQBytes / 2);   // Q of this tile is dead: all its S MMAs are done
        const uint32_t rowaddr = so + row * 128;
        const int sw = row & 7;
        mbar_wait(o_full(tile), it & 1u);
        tc_fence_after();
        float l = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(xsum_base + ((tile * 4 + c) * 128 + row) * 4));
          l += v;
        }
        const float inv = 1.0f / l;
        const f32x2 inv2 = pk2(inv, inv);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(T + 192 + 32 * half, o);
          tmem_ld_wait();
          if (half == 1) {
            // O and the row sums of this tile have been read: the next problem may overwrite them
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_free(tile));
          }
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            uint32_t wv[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              float a, b;
              upk2(mul2(pk2u(o[g4 * 8 + jj * 2], o[g4 * 8 + jj * 2 + 1]), inv2), a, b);
              wv[jj] = pack_bf16x2(a, b);
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (((4 * half + g4) ^ sw) << 4)), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]), "r"(wv[3]) : "memory");
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);   // the four drain warps: the tile is complete
        if (leader) {
          tma_store_2d(&tmO, so, h * 64, row0 + qb * 256 + tile * 128);
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(q_empty(qbuf));
        }
      }
    }
    if (leader) tma_store_wait<0>();
  } else if (warp >= 4 && warp < 4 + kSoftmaxWarps) {
    // -------------------------------------------------------------- softmax warps: (lane quarter q, key slice c)
    const int q = warp & 3;
    const int c = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const f32x2 B0 = pk2(p.b0, p.b0), B1 = pk2(p.b1, p.b1), B2 = pk2(p.b2, p.b2);
    f32x2 sum2[2] = {pk2(0.f, 0.f), pk2(0.f, 0.f)};
    int g = 0;
    for (int it = 0; it < n_it; ++it) {
      for (int j = 0; j < nkb; ++j, ++g) {
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          const uint32_t T = tmem_base + tile * 256 + lane_off;
          mbar_wait(s_full(tile), g & 1u);
          tc_fence_after();
          uint32_t r[32];
          tmem_ld_32x32b_x32(T + 32 * c, r);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free(tile));   // S(g) is in registers: S(g+1) may overwrite it
          float am[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 16; ++i) am[i & 3] = max3(am[i & 3], fabsf(__uint_as_float(r[2 * i])), fabsf(__uint_as_float(r[2 * i + 1])));
          const float amax = fmaxf(max3(am[0], am[1], am[2]), am[3]);
          uint32_t w[16];
          f32x2 acc = sum2[tile];
          if (amax <= p.range) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const f32x2 v = pk2u(r[2 * i], r[2 * i + 1]);
              const f32x2 u = mul2(v, v);
              f32x2 t = fma2(u, B2, B1);
              t = fma2(t, u, B0);
              float a, b;
              upk2(mul2(t, v), a, b);
              // P is the exponential TRUNCATED to bf16; the row sum is taken over the truncated values, so the weights the
              // tensor core multiplies with V sum to exactly the normaliser
              const uint32_t e0 = __float_as_uint(ex2_approx(a)) & 0xFFFF0000u, e1 = __float_as_uint(ex2_approx(b)) & 0xFFFF0000u;
              acc = add2(acc, pk2u(e0, e1));
              w[i] = __byte_perm(e0, e1, 0x7632);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a = p.cap_l2 * tanh_approx(__uint_as_float(r[2 * i]) * p.inv_cap);
              const float b = p.cap_l2 * tanh_approx(__uint_as_float(r[2 * i + 1]) * p.inv_cap);
              const uint32_t e0 = __float_as_uint(ex2_approx(a)) & 0xFFFF0000u, e1 = __float_as_uint(ex2_approx(b)) & 0xFFFF0000u;
              acc = add2(acc, pk2u(e0, e1));
              w[i] = __byte_perm(e0, e1, 0x7632);
            }
          }
          sum2[tile] = acc;
          if (g > 0) {
            mbar_wait(p_free(tile), (g - 1) & 1u);   // PV(g-1) has consumed the previous P
            tc_fence_after();
          }
          tmem_st_32x32b_x16(T + 128 + 16 * c, w);
          if (j == nkb - 1) {
            // last key block of the problem: publish this slice's share of the row sum (read by the drain warps after o_full)
            float s0, s1;
            upk2(sum2[tile], s0, s1);
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(xsum_base + ((tile * 4 + c) * 128 + row) * 4), "f"(s0 + s1) : "memory");
            sum2[tile] = pk2(0.f, 0.f);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full(tile));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Returns cudaErrorNotSupported when the problem does not fit this kernel (the caller falls back to the mma.sync
// flash kernel of attention.cu).
cudaError_t launch_attention_long_tcgen05(cudaStream_t s, const AttnArgs& a) {
  const int D = a.heads * a.dh;
  // (S = 256 is faster on the whole-row kernel of attention_tcgen05.cu: 183 vs 205 us at B = 32)
  if (a.S < 512 || (a.S % 256) || a.dh != 64 || a.group != 1 || a.key_pad != nullptr || a.causal) return cudaErrorNotSupported;
  // no row maximum is taken: the logit cap must bound the exponent (cap * log2e * ... < 100 keeps exp2 and the sums finite)
  if (!(a.cap > 0.f) || a.cap * kLog2e >= 100.0f) return cudaErrorNotSupported;
  if (a.k != a.q + D || a.v != a.q + 2 * D || (a.ld % 8) || (a.ldo % 8)) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(a.q) & 15) || (reinterpret_cast<uintptr_t>(a.out) & 15)) return cudaErrorNotSupported;
  const uint64_t rows = static_cast<uint64_t>(a.num_seq) * a.S;
  CUtensorMap tq, tkv, to;
  if (!make_tmap_2d_bf16(&tq, a.q, rows, 3 * D, a.ld, 256, 64, 128)) return cudaErrorUnknown;
  if (!make_tmap_2d_bf16(&tkv, a.q, rows, 3 * D, a.ld, 128, 64, 128)) return cudaErrorUnknown;
  if (!make_tmap_2d_bf16(&to, a.out, rows, D, a.ldo, 128, 64, 128)) return cudaErrorUnknown;
  LongParams p;
  p.num_problems = a.num_seq * a.heads * (a.S / 256);
  p.heads = a.heads;
  p.D = D;
  p.S = a.S;
  // same cap polynomial as attention_tcgen05.cu: tanh(x)/x = 1 + t1 x^2 + t2 x^4 on |x| <= 0.5
  const double t1 = -0.3320883236095333, t2 = 0.11653281228448388;
  const double cc = a.cap, c2 = cc * cc;
  p.b0 = kLog2e;
  p.b1 = static_cast<float>(kLog2e * t1 / c2);
  p.b2 = static_cast<float>(kLog2e * t2 / (c2 * c2));
  p.range = 0.5f * a.cap;
  p.cap_l2 = a.cap * kLog2e;
  p.inv_cap = 1.0f / a.cap;
  static int granted[kMaxDevices] = {};
  {
    const cudaError_t e = ensure_dynamic_smem(attn_long_tcgen05_kernel, kSmemBytes, granted);
    if (e != cudaSuccess) return e;
  }
  const int grid = p.num_problems < num_sms() ? p.num_problems : num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, attn_long_tcgen05_kernel, tq, tkv, to, p);
}

}  // namespace vp
