// tcgen05 / TMEM fused attention for the spatial stack: S = 256 tokens per frame, dh = 64, no mask.
//
// One persistent CTA per SM walks (frame, head) problems.  Per problem the whole 256x256 fp32 score
// matrix lives in TMEM (2 query tiles x 256 columns = all 512 columns), so there is no K/V loop and no
// online-softmax rescale:
//   warp 0     : TMA producer (Q, K, V tiles of the packed qkv buffer -> 128B-swizzled smem, 2 stages)
//   warp 1     : MMA issuer   (S = Q K^T : tcgen05.mma M128 N256 K16 x4 per query tile, operands in smem;
//                              O = P V   : tcgen05.mma M128 N64 K16 x16, A = P in TMEM, B = V MN-major smem)
//   warp 2     : TMEM allocator
//   warps 4-7  : softmax warpgroup for query rows   0..127 (TMEM columns   0..255)
//   warps 8-11 : softmax warpgroup for query rows 128..255 (TMEM columns 256..511)
// A softmax thread owns one score row and reads it from TMEM exactly once (tcgen05.ld runs at 64 B/clk per SM, so
// every extra pass over the 256 KB of scores costs as much as all of the kernel's MUFU.EX2 work): the logit cap
// cap*tanh(s/cap) (layers.py:586-594) bounds the exponent, so no row maximum is needed.  The cap is an odd polynomial
// on the FMA pipe (packed f32x2; MUFU.TANH only for 32-column groups with |s| > cap/2), exp2 runs on MUFU.EX2 and
// bf16 P is written back over the dead score columns (tcgen05.st).  O is accumulated next to P, normalised by the fp32 row sum,
// staged in the (dead) Q tile and written with one TMA store per query tile.
//
// Replaces DotProductAttention._dot_atten (layers.py:601-661) for the spatial encoder blocks.
#include <cuda.h>
#include <math_constants.h>

#include "kernels.h"
#include "ptx.cuh"

namespace vp {

bool make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                       uint32_t box_cols, int swizzle_bytes);
int num_sms();

namespace {

constexpr int kTileBytes = 256 * 64 * 2;        // one of Q / K / V for a problem: 32 KB
constexpr int kStageBytes = 3 * kTileBytes;     // 96 KB
constexpr int kStages = 2;
constexpr int kThreads = 640;   // 4 control warps + 16 softmax warps
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 128 + 4096 + 2048;   // + barriers + max/min and sum exchange
constexpr float kLog2e = 1.4426950408889634f;

struct TcParams {
  int num_problems, heads, D;
  float b0, b1, b2, b3;   // cap*log2e*tanh(s/cap) ~= s*(b0 + b1 s^2 + b2 s^4 + b3 s^6) for |s| <= range
  float range;
  float cap_l2, inv_cap;  // slow path: cap_l2 * tanh(s * inv_cap)
  int single_pass;        // capped logits are bounded (|cap * log2e| < 100): no row maximum is needed for exp2 to stay finite
};

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) { return mbar_try_wait(bar, parity); }

// TMEM layout of one query tile (256 columns at `T`):
//   scores S        : [0, 256)                     fp32, written by the S MMA
//   P (keys 0..127) : [0, 64)    P (keys 128..255) : [128, 192)   bf16 pairs, each half written over score
//                                                                 columns its own warp has already consumed
//   O               : [64, 128)                    fp32, written by the PV MMA after both P halves are complete
__global__ void __launch_bounds__(kThreads, 1)
attn256_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_qk = [&](int s) { return bar_base + 8u * s; };
  auto full_v = [&](int s) { return bar_base + 8u * (2 + s); };
  auto empty = [&](int s) { return bar_base + 8u * (4 + s); };
  auto s_full = [&](int t) { return bar_base + 8u * (6 + t); };
  auto p_full = [&](int t) { return bar_base + 8u * (8 + t); };
  auto o_full = [&](int t) { return bar_base + 8u * (10 + t); };
  auto tmem_free = [&](int t) { return bar_base + 8u * (12 + t); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * 14;
  const uint32_t xchg_base = bar_base + 128;   // float [2 tiles][2 halves][128 rows][2] max/min (4 KB) + [2][2][128] sums (2 KB)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_it = (p.num_problems - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_qk(s), 1);
      mbar_init(full_v(s), 1);
      mbar_init(empty(s), 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(s_full(t), 1);
      mbar_init(p_full(t), 8);
      mbar_init(o_full(t), 1);
      mbar_init(tmem_free(t), 8);
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

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int pr = blockIdx.x; pr < p.num_problems; pr += gridDim.x, ++it) {
        const int stage = it & 1;
        const uint32_t sphase = (it >> 1) & 1u;
        const int frame = pr / p.heads, h = pr % p.heads;
        mbar_wait(empty(stage), sphase ^ 1u);
        const uint32_t sq = smem_base + stage * kStageBytes;
        mbar_expect_tx(full_qk(stage), 2 * kTileBytes);
        tma_load_2d(sq, &tmQKV, full_qk(stage), h * 64, frame * 256);
        tma_load_2d(sq + kTileBytes, &tmQKV, full_qk(stage), p.D + h * 64, frame * 256);
        mbar_expect_tx(full_v(stage), kTileBytes);
        tma_load_2d(sq + 2 * kTileBytes, &tmQKV, full_v(stage), 2 * p.D + h * 64, frame * 256);
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer (event driven)
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);   // B = V is MN-major (dh contiguous per key)
      int s_it[2] = {0, 0}, pv_it[2] = {0, 0};
      while (pv_it[0] < n_it || pv_it[1] < n_it) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint32_t T = tmem_base + t * 256;
          if (s_it[t] < n_it && s_it[t] == pv_it[t]) {
            // S(it) of this tile: needs Q,K of the stage and the tile's TMEM (previous O read out)
            const int it = s_it[t];
            const int stage = it & 1;
            if (mbar_test(full_qk(stage), (it >> 1) & 1u) && mbar_test(tmem_free(t), (it & 1u) ^ 1u)) {
              tc_fence_after();
              const uint32_t sq = smem_base + stage * kStageBytes;
              const uint64_t dq = umma_desc_kmajor_sw128(sq + t * (kTileBytes / 2));
              const uint64_t dk = umma_desc_kmajor_sw128(sq + kTileBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_ss(T, dq + 2u * k, dk + 2u * k, idesc_s, k != 0 ? 1u : 0u);
              umma_commit(s_full(t));
              s_it[t] = it + 1;
            }
          } else if (pv_it[t] < s_it[t]) {
            const int it = pv_it[t];
            const int stage = it & 1;
            if (mbar_test(full_v(stage), (it >> 1) & 1u) && mbar_test(p_full(t), it & 1u)) {
              tc_fence_after();
              // V tile: key k at byte k*128 (64 dh values), 8-key swizzle atoms of 1024 B; one K=16 step = 2 atoms
              const uint64_t dv = umma_desc_mnmajor_sw128(smem_base + stage * kStageBytes + 2 * kTileBytes, 1024, 1024);
#pragma unroll
              for (int k = 0; k < 16; ++k)
                umma_bf16_ts(T + 64, T + (k < 8 ? 8 * k : 128 + 8 * (k - 8)), dv + static_cast<uint64_t>(k) * (2048 >> 4), idesc_pv,
                             k != 0 ? 1u : 0u);
              umma_commit(o_full(t));
              pv_it[t] = it + 1;
            }
          }
        }
      }
    }
  } else if (warp >= 4) {
    // -------------------------------------------------------------- softmax warps: (tile, column half, lane quarter)
    const int tile = (warp - 4) >> 3;
    const int ch = ((warp - 4) >> 2) & 1;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;                       // row inside the 128-row query tile
    const uint32_t T = tmem_base + tile * 256 + (static_cast<uint32_t>(wq * 32) << 16);
    const uint32_t t_s = T + ch * 128;                    // this warp's 128 score columns
    const uint32_t xme = xchg_base + ((tile * 2 + ch) * 128 + row) * 8;
    const uint32_t xpartner = xchg_base + ((tile * 2 + (ch ^ 1)) * 128 + row) * 8;
    const uint32_t xsum_me = xchg_base + 4096 + ((tile * 2 + ch) * 128 + row) * 4;         // separate slots: no reuse hazard
    const uint32_t xsum_partner = xchg_base + 4096 + ((tile * 2 + (ch ^ 1)) * 128 + row) * 4;
    const f32x2 B0 = pk2(p.b0, p.b0), B1 = pk2(p.b1, p.b1), B2 = pk2(p.b2, p.b2), B3 = pk2(p.b3, p.b3);
    int it = 0;
    for (int pr = blockIdx.x; pr < p.num_problems; pr += gridDim.x, ++it) {
      const int stage = it & 1;
      const uint32_t tphase = it & 1u;
      const int frame = pr / p.heads, h = pr % p.heads;
      mbar_wait(s_full(tile), tphase);
      tc_fence_after();
      // ---- pass 1 (only without a logit cap): row maximum, exchanged with the partner warp.  With the cap the
      // exponent is bounded by cap*log2e (72 for cap = 50), exp2 cannot overflow and softmax is shift-invariant, so the
      // scores are read from TMEM ONCE (TMEM reads run at 64 B/clk per SM: two passes over 256 KB of scores per
      // problem cost 8192 clk against 4096 clk of MUFU.EX2 and made the kernel TMEM-read bound).
      float m_l2 = 0.f;
      if (!p.single_pass) {
        float mx = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_s + 32 * j, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = max3(mx, __uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
        }
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(xme), "f"(mx) : "memory");
        named_bar_sync(1 + tile, 256);
        float pmx;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pmx) : "r"(xpartner));
        mx = fmaxf(mx, pmx);
        m_l2 = (p.cap_l2 > 0.f) ? p.cap_l2 * tanh_approx(mx * p.inv_cap) : mx * p.b0;
      }
      const f32x2 negm = pk2(-m_l2, -m_l2);
      // ---- cap, exp2, partial row sum, bf16 P written over this warp's own dead score columns
      f32x2 sum2 = pk2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_s + 32 * j, r);
        tmem_ld_wait();
        // the odd polynomial is valid for |s| <= range; larger logits (rare) take MUFU.TANH for this row's 32 columns
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) amax = max3(amax, fabsf(__uint_as_float(r[2 * i])), fabsf(__uint_as_float(r[2 * i + 1])));
        const bool fast = amax <= p.range;
        uint32_t w[16];
        if (fast) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const f32x2 v = pk2u(r[2 * i], r[2 * i + 1]);
            const f32x2 u = mul2(v, v);
            f32x2 q = fma2(u, B3, B2);
            q = fma2(q, u, B1);
            q = fma2(q, u, B0);
            float a, b;
            upk2(fma2(q, v, negm), a, b);
            // P is truncated to bf16 with integer ops (ALU pipe) instead of cvt.rn.bf16x2 (F2FP shares the MUFU pipe
            // with ex2 and was ~1/3 of its load); the row sum is taken over the TRUNCATED values, so the weights the
            // tensor core sees sum to exactly the normaliser.
            const uint32_t e0 = __float_as_uint(ex2_approx(a)) & 0xFFFF0000u, e1 = __float_as_uint(ex2_approx(b)) & 0xFFFF0000u;
            sum2 = add2(sum2, pk2u(e0, e1));
            w[i] = __byte_perm(e0, e1, 0x7632);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = fmaf(p.cap_l2, tanh_approx(__uint_as_float(r[2 * i]) * p.inv_cap), -m_l2);
            const float b = fmaf(p.cap_l2, tanh_approx(__uint_as_float(r[2 * i + 1]) * p.inv_cap), -m_l2);
            const uint32_t e0 = __float_as_uint(ex2_approx(a)) & 0xFFFF0000u, e1 = __float_as_uint(ex2_approx(b)) & 0xFFFF0000u;
            sum2 = add2(sum2, pk2u(e0, e1));
            w[i] = __byte_perm(e0, e1, 0x7632);
          }
        }
        tmem_st_32x32b_x16(t_s + 16 * j, w);   // keys [ch*128 + 32j, +32) -> P columns ch*128 + [16j, 16j+16)
      }
      tmem_st_wait();
      tc_fence_before();
      float s0, s1;
      upk2(sum2, s0, s1);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(xsum_me), "f"(s0 + s1) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(tile));
      named_bar_sync(1 + tile, 256);
      float psum;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(psum) : "r"(xsum_partner));
      const float inv = 1.0f / (s0 + s1 + psum);
      const f32x2 inv2 = pk2(inv, inv);
      // ---- O = P V is ready: this warp normalises output columns [32 ch, 32 ch + 32), stages them in the dead Q tile
      mbar_wait(o_full(tile), tphase);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld_32x32b_x32(T + 64 + 32 * ch, o);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_free(tile));
      const uint32_t so = smem_base + stage * kStageBytes + tile * (kTileBytes / 2);
      const uint32_t rowaddr = so + row * 128;
      const int sw = row & 7;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t wv[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float a, b;
          upk2(mul2(pk2u(o[c * 8 + jj * 2], o[c * 8 + jj * 2 + 1]), inv2), a, b);
          wv[jj] = pack_bf16x2(a, b);
        }
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (((4 * ch + c) ^ sw) << 4)), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]), "r"(wv[3]) : "memory");
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + tile, 256);
      if (ch == 0 && wq == 0 && lane == 0) {
        tma_store_2d(&tmO, so, h * 64, frame * 256 + tile * 128);
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(empty(stage));   // Q/K/V of this stage are dead (both PV MMAs retired before o_full) and O has left smem
      }
    }
    if (ch == 0 && wq == 0 && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Returns cudaErrorNotSupported when the problem does not fit this kernel (caller falls back to the
// mma.sync kernels of attention.cu).
cudaError_t launch_attention_tcgen05(cudaStream_t s, const AttnArgs& a) {
  const int D = a.heads * a.dh;
  if (a.S != 256 || a.dh != 64 || a.group != 1 || a.key_pad != nullptr || a.causal) return cudaErrorNotSupported;
  if (a.k != a.q + D || a.v != a.q + 2 * D || (a.ld % 8) || (a.ldo % 8)) return cudaErrorNotSupported;
  if ((reinterpret_cast<uintptr_t>(a.q) & 15) || (reinterpret_cast<uintptr_t>(a.out) & 15)) return cudaErrorNotSupported;
  const uint64_t rows = static_cast<uint64_t>(a.num_seq) * 256;
  CUtensorMap tq, to;
  if (!make_tmap_2d_bf16(&tq, a.q, rows, 3 * D, a.ld, 256, 64, 128)) return cudaErrorUnknown;
  if (!make_tmap_2d_bf16(&to, a.out, rows, D, a.ldo, 128, 64, 128)) return cudaErrorUnknown;
  TcParams p;
  p.num_problems = a.num_seq * a.heads;
  p.heads = a.heads;
  p.D = D;
  if (a.cap > 0.f) {
    // minimax fit of tanh(x)/x in x^2 on |x| <= 0.5 (max error 3.2e-7 => 1.6e-5 in the capped logit at cap = 50)
    const double t1 = -0.3332843058638077, t2 = 0.13208677223289364, t3 = -0.04484290508460146;
    const double c = a.cap, c2 = c * c;
    p.b0 = kLog2e;
    p.b1 = static_cast<float>(kLog2e * t1 / c2);
    p.b2 = static_cast<float>(kLog2e * t2 / (c2 * c2));
    p.b3 = static_cast<float>(kLog2e * t3 / (c2 * c2 * c2));
    p.range = 0.5f * a.cap;
    p.cap_l2 = a.cap * kLog2e;
    p.inv_cap = 1.0f / a.cap;
    p.single_pass = (p.cap_l2 < 100.0f) ? 1 : 0;
  } else {
    p.b0 = kLog2e; p.b1 = p.b2 = p.b3 = 0.f;
    p.range = 3.0e38f;
    p.cap_l2 = 0.f; p.inv_cap = 0.f;
    p.single_pass = 0;
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn256_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  const int grid = p.num_problems < num_sms() ? p.num_problems : num_sms();
  attn256_tcgen05_kernel<<<grid, kThreads, kSmemBytes, s>>>(tq, to, p);
  return cudaGetLastError();
}

}  // namespace vp
