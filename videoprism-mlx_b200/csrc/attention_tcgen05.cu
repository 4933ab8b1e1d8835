// tcgen05 / TMEM fused attention for the spatial stack: S = 256 tokens per frame, dh = 64, no mask, capped logits.
//
// One persistent CTA per SM walks (frame, head) problems.  Per problem the whole 256x256 fp32 score
// matrix lives in TMEM (2 query tiles x 256 columns = all 512 columns), so there is no K/V loop and no
// online-softmax rescale:
//   warp 0      : TMA producer (Q, K, V tiles of the packed qkv buffer -> 128B-swizzled smem, 2 stages)
//   warp 1      : MMA issuer   (S = Q K^T : tcgen05.mma M128 N256 K16 x4 per query tile, operands in smem;
//                               O = P V   : tcgen05.mma M128 N64 K16 x16, A = P in TMEM, B = V MN-major smem;
//                               l = P 1   : tcgen05.mma M128 N16 K16 x16 against a tile of ones: the softmax
//                                           normaliser, summed over exactly the bf16 weights that multiply V)
//   warps 2, 3, 20, 21 : drain warps (one per TMEM lane quarter): read O and l, normalise, stage bf16 O in the dead Q
//                 tile, TMA store, free the smem stage (warp 2 also allocates TMEM, warp 3 fills the ones tile)
//   warps 4..19 : softmax.  Warp (q, c) owns score rows [32q, 32q+32) x key columns [64c, 64c+64) of BOTH query
//                 tiles; a thread owns one score row.
// MUFU.EX2 is the bound of this kernel (16 per clock and SM: 4096 clk per problem), so the softmax warps do nothing
// but exponentiate, ONE query tile at a time, while the tensor core and the drain warps work for the other one:
//     exp(A) | exp(B) | exp(A') | exp(B') ...    PV(A), drain(A) and the next problem's S(A') all run under exp(B).
// The scores are read from TMEM exactly once: the logit cap cap*tanh(s/cap) (layers.py:586-594) bounds the exponent
// (72 in base 2 for cap = 50), so exp2 cannot overflow without a row maximum and softmax is shift-invariant.  The cap is
// an odd polynomial on the FMA pipe (packed f32x2; MUFU.TANH only for 32-column groups with |s| > cap/2); P is the
// exponential TRUNCATED to bf16 (one PRMT per pair; F2FP would share the MUFU pipe) written over dead score columns.
//
// Replaces DotProductAttention._dot_atten (layers.py:601-661) for the spatial encoder blocks.
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

constexpr int kTileBytes = 256 * 64 * 2;        // one of Q / K / V for a problem: 32 KB
constexpr int kStageBytes = 3 * kTileBytes;     // 96 KB
constexpr int kStages = 2;
constexpr int kSoftmaxWarps = 16;
constexpr int kThreads = 32 * (4 + kSoftmaxWarps + 2);   // 704: 4 control/drain warps, 16 softmax warps, 2 more drain warps
constexpr int kOnesBytes = 2048;   // bf16 1.0 tile: B operand of the row-sum MMA (any 16 x 16 window of it is all ones)
constexpr int kSmemBytes = kStages * kStageBytes + kOnesBytes + 1024 /*align slack*/ + 512 /*barriers*/;
constexpr float kLog2e = 1.4426950408889634f;

struct TcParams {
  int num_problems, heads, D;
  float b0, b1, b2;       // cap*log2e*tanh(s/cap) ~= s*(b0 + b1 s^2 + b2 s^4) for |s| <= range
  float range;
  float cap_l2, inv_cap;  // slow path: cap_l2 * tanh(s * inv_cap)
};

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// TMEM layout of one query tile (256 columns at `T`):
//   scores S            : [0, 256)     fp32, written by the S MMA; warp slice c reads columns [64c, 64c+64)
//   P (keys 64c..64c+63): 32 columns of bf16 pairs at {0, 32, 128, 160}[c].  Slices 1 and 3 write over score columns
//                         that slices 0 and 2 read (their second 32-column chunk): ordered by the rd_done barriers.
//   O                   : [64, 128)    fp32, written by the PV MMA after all of P is complete
//   l (row sums)        : [192, 208)   fp32, 16 identical columns, written by the P x ones MMA
__global__ void __launch_bounds__(kThreads, 1)
attn256_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ones_base = smem_base + kStages * kStageBytes;
  const uint32_t bar_base = ones_base + kOnesBytes;
  auto full_qk = [&](int s) { return bar_base + 8u * s; };
  auto full_v = [&](int s) { return bar_base + 8u * (2 + s); };
  auto empty = [&](int s) { return bar_base + 8u * (4 + s); };
  auto s_full = [&](int t) { return bar_base + 8u * (6 + t); };
  auto p_full = [&](int t) { return bar_base + 8u * (8 + t); };
  auto o_full = [&](int t) { return bar_base + 8u * (10 + t); };
  auto tmem_free = [&](int t) { return bar_base + 8u * (12 + t); };
  auto rd_done = [&](int t, int q, int pair) { return bar_base + 8u * (14 + (t * 4 + q) * 2 + pair); };   // 16 barriers
  const uint32_t tmem_ptr_addr = bar_base + 8u * 30;

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
      mbar_init(p_full(t), kSoftmaxWarps);
      mbar_init(o_full(t), 1);
      mbar_init(tmem_free(t), 4);
      for (int q = 0; q < 4; ++q)
        for (int pr = 0; pr < 2; ++pr) mbar_init(rd_done(t, q, pr), 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_addr, 512);
    tmem_relinquish();
  }
  if (warp == 3) {   // bf16 1.0 everywhere
    for (int i = lane; i < kOnesBytes / 16; i += 32)
      asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(ones_base + i * 16), "r"(0x3F803F80u) : "memory");
    fence_proxy_async_smem();   // read by the tensor core (async proxy)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));
  pdl_launch_dependents();   // the setup above overlapped the tail of the previous kernel (the QKV GEMM); its output is read below
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
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
    // -------------------------------------------------------------- MMA issuer
    // The softmax schedule fixes the order in which MMAs become issuable: S_A(0), S_B(0), then for every problem
    // PV_A(it), S_A(it+1), PV_B(it), S_B(it+1).  The issuer therefore BLOCKS on each event in that order (a polling
    // loop would steal issue slots from the four softmax warps that share its scheduler).
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);   // B = V is MN-major (dh contiguous per key)
      constexpr uint32_t idesc_sum = umma_idesc_bf16(128, 16, 0, 0);  // row sums of P: P x ones[256 x 16]
      const uint64_t d_ones = umma_desc_kmajor_sw128(ones_base);
      auto issue_s = [&](int t, int it) {
        // needs Q,K of the stage and the tile's TMEM (previous O and l read out)
        const int stage = it & 1;
        mbar_wait(full_qk(stage), (it >> 1) & 1u);
        mbar_wait(tmem_free(t), (it & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t T = tmem_base + t * 256;
        const uint32_t sq = smem_base + stage * kStageBytes;
        const uint64_t dq = umma_desc_kmajor_sw128(sq + t * (kTileBytes / 2));
        const uint64_t dk = umma_desc_kmajor_sw128(sq + kTileBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(T, dq + 2u * k, dk + 2u * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full(t));
      };
      auto issue_pv = [&](int t, int it) {
        const int stage = it & 1;
        mbar_wait(full_v(stage), (it >> 1) & 1u);
        mbar_wait(p_full(t), it & 1u);
        tc_fence_after();
        const uint32_t T = tmem_base + t * 256;
        // V tile: key k at byte k*128 (64 dh values), 8-key swizzle atoms of 1024 B; one K=16 step = 2 atoms
        const uint64_t dv = umma_desc_mnmajor_sw128(smem_base + stage * kStageBytes + 2 * kTileBytes, 1024, 1024);
#pragma unroll
        for (int k = 0; k < 16; ++k)
          umma_bf16_ts(T + 64, T + (k < 8 ? 8 * k : 128 + 8 * (k - 8)), dv + static_cast<uint64_t>(k) * (2048 >> 4), idesc_pv,
                       k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 16; ++k)
          umma_bf16_ts(T + 192, T + (k < 8 ? 8 * k : 128 + 8 * (k - 8)), d_ones, idesc_sum, k != 0 ? 1u : 0u);
        umma_commit(o_full(t));
      };
      if (n_it > 0) {
        issue_s(0, 0);
        issue_s(1, 0);
      }
      for (int it = 0; it < n_it; ++it) {
        issue_pv(0, it);
        if (it + 1 < n_it) issue_s(0, it + 1);
        issue_pv(1, it);
        if (it + 1 < n_it) issue_s(1, it + 1);
      }
    }
  } else if (warp == 2 || warp == 3 || warp >= 4 + kSoftmaxWarps) {
    // -------------------------------------------------------------- drain warps, one per TMEM lane quarter
    // O = P V and l = P 1 of a (tile, problem) are ready: normalise the 64 output columns of this thread's row, stage them
    // in the (dead) Q tile of the stage, TMA-store the 128 x 64 tile and release the stage.
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const bool leader = (warp == 3) && elect_one();
    for (int it = 0; it < n_it; ++it) {
      const int pr = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
      const int frame = pr / p.heads, h = pr % p.heads;
      const int stage = it & 1;
#pragma unroll 1
      for (int tile = 0; tile < 2; ++tile) {
        const uint32_t T = tmem_base + tile * 256 + lane_off;
        const uint32_t so = smem_base + stage * kStageBytes + tile * (kTileBytes / 2);
        const uint32_t rowaddr = so + row * 128;
        const int sw = row & 7;
        mbar_wait(o_full(tile), it & 1u);
        tc_fence_after();
        uint32_t rs;
        tmem_ld_32x32b_x1(T + 192, rs);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(T + 64 + 32 * half, o);
          tmem_ld_wait();
          if (half == 1) {
            // everything of this tile has left TMEM: the next problem's S may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_free(tile));
          }
          const float inv = 1.0f / __uint_as_float(rs);
          const f32x2 inv2 = pk2(inv, inv);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t wv[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              float a, b;
              upk2(mul2(pk2u(o[g * 8 + jj * 2], o[g * 8 + jj * 2 + 1]), inv2), a, b);
              wv[jj] = pack_bf16x2(a, b);
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (((4 * half + g) ^ sw) << 4)), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]), "r"(wv[3]) : "memory");
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);   // the four drain warps: the tile is complete
        if (leader) {
          tma_store_2d(&tmO, so, h * 64, frame * 256 + tile * 128);
          tma_store_commit();
          tma_store_wait_read<0>();
          mbar_arrive(empty(stage));   // this tile's share of the stage is dead and its O has left smem
        }
      }
    }
    if (leader) tma_store_wait<0>();
  } else if (warp >= 4 && warp < 4 + kSoftmaxWarps) {
    // -------------------------------------------------------------- softmax warps: (lane quarter q, key slice c)
    const int q = warp & 3;
    const int c = (warp - 4) >> 2;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int scol = 64 * c;                              // this warp's 64 score columns
    const int pcol = (c >> 1) * 128 + (c & 1) * 32;       // where its 32 columns of bf16 P go
    const f32x2 B0 = pk2(p.b0, p.b0), B1 = pk2(p.b1, p.b1), B2 = pk2(p.b2, p.b2);

    // cap, exp2 and bf16 P (w) for 32 keys of one query tile
    auto exp_chunk = [&](int tile, int j, uint32_t (&w)[16]) {
      const uint32_t T = tmem_base + tile * 256 + lane_off;
      uint32_t r[32];
      tmem_ld_32x32b_x32(T + scol + 32 * j, r);
      tmem_ld_wait();
      if ((c & 1) == 0 && j == 1) {
        // columns [64c+32, 64c+64) are consumed: slice c+1 may now write its P over them
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(rd_done(tile, q, c >> 1));
      }
      // the odd polynomial is valid for |s| <= range; larger logits (rare) take MUFU.TANH for this row's 32 columns
      float am[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 16; ++i) am[i & 3] = max3(am[i & 3], fabsf(__uint_as_float(r[2 * i])), fabsf(__uint_as_float(r[2 * i + 1])));
      const float amax = fmaxf(max3(am[0], am[1], am[2]), am[3]);
      if (amax <= p.range) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const f32x2 v = pk2u(r[2 * i], r[2 * i + 1]);
          const f32x2 u = mul2(v, v);
          f32x2 t = fma2(u, B2, B1);
          t = fma2(t, u, B0);
          const f32x2 x = mul2(t, v);   // capped logit, base-2 exponent, |x| <= 72.2
          if ((i & 3) == 3) {
            // every 4th pair takes exp2 on the FMA / ALU pipes instead of MUFU (the kernel's bound): round-to-nearest
            // split x = n + f by the 1.5*2^23 trick, 2^f by a cubic (relative error 1.0e-4, a 40th of the bf16 step of P),
            // n added into the exponent field
            const f32x2 tt = add2(x, pk2(12582912.f, 12582912.f));
            const f32x2 nn = add2(tt, pk2(-12582912.f, -12582912.f));
            const f32x2 fr = fma2(nn, pk2(-1.f, -1.f), x);
            f32x2 pp = fma2(fr, pk2(0.055008938f, 0.055008938f), pk2(0.24221096f, 0.24221096f));
            pp = fma2(pp, fr, pk2(0.69328293f, 0.69328293f));
            pp = fma2(pp, fr, pk2(1.f, 1.f));
            float ta, tb, pa, pb;
            upk2(tt, ta, tb);
            upk2(pp, pa, pb);
            const uint32_t ea = __float_as_uint(pa) + (__float_as_uint(ta) << 23);
            const uint32_t eb = __float_as_uint(pb) + (__float_as_uint(tb) << 23);
            w[i] = __byte_perm(ea, eb, 0x7632);
          } else {
            float a, b;
            upk2(x, a, b);
            w[i] = __byte_perm(__float_as_uint(ex2_approx(a)), __float_as_uint(ex2_approx(b)), 0x7632);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = p.cap_l2 * tanh_approx(__uint_as_float(r[2 * i]) * p.inv_cap);
          const float b = p.cap_l2 * tanh_approx(__uint_as_float(r[2 * i + 1]) * p.inv_cap);
          w[i] = __byte_perm(__float_as_uint(ex2_approx(a)), __float_as_uint(ex2_approx(b)), 0x7632);
        }
      }
    };
    // Even slices store each chunk's P at once (over their own consumed columns).  Odd slices write over columns the
    // even slice of their pair reads as its SECOND chunk, so they keep the first chunk's P in registers and store both
    // once that read is done.
    auto p_store = [&](int tile, int j, const uint32_t (&w)[16]) {
      tmem_st_32x32b_x16(tmem_base + tile * 256 + lane_off + pcol + 16 * j, w);
    };
    auto exp_first = [&](int tile, uint32_t (&w0)[16]) {
      exp_chunk(tile, 0, w0);
      if ((c & 1) == 0) p_store(tile, 0, w0);
    };
    auto exp_second = [&](int tile, uint32_t parity, const uint32_t (&w0)[16]) {
      uint32_t w1[16];
      exp_chunk(tile, 1, w1);
      if ((c & 1) == 1) {
        mbar_wait(rd_done(tile, q, c >> 1), parity);
        tc_fence_after();
        p_store(tile, 0, w0);
      }
      p_store(tile, 1, w1);
    };
    auto p_done = [&](int tile) {
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(tile));
    };
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ph = it & 1u;
      // ---- query tile A
      mbar_wait(s_full(0), ph);
      tc_fence_after();
      uint32_t w0[16];
      exp_first(0, w0);
      exp_second(0, ph, w0);
      p_done(0);
      // ---- query tile B
      mbar_wait(s_full(1), ph);
      tc_fence_after();
      exp_first(1, w0);
      exp_second(1, ph, w0);
      p_done(1);
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

// Returns cudaErrorNotSupported when the problem does not fit this kernel (caller falls back to the
// mma.sync kernels of attention.cu).
cudaError_t launch_attention_tcgen05(cudaStream_t s, const AttnArgs& a) {
  const int D = a.heads * a.dh;
  if (a.S != 256 || a.dh != 64 || a.group != 1 || a.key_pad != nullptr || a.causal) return cudaErrorNotSupported;
  // no row maximum is taken: the logit cap must bound the exponent (cap * log2e < 100 keeps exp2 and its row sums finite)
  if (!(a.cap > 0.f) || a.cap * kLog2e >= 100.0f) return cudaErrorNotSupported;
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
  // minimax fit of tanh(x)/x = 1 + t1 x^2 + t2 x^4 on |x| <= 0.5 (max error 2.7e-5 => <= 6.7e-4 in the capped logit at
  // |s| = cap/2 = 25, i.e. < 0.07 % in a softmax weight, a tenth of the bf16 step of P; exact in the limit s -> 0)
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
    const cudaError_t e = ensure_dynamic_smem(attn256_tcgen05_kernel, kSmemBytes, granted);
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
  return cudaLaunchKernelEx(&cfg, attn256_tcgen05_kernel, tq, to, p);
}

}  // namespace vp
