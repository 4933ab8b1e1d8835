// bf16 GEMM for sm_100a: C[M,N] = A[M,K] * Wt[N,K]^T with a fused epilogue.
//
// Persistent, warp-specialised:
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B-swizzled K-major tiles)
//   warp 1      : MMA issuer    (one elected thread, tcgen05.mma kind::f16, M=128 x N=BN x K=16)
//   warp 2      : TMEM allocator
//   warps 4..11 : epilogue      (tcgen05.ld -> bias / GELU / row-scale / pos-emb / residual -> bf16)
// The fp32 accumulator lives in TMEM and is double buffered (2 x BN columns), so the epilogue of
// tile i overlaps the main loop of tile i+1.  The smem ring has kStages slots of (128 x 64 A,
// BN x 64 B) bf16.
//
// Epilogue data movement is all TMA: each epilogue warp owns 32 accumulator rows and walks its
// columns in 32-column chunks; a chunk is converted in registers (packed f32x2 math), written to a
// per-warp 32x32 bf16 staging tile (64B swizzle, conflict-free, two ping-pong buffers) and stored
// with cp.async.bulk.tensor; the bf16 residual tile is TMA-loaded into the same staging buffer
// ahead of time and added in place.  (Row-per-thread global stores cost 32 L1 wavefronts per
// instruction and made the K=768 GEMMs epilogue-bound.)
//
// Replaces the reference's nn.Dense / einsum projections (layers.py:304-312, :486-488, :483-498).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"
#include "ptx.cuh"

namespace vp {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kNumThreads = 384;
constexpr int kFirstEpiWarp = 4;
constexpr int kNumEpiWarps = 8;
constexpr int kStageTileBytes = 32 * 32 * 2;            // one 32x32 bf16 staging tile

// CG = CTAs per MMA (cta_group): 1 = one SM per 128 x BN tile; 2 = an SM pair per 256 x BN tile, each CTA holding
// its 128 A rows and HALF of the B rows (BN/2), which halves the per-SM smem fill traffic and allows 6 stages.
// NBUF = staging tiles per epilogue warp: 2 (ping-pong), or 4 for the residual GEMMs on SM pairs, whose smaller
// pipeline stages leave room to prefetch the residual tiles of all four column chunks at tile start.
template <int BN, int CG, int NBUF>
struct Cfg {
  static constexpr int kBRows = BN / CG;                      // B rows resident per CTA
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = kNumEpiWarps * NBUF * kStageTileBytes;               // 32 or 64 KB
  static constexpr int kStages = ((229376 - kStagingBytes) / kStageBytes) > 8 ? 8 : ((229376 - kStagingBytes) / kStageBytes);
  static constexpr int kTmemCols = 2 * BN;  // 512 or 256: power of two
  static constexpr int kPipeBytes = kStages * kStageBytes;
  static constexpr int kSmemBytes = kPipeBytes + kStagingBytes + 1024 /*align slack*/ + 512 /*barriers*/;
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
  const float* ln_stats_in;
  int ln_slots;
  const float* ln_colsum;
  float ln_inv_dim;
  float* stats_out;
  int stats_slots;
};

__device__ __forceinline__ float gelu_erf_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN, int ACT, bool RESID, bool OUT_F32, int CG>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const KParams p) {
  constexpr int NBUF = (RESID && CG == 2 && !OUT_F32) ? 4 : 2;
  using C = Cfg<BN, CG, NBUF>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + C::kPipeBytes;
  const uint32_t bar_base = staging_base + C::kStagingBytes;
  // barriers (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], resid[8 warps][2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::kStages + 2 + a); };
  auto resid_bar = [&](int w, int b) { return bar_base + 8u * (2 * C::kStages + 4 + w * NBUF + b); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * C::kStages + 4 + NBUF * kNumEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  constexpr int TM = BM * CG;                                  // rows of one (pair) tile
  const int num_m_tiles = (p.M + TM - 1) / TM;
  const int num_n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int num_kb = (p.K + BK - 1) / BK;
  const int cta_rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int tile0 = static_cast<int>(blockIdx.x) / CG;         // first tile of this CTA (pair)
  const int tile_step = static_cast<int>(gridDim.x) / CG;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (!OUT_F32) tma_prefetch_desc(&tmC);
    if (RESID) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kNumEpiWarps * CG);   // pair: the leader's barrier collects both CTAs' epilogue warps
    }
    for (int w = 0; w < kNumEpiWarps; ++w)
      for (int b = 0; b < NBUF; ++b) mbar_init(resid_bar(w, b), 1);
    fence_mbar_init();
  }
  if (CG == 2) cluster_sync_all();   // barrier inits of both CTAs are visible before any remote arrive / multicast commit
  if (warp == 2) {
    if (CG == 2) { tmem_alloc_pair(tmem_ptr_addr, C::kTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_ptr_addr, C::kTmemCols); tmem_relinquish(); }
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
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m0 = (tile / num_n_tiles) * TM + cta_rank * BM;            // this CTA's 128 A rows
        const int n0 = (tile % num_n_tiles) * BN + cta_rank * C::kBRows;     // this CTA's share of the B rows
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint32_t sb = sa + C::kABytes;
          if (CG == 2) {
            // both CTAs' bytes are credited to the leader's full barrier; only the leader arms it
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * C::kStageBytes);
            tma_load_2d_pair(sa, &tmA, full_bar(stage), kb * BK, m0);
            tma_load_2d_pair(sb, &tmB, full_bar(stage), kb * BK, n0);
          } else {
            mbar_expect_tx(full_bar(stage), C::kStageBytes);
            tma_load_2d(sa, &tmA, full_bar(stage), kb * BK, m0);
            tma_load_2d(sb, &tmB, full_bar(stage), kb * BK, n0);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(TM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
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
            if (CG == 2) umma_bf16_ss_pair(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16_ss(d_tmem, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs retire
          if (CG == 2) umma_commit_pair(empty_bar(stage)); else umma_commit(empty_bar(stage));
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
        if (CG == 2) umma_commit_pair(tfull_bar(acc)); else umma_commit(tfull_bar(acc));  // accumulator complete
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ---------------------------------------------------------------- epilogue
    const int e = warp - kFirstEpiWarp;
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int half = e >> 2;         // which half of the BN columns
    constexpr int kColsPerWarp = BN / 2;
    constexpr int NCH = kColsPerWarp / 32;
    const uint32_t stg = staging_base + e * NBUF * kStageTileBytes;
    uint32_t rphase = 0;   // bit b = parity of residual barrier b
    int it = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
      const int m0 = (tile / num_n_tiles) * TM + cta_rank * BM;
      const int n0 = (tile % num_n_tiles) * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int mrow0 = m0 + q * 32;
      const int ncol0 = n0 + half * kColsPerWarp;
      if (RESID && !OUT_F32) {
        // prefetch the residual tiles of the first two chunks into the two staging buffers
        if (lane == 0) {
          if (NBUF == 4) {
            // one buffer per column chunk: the previous tile's four stores were issued in buffer order
            tma_store_wait_read<3>();
            mbar_expect_tx(resid_bar(e, 0), kStageTileBytes);
            tma_load_2d(stg, &tmR, resid_bar(e, 0), ncol0, mrow0);
            tma_store_wait_read<2>();
            mbar_expect_tx(resid_bar(e, 1), kStageTileBytes);
            tma_load_2d(stg + kStageTileBytes, &tmR, resid_bar(e, 1), ncol0 + 32, mrow0);
            tma_store_wait_read<1>();
            mbar_expect_tx(resid_bar(e, 2), kStageTileBytes);
            tma_load_2d(stg + 2 * kStageTileBytes, &tmR, resid_bar(e, 2), ncol0 + 64, mrow0);
            tma_store_wait_read<0>();
            mbar_expect_tx(resid_bar(e, 3), kStageTileBytes);
            tma_load_2d(stg + 3 * kStageTileBytes, &tmR, resid_bar(e, 3), ncol0 + 96, mrow0);
          } else {
            tma_store_wait_read<1>();   // the store that last used buffer 0 has drained
            mbar_expect_tx(resid_bar(e, 0), kStageTileBytes);
            tma_load_2d(stg, &tmR, resid_bar(e, 0), ncol0, mrow0);
            if (NCH > 1) {
              tma_store_wait_read<0>();
              mbar_expect_tx(resid_bar(e, 1), kStageTileBytes);
              tma_load_2d(stg + kStageTileBytes, &tmR, resid_bar(e, 1), ncol0 + 32, mrow0);
            }
          }
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int m = mrow0 + lane;
      const bool row_ok = m < p.M;
      const float rscale = (p.row_scale != nullptr && row_ok) ? __ldg(p.row_scale + m) : 1.0f;
      const f32x2 rscale2 = pk2(rscale, rscale);
      const float* pos_row = nullptr;
      if (p.pos_table != nullptr) pos_row = p.pos_table + static_cast<size_t>(m % p.pos_period) * p.N;
      // folded LayerNorm of the A rows: v = ln_a * acc + ln_b * colsum[n] + bias[n]
      f32x2 ln_a2 = pk2(1.f, 1.f), ln_b2 = pk2(0.f, 0.f);
      if (p.ln_stats_in != nullptr) {
        float2 ss = make_float2(0.f, 0.f);
        if (row_ok) {
          const float2* sp = reinterpret_cast<const float2*>(p.ln_stats_in) + static_cast<size_t>(m) * p.ln_slots;
          for (int sl = 0; sl < p.ln_slots; ++sl) {   // fixed order: bit-reproducible statistics
            const float2 t = __ldg(sp + sl);
            ss.x += t.x; ss.y += t.y;
          }
        }
        const float mean = ss.x * p.ln_inv_dim;
        const float var = fmaxf(ss.y * p.ln_inv_dim - mean * mean, 0.f);
        const float rstd = rsqrtf(var + 1e-6f);
        ln_a2 = pk2(rstd, rstd);
        ln_b2 = pk2(-rstd * mean, -rstd * mean);
      }
      float st_sum = 0.f, st_sq = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < NCH; ++ch) {
        const int col = half * kColsPerWarp + ch * 32;
        // bias for this chunk is fetched before the TMEM load so the two latencies overlap
        float4 bv[8];
        if (p.bias != nullptr) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            bv[g] = (n0 + col + g * 4 < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + n0 + col + g * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 cs[8];
        if (p.ln_stats_in != nullptr) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            cs[g] = (n0 + col + g * 4 < p.N) ? __ldg(reinterpret_cast<const float4*>(p.ln_colsum + n0 + col + g * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + acc * BN + col + (static_cast<uint32_t>(q * 32) << 16), r);
        tmem_ld_wait();
        if (ch == NCH - 1) {
          // all TMEM reads of this warp for this tile are done: release the accumulator early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(tempty_bar(acc), 0); else mbar_arrive(tempty_bar(acc));
          }
        }
        const int n = n0 + col;
        f32x2 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = pk2u(r[2 * i], r[2 * i + 1]);
        if (p.ln_stats_in != nullptr) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            v[2 * g] = fma2(v[2 * g], ln_a2, mul2(pk2(cs[g].x, cs[g].y), ln_b2));
            v[2 * g + 1] = fma2(v[2 * g + 1], ln_a2, mul2(pk2(cs[g].z, cs[g].w), ln_b2));
          }
        }
        if (p.bias != nullptr) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            v[2 * g] = add2(v[2 * g], pk2(bv[g].x, bv[g].y));
            v[2 * g + 1] = add2(v[2 * g + 1], pk2(bv[g].z, bv[g].w));
          }
        }
        if (ACT == ACT_GELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = gelu2(v[i]);
        } else if (ACT == ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float a, b;
            upk2(v[i], a, b);
            v[i] = pk2(fmaxf(a, 0.f), fmaxf(b, 0.f));
          }
        }
        if (p.row_scale != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = mul2(v[i], rscale2);
        }
        if (pos_row != nullptr && row_ok) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (n + g * 4 < p.N) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(pos_row + n + g * 4));
              v[2 * g] = add2(v[2 * g], pk2(b.x, b.y));
              v[2 * g + 1] = add2(v[2 * g + 1], pk2(b.z, b.w));
            }
          }
        }
        if (OUT_F32) {
          // diagnostic / test path: direct fp32 stores (residual, if any, read directly)
          if (row_ok) {
            float* cp = reinterpret_cast<float*>(p.C) + static_cast<size_t>(m) * p.ldc + n;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (n + g * 4 < p.N) {
                float a, b, c, d;
                upk2(v[2 * g], a, b);
                upk2(v[2 * g + 1], c, d);
                if (RESID) {
                  const uint2 rr = *reinterpret_cast<const uint2*>(p.resid + static_cast<size_t>(m) * p.ldr + n + g * 4);
                  a += bf16_lo(rr.x); b += bf16_hi(rr.x); c += bf16_lo(rr.y); d += bf16_hi(rr.y);
                }
                *reinterpret_cast<float4*>(cp + g * 4) = make_float4(a, b, c, d);
              }
            }
          }
        } else {
          const int b = (NBUF == 4) ? ch : (ch & 1);
          const uint32_t buf = stg + b * kStageTileBytes;
          // 64-byte rows, 16-byte chunk c of row `lane` lives at chunk (c ^ ((lane >> 1) & 3))  (TMA SWIZZLE_64B)
          const uint32_t rowaddr = buf + lane * 64;
          const int sw = (lane >> 1) & 3;
          if (RESID) {
            mbar_wait(resid_bar(e, b), (rphase >> b) & 1u);
            rphase ^= 1u << b;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t w0, w1, w2, w3;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(rowaddr + ((c ^ sw) << 4)));
              v[4 * c + 0] = add2(v[4 * c + 0], pk2(bf16_lo(w0), bf16_hi(w0)));
              v[4 * c + 1] = add2(v[4 * c + 1], pk2(bf16_lo(w1), bf16_hi(w1)));
              v[4 * c + 2] = add2(v[4 * c + 2], pk2(bf16_lo(w2), bf16_hi(w2)));
              v[4 * c + 3] = add2(v[4 * c + 3], pk2(bf16_lo(w3), bf16_hi(w3)));
            }
          } else {
            if (lane == 0) tma_store_wait_read<1>();   // the store issued two chunks ago (same buffer) has drained
            __syncwarp();
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float lo, hi;
              upk2(v[4 * c + j], lo, hi);
              w[j] = pack_bf16x2(lo, hi);
            }
            if (p.stats_out != nullptr && n + c * 8 < p.N) {
              // statistics of the ROUNDED values: exactly what the next (LayerNorm-folded) GEMM will read
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float lo = bf16_lo(w[j]), hi = bf16_hi(w[j]);
                st_sum += lo + hi;
                st_sq = fmaf(lo, lo, fmaf(hi, hi, st_sq));
              }
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + ((c ^ sw) << 4)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, buf, n, mrow0);   // clipped against [M, N] by the tensor map
            tma_store_commit();
            if (RESID && NBUF == 2 && ch + 2 < NCH) {
              // buffer (ch+1)&1 ... is busy; the buffer for chunk ch+2 is this one: wait for the store just issued
              // to finish reading it, then prefetch that chunk's residual tile
              tma_store_wait_read<0>();
              mbar_expect_tx(resid_bar(e, b), kStageTileBytes);
              tma_load_2d(buf, &tmR, resid_bar(e, b), n + 64, mrow0);
            }
          }
          __syncwarp();
        }
      }
      if (!OUT_F32 && p.stats_out != nullptr && row_ok) {
        const int slot = (tile % num_n_tiles) * 2 + half;
        reinterpret_cast<float2*>(p.stats_out)[static_cast<size_t>(m) * p.stats_slots + slot] = make_float2(st_sum, st_sq);
      }
    }
    if (!OUT_F32 && lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // no CTA of the pair exits (or frees TMEM) while its peer may still signal it
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, C::kTmemCols); else tmem_dealloc(tmem_base, C::kTmemCols);
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
// box = box_cols x box_rows, zero fill out of bounds.  swizzle_bytes: 128 or 64 (= box_cols * 2).
bool make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                       uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(bf16)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

namespace {

struct Maps {
  CUtensorMap a, b, c, r;
};

template <int BN, int ACT, bool RESID, bool OUT_F32, int CG>
cudaError_t launch_gemm_t(cudaStream_t s, const Maps& m, const KParams& kp, int grid) {
  auto kern = gemm_bf16_kernel<BN, ACT, RESID, OUT_F32, CG>;
  constexpr int kSmem = Cfg<BN, CG, (RESID && CG == 2 && !OUT_F32) ? 4 : 2>::kSmemBytes;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  if (CG == 1) {
    kern<<<grid, kNumThreads, kSmem, s>>>(m.a, m.b, m.c, m.r, kp);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, m.a, m.b, m.c, m.r, kp);
}

template <int BN, bool RESID, bool OUT_F32, int CG>
cudaError_t launch_gemm_act(cudaStream_t s, const Maps& m, const KParams& kp, int grid, int act) {
  switch (act) {
    case ACT_NONE: return launch_gemm_t<BN, ACT_NONE, RESID, OUT_F32, CG>(s, m, kp, grid);
    case ACT_GELU: return launch_gemm_t<BN, ACT_GELU, RESID, OUT_F32, CG>(s, m, kp, grid);
    case ACT_RELU: return launch_gemm_t<BN, ACT_RELU, RESID, OUT_F32, CG>(s, m, kp, grid);
  }
  return cudaErrorInvalidValue;
}

template <int BN, int CG>
cudaError_t launch_gemm_bn(cudaStream_t s, const Maps& m, const KParams& kp, int grid, int act, bool resid, bool out_f32) {
  if (out_f32) return resid ? launch_gemm_act<BN, true, true, CG>(s, m, kp, grid, act) : launch_gemm_act<BN, false, true, CG>(s, m, kp, grid, act);
  return resid ? launch_gemm_act<BN, true, false, CG>(s, m, kp, grid, act) : launch_gemm_act<BN, false, false, CG>(s, m, kp, grid, act);
}

}  // namespace

int gemm_stats_slots(int N) {
  const int BN = (N % 256 == 0) ? 256 : 128;
  return 2 * ((N + BN - 1) / BN);
}

cudaError_t launch_gemm(cudaStream_t s, const bf16* A, int lda, const bf16* Wt, int ldb, void* Cout, int ldc, int M, int N,
                        int K, const GemmEpilogue& epi) {
  if (M <= 0 || N <= 0 || K <= 0) return cudaErrorInvalidValue;
  if ((K % 8) || (N % 8) || (lda % 8) || (ldb % 8) || (ldc % 8)) return cudaErrorInvalidValue;
  if (epi.resid != nullptr && (epi.ldr % 8)) return cudaErrorInvalidValue;
  const int BN = (N % 256 == 0) ? 256 : 128;
  // SM pairs (cta_group::2, 256 x 256 tiles) when the problem has at least one pair-tile per pair of SMs
  static const int force_cg = getenv("VP_GEMM_CTA_GROUP") ? atoi(getenv("VP_GEMM_CTA_GROUP")) : 0;
  const int pair_tiles = ((M + 255) / 256) * (N / 256);
  int CG = (BN == 256 && !epi.out_f32 && pair_tiles >= num_sms() / 2) ? 2 : 1;
  if (force_cg == 1) CG = 1;
  if (force_cg == 2 && BN == 256) CG = 2;
  Maps m;
  if (!make_tmap_2d_bf16(&m.a, A, M, K, lda, BM, BK, 128)) return cudaErrorUnknown;
  if (!make_tmap_2d_bf16(&m.b, Wt, N, K, ldb, BN / CG, BK, 128)) return cudaErrorUnknown;
  if (!epi.out_f32) {
    if (!make_tmap_2d_bf16(&m.c, Cout, M, N, ldc, 32, 32, 64)) return cudaErrorUnknown;
  } else {
    m.c = m.a;
  }
  if (epi.resid != nullptr && !epi.out_f32) {
    if (!make_tmap_2d_bf16(&m.r, epi.resid, M, N, epi.ldr, 32, 32, 64)) return cudaErrorUnknown;
  } else {
    m.r = m.a;
  }
  KParams kp;
  kp.M = M; kp.N = N; kp.K = K;
  kp.C = Cout; kp.ldc = ldc;
  kp.bias = epi.bias;
  kp.row_scale = epi.row_scale;
  kp.pos_table = epi.pos_table;
  kp.pos_period = epi.pos_period > 0 ? epi.pos_period : 1;
  kp.resid = epi.resid; kp.ldr = epi.ldr;
  kp.ln_stats_in = epi.ln_stats_in; kp.ln_colsum = epi.ln_colsum; kp.ln_slots = epi.ln_slots > 0 ? epi.ln_slots : 1;
  kp.stats_slots = gemm_stats_slots(N);
  kp.ln_inv_dim = epi.ln_dim > 0 ? 1.0f / static_cast<float>(epi.ln_dim) : 0.f;
  kp.stats_out = epi.stats_out;
  if (epi.ln_stats_in != nullptr && (epi.ln_colsum == nullptr || epi.ln_dim <= 0)) return cudaErrorInvalidValue;
  if (epi.stats_out != nullptr && epi.out_f32) return cudaErrorInvalidValue;
  if (CG == 2) {
    const int pairs = pair_tiles < num_sms() / 2 ? pair_tiles : num_sms() / 2;
    return launch_gemm_bn<256, 2>(s, m, kp, 2 * pairs, epi.act, epi.resid != nullptr, epi.out_f32 != 0);
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (BN == 256) return launch_gemm_bn<256, 1>(s, m, kp, grid, epi.act, epi.resid != nullptr, epi.out_f32 != 0);
  return launch_gemm_bn<128, 1>(s, m, kp, grid, epi.act, epi.resid != nullptr, epi.out_f32 != 0);
}

}  // namespace vp
