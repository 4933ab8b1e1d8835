// bf16 GEMM for sm_100a: C[M,N] = A[M,K] * Wt[N,K]^T with a fused epilogue.
//
// Persistent, warp-specialised (640 threads):
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B-swizzled K-major tiles)
//   warp 1      : MMA issuer    (one elected thread, tcgen05.mma kind::f16; on SM pairs cta_group::2, M=256 x N=256 x K=16)
//   warps 2, 3  : TMEM allocator (warp 2) and "stagers": per-column {colsum, bias} and, for a folded LayerNorm, the
//                 per-row (rstd, -rstd*mean) of the NEXT tile -> shared memory
//   warps 4..19 : epilogue      (tcgen05.ld -> folded LayerNorm / bias / GELU / row-scale / pos-emb / residual -> bf16)
// The fp32 accumulator lives in TMEM and is double buffered (2 x BN columns), so the epilogue of
// tile i overlaps the main loop of tile i+1.  The smem ring has 5-6 slots of (128 x 64 A, BN/CG x 64 B) bf16.
//
// The epilogue is latency-, not throughput-bound (ncu: 2 epilogue warps per scheduler issued 0.19 IPC each and
// made the K=768 GEMMs epilogue-bound), so it runs 16 warps (4 per scheduler), each owning 32 accumulator rows x
// BN/4 columns, and everything but the accumulator is fetched before the MMA completes: the stagers' vectors are
// read back as broadcast 16-byte shared loads, the bf16 residual tile is TMA-loaded into the warp's staging tile.
// Data movement is all TMA: each converted 32 x 32 chunk goes through the warp's swizzled (conflict-free) staging
// tile and out with one cp.async.bulk.tensor.  (Row-per-thread global stores cost 32 L1 wavefronts per instruction.)
// Every single-lane issue site is predicated with elect.sync, not `lane == 0`: ptxas otherwise wraps each
// UTCHMMA / UTMALDG in an ELECT + BRA.U.ANY loop (11 instructions per MMA instead of 4).
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
constexpr int kFirstEpiWarp = 4;
constexpr int kNumEpiWarps = 16;                        // 4 per TMEM lane quarter: each owns 32 rows x BN/4 columns of a tile
constexpr int kNumThreads = 32 * (kFirstEpiWarp + kNumEpiWarps);   // 640 -> at most 96 registers per thread

// CG = CTAs per MMA (cta_group): 1 = one SM per 128 x BN tile; 2 = an SM pair per 256 x BN tile, each CTA holding
// its 128 A rows and HALF of the B rows (BN/2), which halves the per-SM smem fill traffic.
template <int BN, int CG>
struct Cfg {
  static constexpr int kBRows = BN / CG;                      // B rows resident per CTA
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kColsPerWarp = BN / 4;                 // 64 or 32 output columns per epilogue warp (1 or 2 chunks of 32)
  static constexpr int kStageTileBytes = 32 * 32 * 2;         // 32 x 32 bf16 staging tile
  // staging tiles per epilogue warp (run-time, KParams::nbuf): 1, reused by the warp's chunks (32 KB in all), or 2 for
  // short-K residual GEMMs, whose residual tiles are then all requested at tile start (64 KB, one pipeline stage less)
  static constexpr int staging_bytes(int nbuf) { return kNumEpiWarps * nbuf * kStageTileBytes; }
  static constexpr int kCvecBytes = BN * 8;                   // {colsum[n], colsum[n+1], bias[n], bias[n+1]} per column pair
  static constexpr int kRowvecBytes = BM * 8;                 // folded LayerNorm: (rstd, -rstd * mean) per accumulator row
  static constexpr int kBarBytes = 512;                       // (2*stages + 4 + 2*16 + 2) mbarriers + the TMEM base pointer
  static constexpr int kTmemCols = 2 * BN;  // 512 or 256: power of two
  // The main loop needs ~150 KB of loads in flight per SM (64 B/clk at ~1.5 us of L2/HBM latency), so everything
  // else is kept small and the rest of the 227 KB is pipeline: 6 stages of 32 KB for SM pairs (5 when the folded
  // LayerNorm needs its per-row vector).  The stage count is a run-time parameter of the kernel.
  static constexpr int stages(bool ln, int nbuf) {
    const int avail = 232448 - staging_bytes(nbuf) - kCvecBytes - (ln ? kRowvecBytes : 0) - kBarBytes;
    return (avail / kStageBytes) > 8 ? 8 : (avail / kStageBytes);
  }
  static constexpr int smem_bytes(bool ln, int nbuf) {
    return stages(ln, nbuf) * kStageBytes + staging_bytes(nbuf) + kCvecBytes + (ln ? kRowvecBytes : 0) + kBarBytes;
  }
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
  int resid_period;   // > 0: residual rows repeat with this period (a table), see GemmEpilogue
  const float* ln_stats_in;
  int ln_slots;
  const float* ln_colsum;
  float ln_inv_dim;
  float* stats_out;
  int stats_slots;
  int stages;   // smem ring depth (Cfg::stages)
  int nbuf;     // staging tiles per epilogue warp (1 or 2)
  int cl;       // CTAs per cluster: CG, or 4 = two SM pairs on M-adjacent tiles that share (TMA-multicast) their B tile
};

template <int BN, int ACT, bool RESID, bool OUT_F32, int CG>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const KParams p) {
  using C = Cfg<BN, CG>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();   // 128B-swizzled tiles need a 1024-byte aligned base (no slack is reserved for realigning)
  const int kStages = p.stages;
  const bool ln_fold = p.ln_stats_in != nullptr;
  const uint32_t staging_base = smem_base + kStages * C::kStageBytes;
  const uint32_t cvec_base = staging_base + C::staging_bytes(p.nbuf);
  const uint32_t rowvec_base = cvec_base + C::kCvecBytes;
  const uint32_t bar_base = rowvec_base + (ln_fold ? C::kRowvecBytes : 0);
  // barriers (8 B each): full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], resid[16 warps][2], vec_full, vec_empty, then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  auto resid_bar = [&](int w, int b) { return bar_base + 8u * (2 * kStages + 4 + 2 * w + b); };
  const uint32_t vec_full_bar = bar_base + 8u * (2 * kStages + 4 + 2 * kNumEpiWarps);
  const uint32_t vec_empty_bar = vec_full_bar + 8u;
  const uint32_t tmem_ptr_addr = vec_empty_bar + 8u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  constexpr int TM = BM * CG;                                  // rows of one (pair) tile
  // A cluster holds npairs SM pairs (1, or 2 with p.cl == 4) that work on M-adjacent tiles of the same N tile, so
  // they read the same B (weight) tile: each CTA fetches half of its B share and multicasts it to its twin in the other
  // pair.  A "tile" below is the cluster's unit: npairs x TM rows by BN columns.
  const int crank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const int cta_rank = crank & 1;                              // rank inside the MMA pair
  const int pi = crank >> 1;                                   // which pair of the cluster
  const int npairs = (CG == 2) ? p.cl / 2 : 1;
  const int num_m_tiles = ((p.M + TM - 1) / TM + npairs - 1) / npairs;
  const int num_n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = num_m_tiles * num_n_tiles;
  const int num_kb = (p.K + BK - 1) / BK;
  const int tile0 = static_cast<int>(blockIdx.x) / ((CG == 2) ? p.cl : 1);         // first tile of this cluster
  const int tile_step = static_cast<int>(gridDim.x) / ((CG == 2) ? p.cl : 1);
  auto tile_m0 = [&](int tile) { return ((tile / num_n_tiles) * npairs + pi) * TM + cta_rank * BM; };   // this CTA's 128 A rows

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (!OUT_F32) tma_prefetch_desc(&tmC);
    if (RESID) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), npairs);   // every pair of the cluster must have consumed the slot (B is multicast into it)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kNumEpiWarps * CG);   // pair: the leader's barrier collects both CTAs' epilogue warps
    }
    for (int w = 0; w < kNumEpiWarps; ++w) { mbar_init(resid_bar(w, 0), 1); mbar_init(resid_bar(w, 1), 1); }
    mbar_init(vec_full_bar, 2);
    mbar_init(vec_empty_bar, kNumEpiWarps);
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
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of the previous kernel
  // in the stream; from here on its output is read.
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // (elect_one, not lane == 0: ptxas then issues the warp-uniform TMA / tcgen05 instructions directly instead of
    // wrapping each one in an ELECT / BRA.U.ANY loop)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m0 = tile_m0(tile);
        const int n0 = (tile % num_n_tiles) * BN + cta_rank * C::kBRows;     // this CTA's share of the B rows
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::kStageBytes;
          const uint32_t sb = sa + C::kABytes;
          if (CG == 2 && npairs == 2) {
            // Clusters of 4: as for plain pairs the bytes are credited to the pair leader's full barrier; this CTA fetches
            // half of its B share (64 rows) and multicasts it to itself and its twin (same pair rank, other pair), the
            // twin supplies the other half.
            const int hrows = C::kBRows / 2;
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * C::kStageBytes);
            tma_load_2d_pair(sa, &tmA, full_bar(stage), kb * BK, m0);
            tma_load_2d_pair_mc(sb + pi * hrows * (BK * 2), &tmB, full_bar(stage), kb * BK, n0 + pi * hrows,
                                static_cast<uint16_t>(5u << cta_rank));
          } else if (CG == 2) {
            // both CTAs' bytes are credited to the leader's full barrier; only the leader arms it
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * C::kStageBytes);
            tma_load_2d_pair(sa, &tmA, full_bar(stage), kb * BK, m0);
            tma_load_2d_pair(sb, &tmB, full_bar(stage), kb * BK, n0);
          } else {
            mbar_expect_tx(full_bar(stage), C::kStageBytes);
            tma_load_2d(sa, &tmA, full_bar(stage), kb * BK, m0);
            tma_load_2d(sb, &tmB, full_bar(stage), kb * BK, n0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    if (cta_rank == 0 && elect_one()) {
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
          if (CG == 2) umma_commit_pair(empty_bar(stage), static_cast<uint16_t>(npairs == 2 ? 0xF : 0x3)); else umma_commit(empty_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (CG == 2) umma_commit_pair(tfull_bar(acc), static_cast<uint16_t>(0x3 << (2 * pi))); else umma_commit(tfull_bar(acc));  // accumulator complete
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ------------------------------------------------- vector stagers (64 threads)
    // Per tile: the per-column constants {colsum, bias} of its BN columns and, for a folded LayerNorm, (rstd,
    // -rstd * mean) of this CTA's 128 rows.  The global loads of tile i+1 are issued while the epilogue still works
    // on tile i (their L2/HBM latency used to sit on the epilogue's critical path); the single shared-memory copy is
    // rewritten once every epilogue warp has released it.
    const int ht = static_cast<int>(threadIdx.x) - 64;   // 0..63
    int it = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
      const int m0 = tile_m0(tile);
      const int n0 = (tile % num_n_tiles) * BN;
      float4 cs = make_float4(0.f, 0.f, 0.f, 0.f), bs = make_float4(0.f, 0.f, 0.f, 0.f);
      const int n = n0 + ht * 4;
      if (ht * 4 < BN && n < p.N) {   // N % 8 == 0: groups of 4 columns are all-or-nothing
        if (ln_fold) cs = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + n));
        if (p.bias != nullptr) bs = __ldg(reinterpret_cast<const float4*>(p.bias + n));
      }
      float2 ab[2] = {make_float2(1.f, 0.f), make_float2(1.f, 0.f)};
      if (ln_fold) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int m = m0 + ht + rr * 64;
          float2 ss = make_float2(0.f, 0.f);
          if (m < p.M) {
            const float2* sp = reinterpret_cast<const float2*>(p.ln_stats_in) + static_cast<size_t>(m) * p.ln_slots;
            for (int sl = 0; sl < p.ln_slots; ++sl) {   // fixed order: bit-reproducible statistics
              const float2 t = __ldg(sp + sl);
              ss.x += t.x; ss.y += t.y;
            }
          }
          const float mean = ss.x * p.ln_inv_dim;
          const float var = fmaxf(ss.y * p.ln_inv_dim - mean * mean, 0.f);
          const float rstd = rsqrtf(var + 1e-6f);
          ab[rr] = make_float2(rstd, -rstd * mean);
        }
      }
      if (it > 0) mbar_wait(vec_empty_bar, (it - 1) & 1u);
      if (ht * 4 < BN) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cvec_base + ht * 32), "f"(cs.x), "f"(cs.y), "f"(bs.x), "f"(bs.y) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cvec_base + ht * 32 + 16), "f"(cs.z), "f"(cs.w), "f"(bs.z), "f"(bs.w) : "memory");
      }
      if (ln_fold) {
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(rowvec_base + ht * 8), "f"(ab[0].x), "f"(ab[0].y) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(rowvec_base + (ht + 64) * 8), "f"(ab[1].x), "f"(ab[1].y) : "memory");
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(vec_full_bar);
    }
  } else if (warp >= kFirstEpiWarp) {
    // ---------------------------------------------------------------- epilogue
    // 16 warps: warp (q, slice) owns accumulator rows [32q, 32q+32) (its TMEM lane quarter) x columns
    // [slice*CW, (slice+1)*CW).  Everything a tile needs besides the accumulator is fetched BEFORE the wait on the
    // accumulator barrier, so it overlaps the main loop: the per-column constants {colsum, bias} go to shared
    // memory once per tile (one float per epilogue thread, read back as broadcast 16-byte loads), the folded
    // LayerNorm's per-row statistics and the residual tile (TMA into the staging buffer) likewise.
    const int e = warp - kFirstEpiWarp;
    const bool leader = elect_one();   // the one lane of this warp that issues (and later waits for) its TMA operations
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int slice = e >> 2;        // which quarter of the BN columns
    constexpr int CW = C::kColsPerWarp;
    constexpr int NCH = CW / 32;
    const bool two_bufs = p.nbuf == 2;   // chunk ch of a tile uses staging tile ch (else all chunks share tile 0)
    const uint32_t stg = staging_base + e * p.nbuf * C::kStageTileBytes;
    // 64-byte rows, 16-byte chunk c of row `lane` lives at chunk (c ^ ((lane >> 1) & 3))  (TMA SWIZZLE_64B)
    const uint32_t rowaddr = stg + lane * 64;
    const int sw = (lane >> 1) & 3;
    uint32_t rphase = 0;
    int it = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step, ++it) {
      const int m0 = tile_m0(tile);
      const int n0 = (tile % num_n_tiles) * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int mrow0 = m0 + q * 32;
      const int ncol0 = n0 + slice * CW;
      const int rrow0 = (RESID && p.resid_period > 0) ? mrow0 % p.resid_period : mrow0;   // first row of the residual tile
      if (!OUT_F32 && leader) {
        tma_store_wait_read<0>();   // the previous tile's last store has finished reading the staging tile
        if (RESID) {
          mbar_expect_tx(resid_bar(e, 0), C::kStageTileBytes);
          tma_load_2d(stg, &tmR, resid_bar(e, 0), ncol0, rrow0);
          if (two_bufs && NCH > 1) {
            mbar_expect_tx(resid_bar(e, 1), C::kStageTileBytes);
            tma_load_2d(stg + C::kStageTileBytes, &tmR, resid_bar(e, 1), ncol0 + 32, rrow0);
          }
        }
      }
      const int m = mrow0 + lane;
      const bool row_ok = m < p.M;
      const float rscale = (p.row_scale != nullptr && row_ok) ? __ldg(p.row_scale + m) : 1.0f;
      const float* pos_row = nullptr;
      if (p.pos_table != nullptr) pos_row = p.pos_table + static_cast<size_t>(m % p.pos_period) * p.N;
      mbar_wait(vec_full_bar, it & 1u);   // the stager warps have published this tile's column / row vectors
      // folded LayerNorm of the A rows: v = ln_a * acc + ln_b * colsum[n] + bias[n]
      f32x2 ln_a2 = pk2(1.f, 1.f), ln_b2 = pk2(0.f, 0.f);
      if (ln_fold) {
        float a, b;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(rowvec_base + (q * 32 + lane) * 8));
        ln_a2 = pk2(a, a);
        ln_b2 = pk2(b, b);
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (!OUT_F32 && !RESID) __syncwarp();   // lane 0 saw the staging tile drain
      float st_sum = 0.f, st_sq = 0.f;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int col = slice * CW + ch * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + acc * BN + col + (static_cast<uint32_t>(q * 32) << 16), r);
        tmem_ld_wait();
        if (ch == NCH - 1) {
          // all TMEM reads of this warp for this tile are done: release the accumulator early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(tempty_bar(acc), 2 * pi); else mbar_arrive(tempty_bar(acc));
          }
        }
        const int n = n0 + col;
        const uint32_t cv = cvec_base + col * 8;
        f32x2 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float c0, c1, b0, b1;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(c0), "=f"(c1), "=f"(b0), "=f"(b1) : "r"(cv + i * 16));
          v[i] = fma2(pk2u(r[2 * i], r[2 * i + 1]), ln_a2, fma2(ln_b2, pk2(c0, c1), pk2(b0, b1)));
        }
        if (ch == NCH - 1) {
          __syncwarp();
          if (lane == 0) mbar_arrive(vec_empty_bar);   // this warp no longer reads the tile's column / row vectors
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
          const f32x2 rscale2 = pk2(rscale, rscale);
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
                  const uint2 rr = *reinterpret_cast<const uint2*>(p.resid + static_cast<size_t>(rrow0 + lane) * p.ldr + n + g * 4);
                  a += bf16_lo(rr.x); b += bf16_hi(rr.x); c += bf16_lo(rr.y); d += bf16_hi(rr.y);
                }
                *reinterpret_cast<float4*>(cp + g * 4) = make_float4(a, b, c, d);
              }
            }
          }
        } else {
          const int bsel = (two_bufs && ch > 0) ? 1 : 0;
          const uint32_t boff = bsel * C::kStageTileBytes;
          if (RESID) {
            // chunk 0's residual tile was requested at tile start; chunk 1's too when it has its own staging tile,
            // else right after chunk 0's store (below)
            mbar_wait(resid_bar(e, bsel), (rphase >> bsel) & 1u);
            rphase ^= 1u << bsel;
          } else if (ch > 0 && !two_bufs) {
            if (leader) tma_store_wait_read<0>();   // chunk 0's store has finished reading the staging tile
            __syncwarp();
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t addr = rowaddr + boff + ((c ^ sw) << 4);
            if (RESID) {
              uint32_t w0, w1, w2, w3;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(addr));
              v[4 * c + 0] = add2(v[4 * c + 0], pk2(bf16_lo(w0), bf16_hi(w0)));
              v[4 * c + 1] = add2(v[4 * c + 1], pk2(bf16_lo(w1), bf16_hi(w1)));
              v[4 * c + 2] = add2(v[4 * c + 2], pk2(bf16_lo(w2), bf16_hi(w2)));
              v[4 * c + 3] = add2(v[4 * c + 3], pk2(bf16_lo(w3), bf16_hi(w3)));
            }
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
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (leader) {
            tma_store_2d(&tmC, stg + boff, n, mrow0);   // clipped against [M, N] by the tensor map
            tma_store_commit();
            if (RESID && !two_bufs && ch + 1 < NCH) {
              tma_store_wait_read<0>();
              mbar_expect_tx(resid_bar(e, 0), C::kStageTileBytes);
              tma_load_2d(stg, &tmR, resid_bar(e, 0), n + 32, rrow0);
            }
          }
          __syncwarp();
        }
      }
      if (!OUT_F32 && p.stats_out != nullptr && row_ok) {
        const int slot = (tile % num_n_tiles) * 4 + slice;
        reinterpret_cast<float2*>(p.stats_out)[static_cast<size_t>(m) * p.stats_slots + slot] = make_float2(st_sum, st_sq);
      }
    }
    if (!OUT_F32 && leader) tma_store_wait<0>();
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

bool pdl_enabled() {
  static const bool on = !(getenv("VP_PDL") && atoi(getenv("VP_PDL")) == 0);
  return on;
}

int num_sms() {   // of the current device (cached per device ordinal)
  static int cache[kMaxDevices] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

namespace {

struct Maps {
  CUtensorMap a, b, c, r;
};

template <int BN, int ACT, bool RESID, bool OUT_F32, int CG>
cudaError_t launch_gemm_t(cudaStream_t s, const Maps& m, const KParams& kp_in, int grid) {
  auto kern = gemm_bf16_kernel<BN, ACT, RESID, OUT_F32, CG>;
  KParams kp = kp_in;
  if (CG != 2) kp.cl = 1;
  // A second staging tile per warp (all residual tiles of a tile requested at its start, one pipeline stage less) was
  // measured neutral for the out-projection (178 vs 175 us in situ): off unless VP_GEMM_NBUF=2.
  static const int nbuf_env = getenv("VP_GEMM_NBUF") ? atoi(getenv("VP_GEMM_NBUF")) : 1;
  kp.nbuf = (nbuf_env == 2 && RESID && !OUT_F32 && BN == 256 && kp.K <= 1024) ? 2 : 1;
  kp.stages = Cfg<BN, CG>::stages(kp.ln_stats_in != nullptr, kp.nbuf);
  const int kSmem = Cfg<BN, CG>::smem_bytes(kp.ln_stats_in != nullptr, kp.nbuf);
  static int granted[kMaxDevices] = {};   // one per template instantiation
  {
    const cudaError_t e = ensure_dynamic_smem(kern, 232448, granted);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = kSmem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = kp.cl; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
    if (kp.cl == 4) {
      // clusters of four 227 KB CTAs do not tile all 148 SMs (GPCs of 16 / 18 / 20 SMs): ask how many fit
      static int max_clusters = 0;
      if (max_clusters == 0) {
        cfg.attrs = attr; cfg.numAttrs = na;
        cfg.gridDim = dim3(4 * 37);
        if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters <= 0) max_clusters = 32;
      }
      const int want = grid / 4;
      cfg.gridDim = dim3(4 * (want < max_clusters ? want : max_clusters));
    }
  }
  if (pdl_enabled()) {   // programmatic dependent launch: the prologue overlaps the previous kernel's tail (ptx.cuh: pdl_wait)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
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
  return 4 * ((N + BN - 1) / BN);
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
  // Clusters of 4 (two SM pairs on M-adjacent tiles sharing one TMA-multicast B tile: 25 % fewer L2 reads) are
  // implemented but OFF: only 33 such clusters fit (132 of 148 SMs) and the forward measured 1077 vs 1050-1083 clips/s,
  // i.e. the power the main loop spends on operand delivery is not in the L2 read that multicast saves.
  static const int cluster_env = getenv("VP_GEMM_CLUSTER") ? atoi(getenv("VP_GEMM_CLUSTER")) : 2;
  const int cl = (CG == 2 && cluster_env == 4 && pair_tiles >= num_sms()) ? 4 : CG;
  Maps m;
  if (!make_tmap_2d_bf16(&m.a, A, M, K, lda, BM, BK, 128)) return cudaErrorUnknown;
  if (!make_tmap_2d_bf16(&m.b, Wt, N, K, ldb, cl == 4 ? BN / 4 : BN / CG, BK, 128)) return cudaErrorUnknown;
  if (!epi.out_f32) {
    if (!make_tmap_2d_bf16(&m.c, Cout, M, N, ldc, 32, 32, 64)) return cudaErrorUnknown;
  } else {
    m.c = m.a;
  }
  if (epi.resid_period < 0 || (epi.resid_period % 32) != 0) return cudaErrorInvalidValue;
  if (epi.resid != nullptr && !epi.out_f32) {
    if (!make_tmap_2d_bf16(&m.r, epi.resid, epi.resid_period > 0 ? epi.resid_period : M, N, epi.ldr, 32, 32, 64)) return cudaErrorUnknown;
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
  kp.resid = epi.resid; kp.ldr = epi.ldr; kp.resid_period = epi.resid != nullptr ? epi.resid_period : 0;
  kp.ln_stats_in = epi.ln_stats_in; kp.ln_colsum = epi.ln_colsum; kp.ln_slots = epi.ln_slots > 0 ? epi.ln_slots : 1;
  kp.stats_slots = gemm_stats_slots(N);
  kp.ln_inv_dim = epi.ln_dim > 0 ? 1.0f / static_cast<float>(epi.ln_dim) : 0.f;
  kp.stats_out = epi.stats_out;
  if (epi.ln_stats_in != nullptr && (epi.ln_colsum == nullptr || epi.ln_dim <= 0)) return cudaErrorInvalidValue;
  if (epi.stats_out != nullptr && epi.out_f32) return cudaErrorInvalidValue;
  kp.cl = cl;
  if (CG == 2) {
    if (cl == 4) {   // grid in CTAs; launch_gemm_t clamps it to the number of clusters that fit
      const int cluster_tiles = (((M + 255) / 256 + 1) / 2) * (N / 256);
      return launch_gemm_bn<256, 2>(s, m, kp, 4 * cluster_tiles, epi.act, epi.resid != nullptr, epi.out_f32 != 0);
    }
    const int pairs = pair_tiles < num_sms() / 2 ? pair_tiles : num_sms() / 2;
    return launch_gemm_bn<256, 2>(s, m, kp, 2 * pairs, epi.act, epi.resid != nullptr, epi.out_f32 != 0);
  }
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (BN == 256) return launch_gemm_bn<256, 1>(s, m, kp, grid, epi.act, epi.resid != nullptr, epi.out_f32 != 0);
  return launch_gemm_bn<128, 1>(s, m, kp, grid, epi.act, epi.resid != nullptr, epi.out_f32 != 0);
}

}  // namespace vp
