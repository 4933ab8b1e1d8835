// Fused softmax attention (scores never leave the SM): q.k^T -> cap*tanh(./cap) -> mask -> fp32
// online softmax -> P.V, for sequences embedded in a packed [rows, ld] bf16 qkv buffer.
//
// Replaces DotProductAttention._dot_atten (layers.py:601-661) incl. the logit cap (:586-594) and
// the mask semantics of :51-179.  Two kernels share the per-warp tile routine:
//   attn_flash_kernel : 64 queries per CTA (4 warps x 16 rows), K/V streamed in 64-key tiles through a
//                       double-buffered cp.async ring (spatial S=256, auxiliary S=4096, text S=65)
//   attn_small_kernel : S <= 16 (temporal stack): one warp per (sequence, head), 4 heads per CTA
// Tensor work is mma.sync m16n8k16 bf16 with fp32 accumulation; the tcgen05 kernel for the unmasked
// S % 256 == 0, dh = 64 cases (spatial stack, auxiliary encoder) lives in attention_kloop_tcgen05.cu.
#include <math_constants.h>

#include <stdlib.h>

#include "kernels.h"
#include "ptx.cuh"

namespace vp {

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kMasked = -1.0e30f;  // finite: a fully masked row becomes uniform, as in the reference

template <int DH>
__device__ __forceinline__ uint32_t swz(int r, int c) {
  // byte offset of 16-byte chunk c of row r in a [rows][DH] bf16 tile (conflict-free for ldmatrix)
  if (DH == 128) return static_cast<uint32_t>(r * 256 + ((c ^ (r & 7)) << 4));
  if (DH == 64) return static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4));
  return static_cast<uint32_t>(r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int DH>
struct WarpState {
  float o[DH / 8][4];
  float m[2];
  float l[2];
};

// One warp, 16 query rows (A fragments in qf), NT*8 keys resident in smem at sK / sV.
//   key0   : sequence index of the tile's first key
//   kflag  : smem floats, 1 = key padded (per key of the tile), or nullptr
//   qi0/1  : sequence index of this thread's two rows; qpad0/1: query-padded (only read when causal)
template <int DH, int NT>
__device__ __forceinline__ void warp_attend_tile(WarpState<DH>& st, const uint32_t (&qf)[DH / 16][4], uint32_t sK, uint32_t sV,
                                                 int key0, int S, float cap, const float* kflag, bool causal, int qi0, int qi1,
                                                 bool qpad0, bool qpad1, int lane) {
  float c[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { c[nt][0] = c[nt][1] = c[nt][2] = c[nt][3] = 0.f; }
  // ---- S = Q K^T
#pragma unroll
  for (int ks = 0; ks < DH / 16; ++ks) {
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
      const int chunk = ks * 2 + ((lane >> 3) & 1);
      uint32_t b0, b1, b2, b3;
      ldsm_x4(sK + swz<DH>(row, chunk), b0, b1, b2, b3);
      mma_bf16(c[2 * np], qf[ks], b0, b1);
      mma_bf16(c[2 * np + 1], qf[ks], b2, b3);
    }
  }
  // ---- cap, mask, to log2 domain
  const float inv_cap = cap > 0.f ? 1.0f / cap : 0.f;
  const float cap_l2 = cap * kLog2e;
  float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int jl = nt * 8 + (lane & 3) * 2 + (e & 1);
      const int j = key0 + jl;
      const int qi = (e < 2) ? qi0 : qi1;
      float s = c[nt][e];
      s = cap > 0.f ? cap_l2 * tanh_approx(s * inv_cap) : s * kLog2e;
      bool masked = (kflag != nullptr) && (kflag[jl] > 0.5f);
      if (causal) masked = masked || (j > qi) || ((e < 2) ? qpad0 : qpad1);
      s = masked ? kMasked : s;
      s = (j >= S) ? -CUDART_INF_F : s;
      c[nt][e] = s;
      if (e < 2) mx0 = fmaxf(mx0, s); else mx1 = fmaxf(mx1, s);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  const float mn0 = fmaxf(st.m[0], mx0), mn1 = fmaxf(st.m[1], mx1);
  const float sc0 = ex2_approx(st.m[0] - mn0), sc1 = ex2_approx(st.m[1] - mn1);
  st.m[0] = mn0; st.m[1] = mn1;
  float rs0 = 0.f, rs1 = 0.f;
  uint32_t pa[NT / 2][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const float p0 = ex2_approx(c[nt][0] - mn0), p1 = ex2_approx(c[nt][1] - mn0);
    const float p2 = ex2_approx(c[nt][2] - mn1), p3 = ex2_approx(c[nt][3] - mn1);
    rs0 += p0 + p1; rs1 += p2 + p3;
    pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
    pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
  }
  st.l[0] = st.l[0] * sc0 + rs0;
  st.l[1] = st.l[1] * sc1 + rs1;
#pragma unroll
  for (int dt = 0; dt < DH / 8; ++dt) { st.o[dt][0] *= sc0; st.o[dt][1] *= sc0; st.o[dt][2] *= sc1; st.o[dt][3] *= sc1; }
  // ---- O += P V
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
#pragma unroll
    for (int dp = 0; dp < DH / 16; ++dp) {
      const int row = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
      const int chunk = dp * 2 + (lane >> 4);
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(sV + swz<DH>(row, chunk), b0, b1, b2, b3);
      mma_bf16(st.o[2 * dp], pa[kk], b0, b1);
      mma_bf16(st.o[2 * dp + 1], pa[kk], b2, b3);
    }
  }
}

template <int DH>
__device__ __forceinline__ void warp_load_q_frags(uint32_t (&qf)[DH / 16][4], uint32_t sQ, int row_base, int lane) {
#pragma unroll
  for (int ks = 0; ks < DH / 16; ++ks) {
    const int row = row_base + (lane & 15);
    const int chunk = ks * 2 + (lane >> 4);
    ldsm_x4(sQ + swz<DH>(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  }
}

// normalise, stage through smem (rows row_base..+15 of sO, this warp only), 16-byte stores
template <int DH>
__device__ __forceinline__ void warp_store_out(WarpState<DH>& st, uint8_t* sO_gen, uint32_t sO, int row_base, int lane,
                                               bf16* out, int ldo, size_t row_first, int row_stride, int valid_rows, int col0,
                                               int real_chunks) {
  float l0 = st.l[0], l1 = st.l[1];
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __syncwarp();
#pragma unroll
  for (int dt = 0; dt < DH / 8; ++dt) {
    const int r0 = row_base + (lane >> 2), r1 = r0 + 8;
    const uint32_t off = (lane & 3) * 4;
    *reinterpret_cast<uint32_t*>(sO_gen + swz<DH>(r0, dt) + off) = pack_bf16x2(st.o[dt][0] * i0, st.o[dt][1] * i0);
    *reinterpret_cast<uint32_t*>(sO_gen + swz<DH>(r1, dt) + off) = pack_bf16x2(st.o[dt][2] * i1, st.o[dt][3] * i1);
  }
  __syncwarp();
  constexpr int CPR = DH / 8;  // 16-byte chunks per row
  for (int idx = lane; idx < 16 * CPR; idx += 32) {
    const int r = idx / CPR, ch = idx % CPR;
    if (r < valid_rows && ch < real_chunks) {
      const uint4 v = *reinterpret_cast<const uint4*>(sO_gen + swz<DH>(row_base + r, ch));
      *reinterpret_cast<uint4*>(out + (row_first + static_cast<size_t>(r) * row_stride) * ldo + col0 + ch * 8) = v;
    }
  }
  (void)sO;
}

// ---------------------------------------------------------------- flash kernel
template <int DH>
__global__ void __launch_bounds__(128) attn_flash_kernel(const AttnArgs a) {
  // DH is the padded head dimension of the smem tiles; a.dh (<= DH, a multiple of 8) the real one: 16-byte chunks past
  // it are zero-filled on load (they add nothing to q.k and give zero context columns) and never stored.
  constexpr int BQ = 64, BKV = 64, CPR = DH / 8;
  extern __shared__ __align__(128) uint8_t attn_smem[];
  uint8_t* sQ = attn_smem;
  uint8_t(*sK)[BKV * DH * 2] = reinterpret_cast<uint8_t(*)[BKV * DH * 2]>(attn_smem + BQ * DH * 2);
  uint8_t(*sV)[BKV * DH * 2] = reinterpret_cast<uint8_t(*)[BKV * DH * 2]>(attn_smem + BQ * DH * 2 + 2 * BKV * DH * 2);
  float(*sFlag)[BKV] = reinterpret_cast<float(*)[BKV]>(attn_smem + BQ * DH * 2 + 4 * BKV * DH * 2);
  const int real_chunks = a.dh / 8;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nqb = (a.S + BQ - 1) / BQ;
  const int q0 = (blockIdx.x % nqb) * BQ;
  const int h = blockIdx.y;
  const int sid = blockIdx.x / nqb;
  const size_t row_first = static_cast<size_t>(sid / a.group) * a.group * a.S + (sid % a.group);
  const int rstride = a.group;
  const int col0 = h * a.dh;
  const uint32_t sQa = smem_u32(sQ);

  // Q tile
  for (int idx = tid; idx < BQ * CPR; idx += 128) {
    const int r = idx / CPR, ch = idx % CPR;
    const int qi = q0 + r;
    const bool ok = qi < a.S && ch < real_chunks;
    const bf16* src = a.q + (row_first + static_cast<size_t>(ok ? qi : 0) * rstride) * a.ld + col0 + (ok ? ch : 0) * 8;
    cp_async16(sQa + swz<DH>(r, ch), src, ok);
  }
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * BKV;
    const uint32_t sKa = smem_u32(sK[buf]), sVa = smem_u32(sV[buf]);
    for (int idx = tid; idx < BKV * CPR; idx += 128) {
      const int r = idx / CPR, ch = idx % CPR;
      const int kj = k0 + r;
      const bool ok = kj < a.S && ch < real_chunks;
      const size_t row = row_first + static_cast<size_t>(ok ? kj : 0) * rstride;
      cp_async16(sKa + swz<DH>(r, ch), a.k + row * a.ld + col0 + (ok ? ch : 0) * 8, ok);
      cp_async16(sVa + swz<DH>(r, ch), a.v + row * a.ld + col0 + (ok ? ch : 0) * 8, ok);
    }
    if (a.key_pad != nullptr && tid < BKV) {
      const int kj = k0 + tid;
      sFlag[buf][tid] = (kj < a.S) ? a.key_pad[static_cast<size_t>(sid) * a.S + kj] : 0.f;
    }
  };
  const int num_tiles = (a.S + BKV - 1) / BKV;
  load_kv(0, 0);
  cp_async_commit();

  WarpState<DH> st;
#pragma unroll
  for (int dt = 0; dt < DH / 8; ++dt) st.o[dt][0] = st.o[dt][1] = st.o[dt][2] = st.o[dt][3] = 0.f;
  st.m[0] = st.m[1] = -CUDART_INF_F;
  st.l[0] = st.l[1] = 0.f;

  const int qi0 = q0 + warp * 16 + (lane >> 2), qi1 = qi0 + 8;
  bool qpad0 = false, qpad1 = false;
  if (a.causal && a.key_pad != nullptr) {
    qpad0 = qi0 < a.S && a.key_pad[static_cast<size_t>(sid) * a.S + qi0] > 0.5f;
    qpad1 = qi1 < a.S && a.key_pad[static_cast<size_t>(sid) * a.S + qi1] > 0.5f;
  }
  uint32_t qf[DH / 16][4];
  for (int t = 0; t < num_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < num_tiles) {
      load_kv(t + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) warp_load_q_frags<DH>(qf, sQa, warp * 16, lane);
    warp_attend_tile<DH, BKV / 8>(st, qf, smem_u32(sK[buf]), smem_u32(sV[buf]), t * BKV, a.S, a.cap,
                                  a.key_pad != nullptr ? sFlag[buf] : nullptr, a.causal != 0, qi0, qi1, qpad0, qpad1, lane);
    __syncthreads();
  }
  const int valid = min(16, a.S - (q0 + warp * 16));
  if (valid > 0) {
    warp_store_out<DH>(st, sQ, sQa, warp * 16, lane, a.out, a.ldo, row_first + static_cast<size_t>(q0 + warp * 16) * rstride,
                       rstride, valid, col0, real_chunks);
  }
}

// ---------------------------------------------------------------- small kernel (S <= 16)
template <int DH>
__global__ void __launch_bounds__(128) attn_small_kernel(const AttnArgs a, int total_problems) {
  constexpr int CPR = DH / 8;
  constexpr int WPB = DH > 64 ? 2 : 4;   // warps (problems) per block: 3 tiles of 16 x DH bf16 each, within 48 KB of static smem
  __shared__ __align__(128) uint8_t sQ[WPB][16 * DH * 2];
  __shared__ __align__(128) uint8_t sK[WPB][16 * DH * 2];
  __shared__ __align__(128) uint8_t sV[WPB][16 * DH * 2];
  __shared__ float sFlag[WPB][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int real_chunks = a.dh / 8;
  const int prob = blockIdx.x * WPB + warp;  // = sid * heads + h
  if (prob >= total_problems) return;
  const int sid = prob / a.heads, h = prob % a.heads;
  const size_t row_first = static_cast<size_t>(sid / a.group) * a.group * a.S + (sid % a.group);
  const int rstride = a.group;
  const int col0 = h * a.dh;
  const uint32_t sQa = smem_u32(sQ[warp]), sKa = smem_u32(sK[warp]), sVa = smem_u32(sV[warp]);
  for (int idx = lane; idx < 16 * CPR; idx += 32) {
    const int r = idx / CPR, ch = idx % CPR;
    const bool ok = r < a.S && ch < real_chunks;
    const size_t off = (row_first + static_cast<size_t>(ok ? r : 0) * rstride) * a.ld + col0 + (ok ? ch : 0) * 8;
    cp_async16(sQa + swz<DH>(r, ch), a.q + off, ok);
    cp_async16(sKa + swz<DH>(r, ch), a.k + off, ok);
    cp_async16(sVa + swz<DH>(r, ch), a.v + off, ok);
  }
  cp_async_commit();
  if (a.key_pad != nullptr && lane < 16) sFlag[warp][lane] = (lane < a.S) ? a.key_pad[static_cast<size_t>(sid) * a.S + lane] : 0.f;
  WarpState<DH> st;
#pragma unroll
  for (int dt = 0; dt < DH / 8; ++dt) st.o[dt][0] = st.o[dt][1] = st.o[dt][2] = st.o[dt][3] = 0.f;
  st.m[0] = st.m[1] = -CUDART_INF_F;
  st.l[0] = st.l[1] = 0.f;
  const int qi0 = lane >> 2, qi1 = qi0 + 8;
  bool qpad0 = false, qpad1 = false;
  if (a.causal && a.key_pad != nullptr) {
    qpad0 = qi0 < a.S && a.key_pad[static_cast<size_t>(sid) * a.S + qi0] > 0.5f;
    qpad1 = qi1 < a.S && a.key_pad[static_cast<size_t>(sid) * a.S + qi1] > 0.5f;
  }
  cp_async_wait<0>();
  __syncwarp();
  uint32_t qf[DH / 16][4];
  warp_load_q_frags<DH>(qf, sQa, 0, lane);
  warp_attend_tile<DH, 2>(st, qf, sKa, sVa, 0, a.S, a.cap, a.key_pad != nullptr ? sFlag[warp] : nullptr, a.causal != 0, qi0, qi1,
                          qpad0, qpad1, lane);
  warp_store_out<DH>(st, sQ[warp], sQa, 0, lane, a.out, a.ldo, row_first, rstride, a.S, col0, real_chunks);
}

// Sequences whose keys are ALL padded (a padded frame in the spatial stack): every logit becomes min_value
// (layers.py:155-179), the softmax is uniform, and each query row receives the mean of V.  Lets frame-padded batches keep
// the tcgen05 kernel: it runs unmasked on every frame and this kernel overwrites the (rare) padded ones.
// grid (num_seq, heads), block 256 = 4 row groups x 64 lanes (2 bf16 each for dh = 64... one per lane pair for dh = 32).
__global__ void __launch_bounds__(256) attn_uniform_rows_kernel(const AttnArgs a) {
  const int seq = blockIdx.x, h = blockIdx.y;
  if (a.key_pad[static_cast<size_t>(seq) * a.S] < 0.5f) return;
  __shared__ float part[4][64];
  const int g = threadIdx.x >> 6;
  const size_t row0 = static_cast<size_t>(seq / a.group) * a.S * a.group + (seq % a.group);
  for (int j0 = 0; j0 < a.dh; j0 += 64) {
    const int j = j0 + (threadIdx.x & 63);
    float acc = 0.f;
    if (j < a.dh)
      for (int s = g; s < a.S; s += 4) acc += __bfloat162float(a.v[(row0 + static_cast<size_t>(s) * a.group) * a.ld + h * a.dh + j]);
    part[g][j - j0] = acc;
    __syncthreads();
    if (j < a.dh) {
      const bf16 m = __float2bfloat16((part[0][j - j0] + part[1][j - j0] + part[2][j - j0] + part[3][j - j0]) / static_cast<float>(a.S));
      for (int s = g; s < a.S; s += 4) a.out[(row0 + static_cast<size_t>(s) * a.group) * a.ldo + h * a.dh + j] = m;
    }
    __syncthreads();
  }
}

}  // namespace

// attention_kloop_tcgen05.cu: tcgen05 / TMEM kernel for unmasked sequences with S % 256 == 0 and dh = 64 (the spatial stack
// and the auxiliary encoder); returns cudaErrorNotSupported for anything else, which then runs on the kernels of this file
cudaError_t launch_attention_kloop_tcgen05(cudaStream_t s, const AttnArgs& a);

cudaError_t launch_attention(cudaStream_t s, const AttnArgs& a) {
  if (a.num_seq <= 0 || a.S <= 0) return cudaSuccess;
  if ((a.ld % 8) || (a.ldo % 8) || a.dh <= 0 || (a.dh % 8) || a.dh > 128 || a.group < 1) return cudaErrorInvalidValue;
  // smem tile width: 32 / 64 exactly, anything else (e.g. the giant configuration's dim_per_head = 88) zero-padded to 128
  const int DHT = a.dh == 32 ? 32 : a.dh == 64 ? 64 : 128;
  if (a.launched) *a.launched = 1;
  if (!a.force_mma_sync && a.key_pad != nullptr && a.pad_whole_seq && !a.causal) {
    AttnArgs b = a;
    b.key_pad = nullptr;
    cudaError_t e = launch_attention_kloop_tcgen05(s, b);   // unmasked fast path on every sequence ...
    if (e == cudaSuccess) {                           // ... then the fully padded sequences get the uniform-softmax result
      attn_uniform_rows_kernel<<<dim3(a.num_seq, a.heads), 256, 0, s>>>(a);
      if (a.launched) *a.launched = 2;
      return cudaGetLastError();
    }
    if (e != cudaErrorNotSupported) return e;
  }
  if (!a.force_mma_sync) {
    const cudaError_t e = launch_attention_kloop_tcgen05(s, a);   // unmasked, dh = 64, S % 256 == 0: spatial stack / auxiliary encoder
    if (e != cudaErrorNotSupported) return e;
  }
  if (a.S <= 16) {
    const int total = a.num_seq * a.heads;
    if (DHT == 64) attn_small_kernel<64><<<(total + 3) / 4, 128, 0, s>>>(a, total);
    else if (DHT == 32) attn_small_kernel<32><<<(total + 3) / 4, 128, 0, s>>>(a, total);
    else attn_small_kernel<128><<<(total + 1) / 2, 64, 0, s>>>(a, total);
  } else {
    dim3 grid(((a.S + 63) / 64) * a.num_seq, a.heads);
    const size_t smem = static_cast<size_t>(64 + 4 * 64) * DHT * 2 + 2 * 64 * sizeof(float);   // sQ | sK[2] | sV[2] | sFlag[2]
    if (DHT == 64) attn_flash_kernel<64><<<grid, 128, smem, s>>>(a);
    else if (DHT == 32) attn_flash_kernel<32><<<grid, 128, smem, s>>>(a);
    else {
      static int granted[kMaxDevices] = {};
      const cudaError_t e = ensure_dynamic_smem(attn_flash_kernel<128>, static_cast<int>(smem), granted);
      if (e != cudaSuccess) return e;
      attn_flash_kernel<128><<<grid, 128, smem, s>>>(a);
    }
  }
  return cudaGetLastError();
}

}  // namespace vp
