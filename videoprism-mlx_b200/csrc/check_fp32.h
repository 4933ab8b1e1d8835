// fp32 check-mode kernels (check_fp32.cu): internal interface used by the engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace vp {
namespace f32 {

// C[M,N] = epi(A[M,K] . W): v = (acc + bias[n]) * alpha; act (0 none, 1 exact-erf GELU, 2 ReLU); * row_scale[m];
// + pos_table[m % pos_period, n]; + resid[m, n].  w_nk = 0: W is [K, N] (ldw = row pitch); w_nk = 1: W is [N, K].
struct Sgemm {
  const float* A = nullptr; int lda = 0;
  const float* W = nullptr; int ldw = 0; int w_nk = 0;
  float* C = nullptr; int ldc = 0;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;
  float alpha = 1.0f;
  int act = 0;
  const float* row_scale = nullptr;
  const float* pos_table = nullptr; int pos_period = 1;
  const float* resid = nullptr; int ldr = 0;   // may alias C
};
cudaError_t sgemm(cudaStream_t s, const Sgemm& a);

// y[m, :] = LN(x[m, :]) (+ add_table[(m / add_div) % add_mod, :]); y2 (optional, dense [M, D]) = LN(x) without the table
struct LayerNorm {
  const float* x = nullptr; int ldx = 0;
  const float* gamma1 = nullptr; const float* beta = nullptr;   // 1 + scale, bias
  float* y = nullptr; int ldy = 0;                                // may alias x
  float* y2 = nullptr;
  const float* add_table = nullptr; int add_div = 1, add_mod = 1;
  int M = 0, D = 0;
};
cudaError_t layernorm(cudaStream_t s, const LayerNorm& a);

// same row mapping and mask semantics as AttnArgs (kernels.h), fp32 operands
struct Attention {
  const float* q = nullptr; const float* k = nullptr; const float* v = nullptr; int ld = 0;
  float* out = nullptr; int ldo = 0;
  int num_seq = 0, S = 0, group = 1, heads = 0, dh = 0;
  float cap = 0.f;
  const float* key_pad = nullptr;
  int causal = 0;
};
cudaError_t attention(cudaStream_t s, const Attention& a);

cudaError_t patchify(cudaStream_t s, const void* video, int is_u8, float* out /*[tokens, p*p*3]*/, int BT, int H, int W, int p);
cudaError_t text_embed(cudaStream_t s, const int32_t* ids, const float* pad, const float* emb, const float* pe, const float* cls,
                       float* x, float* keep, float* pad_ext, int Q, int L, int D, int vocab);
// xbar[seq, h, :] = sum_s softmax_s(scores[seq, s, h]) x[seq, s, :]
cudaError_t pool(cudaStream_t s, const float* x, const float* scores /*[num_seq*S, H]*/, float* xbar /*[num_seq, H, D]*/, int num_seq,
                 int S, int D, int H);
cudaError_t cast_to_bf16(cudaStream_t s, const float* src, void* dst, size_t n);

}  // namespace f32
}  // namespace vp
