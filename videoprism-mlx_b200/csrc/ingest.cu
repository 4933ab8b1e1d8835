// Frame ingest on the device: uint8 RGB frames of any size -> uint8 [T, target, target, 3], bit-exact with what the
// reference's video_utils.load_video gets from OpenCV (video_utils.py:74-84, :97-127):
//   "center_crop": cv2.resize so that the shortest side becomes `target`, then a centre crop;
//   "resize"     : cv2.resize(frame, (target, target)).
// cv2.resize's default INTER_LINEAR on uint8 is a fixed-point bilinear (imgproc/resize.cpp): 11-bit weights
// cvRound((1-f)*2048), cvRound(f*2048) from f = float((d+0.5)*scale - 0.5); along x the fraction is zeroed where the
// 2-tap window leaves the image, along y the row indices are clamped instead; horizontal pass in int32, vertical pass
// (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2; an exact 2x shrink in both directions is INTER_AREA,
// (p00+p01+p10+p11+2)>>2.  Integer arithmetic throughout, so the result is reproduced exactly; the coordinate
// arithmetic uses explicitly rounded double operations (no FMA contraction) to match the host library.
// Only the cropped window of the resized image is ever computed.  HBM-bound and tiny (8 MB read + 4 MB written per
// 16-frame 640x360 clip).
#include "kernels.h"

namespace vp {

namespace {

struct Tap {
  int s0, s1, w0, w1;
};

__device__ __forceinline__ Tap linear_tap(int d, int src, int dst, bool vertical) {
  const double scale = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(dst), static_cast<double>(src)));
  float f = static_cast<float>(__dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(d), 0.5), scale), -0.5));
  int s = static_cast<int>(floorf(f));
  f = __fsub_rn(f, static_cast<float>(s));
  if (!vertical) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  Tap t;
  t.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));   // cvRound: round half to even
  t.w1 = __float2int_rn(__fmul_rn(f, 2048.0f));
  t.s1 = min(max(s + 1, 0), src - 1);
  t.s0 = min(max(s, 0), src - 1);
  return t;
}

// One block per (frame, group of kRows output rows).  The horizontal taps depend only on x and the vertical ones only on
// y, so a block computes them once into shared memory (they cost a handful of double-precision operations each; the first
// version recomputed both for every pixel), assembles its kRows output rows there and writes them as one contiguous run of
// 32-bit words (the rows of a group are adjacent in the output).
// Tried and dropped: staging each block's source window in shared memory with 16-byte loads (wide, coalesced DRAM reads,
// gathers from shared memory) is bit-exact but slower (73 us against 58 us for 128 frames of 640x360): the kernel is bound
// by the ~80 integer / load instructions per output pixel (12 byte gathers, two fixed-point passes), not by memory.
constexpr int kRows = 8;

__global__ void __launch_bounds__(256) resize_frames_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W,
                                                               int new_h, int new_w, int y0, int x0, int target, int groups) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Tap* tx_s = reinterpret_cast<Tap*>(smem_raw);                       // [target]
  uint8_t* rows_out = smem_raw + static_cast<size_t>(target) * sizeof(Tap);   // [kRows][target * 3]
  __shared__ Tap ty_s[kRows];
  const int g = static_cast<int>(blockIdx.x) % groups;
  const int t = static_cast<int>(blockIdx.x) / groups;
  const int ybase = g * kRows;
  const int nrows = min(kRows, target - ybase);
  const uint8_t* frame = src + static_cast<size_t>(t) * H * W * 3;
  const bool area2x = (W == 2 * new_w && H == 2 * new_h);
  if (!area2x) {
    for (int x = threadIdx.x; x < target; x += blockDim.x) tx_s[x] = linear_tap(x + x0, W, new_w, false);
    if (threadIdx.x < nrows) ty_s[threadIdx.x] = linear_tap(ybase + threadIdx.x + y0, H, new_h, true);
  }
  __syncthreads();
  const int row_bytes = target * 3;
  for (int i = threadIdx.x; i < nrows * target; i += blockDim.x) {
    const int r = i / target, x = i - r * target;
    uint8_t* o = rows_out + r * row_bytes + x * 3;
    if (area2x) {
      const uint8_t* p0 = frame + (static_cast<size_t>(2 * (ybase + r + y0)) * W + 2 * (x + x0)) * 3;
      const uint8_t* p1 = p0 + static_cast<size_t>(W) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = static_cast<uint8_t>((p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2);
    } else {
      const Tap tx = tx_s[x];
      const Tap ty = ty_s[r];
      const uint8_t* r0 = frame + static_cast<size_t>(ty.s0) * W * 3;
      const uint8_t* r1 = frame + static_cast<size_t>(ty.s1) * W * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int h0 = r0[tx.s0 * 3 + c] * tx.w0 + r0[tx.s1 * 3 + c] * tx.w1;
        const int h1 = r1[tx.s0 * 3 + c] * tx.w0 + r1[tx.s1 * 3 + c] * tx.w1;
        o[c] = static_cast<uint8_t>((((ty.w0 * (h0 >> 4)) >> 16) + ((ty.w1 * (h1 >> 4)) >> 16) + 2) >> 2);
      }
    }
  }
  __syncthreads();
  const int nbytes = nrows * row_bytes;
  uint8_t* drow = dst + (static_cast<size_t>(t) * target + ybase) * row_bytes;
  if ((reinterpret_cast<uintptr_t>(drow) & 3) == 0 && (nbytes & 3) == 0) {
    for (int i = threadIdx.x; i < nbytes / 4; i += blockDim.x)
      reinterpret_cast<uint32_t*>(drow)[i] = reinterpret_cast<const uint32_t*>(rows_out)[i];
  } else {
    for (int i = threadIdx.x; i < nbytes; i += blockDim.x) drow[i] = rows_out[i];
  }
}

}  // namespace

cudaError_t launch_resize_frames_u8(cudaStream_t s, const uint8_t* src, int T, int H, int W, uint8_t* dst, int target, int mode) {
  if (T <= 0) return cudaSuccess;
  if (H <= 0 || W <= 0 || target <= 0 || (mode != 0 && mode != 1)) return cudaErrorInvalidValue;
  int new_h = target, new_w = target, y0 = 0, x0 = 0;
  if (mode == 0) {   // video_utils.py:108-125, in the same double arithmetic as the Python expressions
    if (H < W) new_w = static_cast<int>(W * (static_cast<double>(target) / H));
    else new_h = static_cast<int>(H * (static_cast<double>(target) / W));
    y0 = (new_h - target) / 2;
    x0 = (new_w - target) / 2;
  }
  const int groups = (target + kRows - 1) / kRows;
  const size_t smem = static_cast<size_t>(target) * sizeof(Tap) + (static_cast<size_t>(kRows) * target * 3 + 15) / 16 * 16;
  if (smem > 48 * 1024 || static_cast<size_t>(T) * groups > 0x7fffffffULL) return cudaErrorInvalidValue;
  resize_frames_u8_kernel<<<static_cast<unsigned>(T) * groups, 256, smem, s>>>(src, dst, H, W, new_h, new_w, y0, x0, target, groups);
  return cudaGetLastError();
}

}  // namespace vp
