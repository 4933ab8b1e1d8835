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

// One block per (frame, output row): the row's vertical tap is computed once, the output row is assembled in shared
// memory and written with coalesced 32-bit stores (byte stores and 288 redundant copies of the row tap made the first
// version LSU-bound at 0.7 TB/s).
__global__ void resize_frames_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, int new_h,
                                        int new_w, int y0, int x0, int target) {
  extern __shared__ __align__(16) uint8_t row_out[];   // target * 3 bytes (padded to a multiple of 4)
  __shared__ Tap ty_s;
  const int y = static_cast<int>(blockIdx.x) % target;
  const int t = static_cast<int>(blockIdx.x) / target;
  const uint8_t* frame = src + static_cast<size_t>(t) * H * W * 3;
  const int dy = y + y0;   // coordinates in the (virtual) resized image
  const bool area2x = (W == 2 * new_w && H == 2 * new_h);
  if (threadIdx.x == 0 && !area2x) ty_s = linear_tap(dy, H, new_h, true);
  __syncthreads();
  for (int x = threadIdx.x; x < target; x += blockDim.x) {
    const int dx = x + x0;
    uint8_t* o = row_out + x * 3;
    if (area2x) {
      const uint8_t* p0 = frame + (static_cast<size_t>(2 * dy) * W + 2 * dx) * 3;
      const uint8_t* p1 = p0 + static_cast<size_t>(W) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) o[c] = static_cast<uint8_t>((p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2);
    } else {
      const Tap tx = linear_tap(dx, W, new_w, false);
      const Tap ty = ty_s;
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
  const size_t row_bytes = static_cast<size_t>(target) * 3;
  uint8_t* drow = dst + (static_cast<size_t>(t) * target + y) * row_bytes;
  if ((reinterpret_cast<uintptr_t>(drow) & 3) == 0 && (row_bytes & 3) == 0) {
    for (int i = threadIdx.x; i < static_cast<int>(row_bytes / 4); i += blockDim.x)
      reinterpret_cast<uint32_t*>(drow)[i] = reinterpret_cast<const uint32_t*>(row_out)[i];
  } else {
    for (int i = threadIdx.x; i < static_cast<int>(row_bytes); i += blockDim.x) drow[i] = row_out[i];
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
  const int threads = target >= 512 ? 512 : ((target + 31) / 32) * 32;
  const size_t smem = (static_cast<size_t>(target) * 3 + 15) / 16 * 16;
  if (smem > 48 * 1024 || static_cast<size_t>(T) * target > 0x7fffffffULL) return cudaErrorInvalidValue;
  resize_frames_u8_kernel<<<static_cast<unsigned>(T) * target, threads, smem, s>>>(src, dst, H, W, new_h, new_w, y0, x0, target);
  return cudaGetLastError();
}

}  // namespace vp
