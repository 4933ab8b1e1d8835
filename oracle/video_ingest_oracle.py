"""CPU restatement (numpy) of the reference's frame preprocessing: TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py` may import this file; the product path
(`videoprism_b200.video_utils`) never does.

What it restates
----------------
`videoprism/video_utils.py:20-94` (`load_video`): per sampled frame BGR->RGB, then either
`_center_crop_resize` (`video_utils.py:97-127`: shortest side -> target_size with `cv2.resize`, then a centre
crop) or a plain `cv2.resize(frame, (target, target))`, finally `astype(float32) / 255.0`.

The arithmetic lives in a third-party dependency that is not under /root/reference: **opencv-python**
(`requirements.txt:7`, unpinned; 4.13.0 in this image).  `cv2.resize` with its default `INTER_LINEAR` on uint8 is
OpenCV's fixed-point bilinear (imgproc/resize.cpp, `HResizeLinear<uchar,int,short,2048>` + `VResizeLinear<uchar,...>`):

  scale   = 1 / (dst / src)                                   (double)
  f       = float((d + 0.5) * scale - 0.5);  s = floor(f);  f -= s
  columns : where the 2-tap window leaves the image (s < 0 or s >= src-1) the fraction is ZEROED and s clamped
  rows    : the fraction is KEPT and the two row indices are clamped
  weights : a0 = cvRound((1 - f) * 2048), a1 = cvRound(f * 2048)          (round half to even; int16)
  h-pass  : r[y][x] = S[y][s0] * a0 + S[y][s1] * a1                        (int32, scaled by 2^11)
  v-pass  : dst = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
  and an EXACT 2x shrink in both directions is computed as INTER_AREA: (p00 + p01 + p10 + p11 + 2) >> 2.

Pinned: `tests/test_video_ingest.py` checks this restatement bit for bit against fixtures produced by cv2 itself
(`tests/golden/make_golden_ingest.py`, frames of the reference's own `assets/water_bottle_drumming.mp4` plus random
images over up- and down-scaling ratios) and, where cv2 is importable, against live `cv2.resize` calls on random sizes.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def linear_coeffs(src: int, dst: int, vertical: bool):
    """(index0, index1, weight0, weight1) of every destination coordinate (resize.cpp: the xofs/ialpha, yofs/ibeta tables)."""
    scale = 1.0 / (float(dst) / float(src))
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int32)
    w1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
    s1 = np.clip(s + 1, 0, src - 1)
    s0 = np.clip(s, 0, src - 1)
    return s0.astype(np.int32), s1.astype(np.int32), w0, w1


def resize_linear_u8(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv2.resize(img, (dst_w, dst_h)) for uint8 HxWxC images (default INTER_LINEAR)."""
    assert img.dtype == np.uint8 and img.ndim == 3
    sh, sw = img.shape[:2]
    x = img.astype(np.int32)
    if sw == 2 * dst_w and sh == 2 * dst_h:
        return ((x[0::2, 0::2] + x[0::2, 1::2] + x[1::2, 0::2] + x[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx0, sx1, a0, a1 = linear_coeffs(sw, dst_w, vertical=False)
    sy0, sy1, b0, b1 = linear_coeffs(sh, dst_h, vertical=True)
    rows = x[:, sx0] * a0[None, :, None] + x[:, sx1] * a1[None, :, None]
    r0, r1 = rows[sy0] >> 4, rows[sy1] >> 4
    out = (((b0[:, None, None] * r0) >> 16) + ((b1[:, None, None] * r1) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def resized_shape(h: int, w: int, target: int, resize_mode: str):
    """(new_h, new_w, start_y, start_x) of video_utils.py:75-82 / :108-125."""
    if resize_mode == "resize":
        return target, target, 0, 0
    if resize_mode != "center_crop":
        raise ValueError(f"Unknown resize_mode: {resize_mode}")
    if h < w:
        new_h, new_w = target, int(w * (target / h))
    else:
        new_w, new_h = target, int(h * (target / w))
    return new_h, new_w, (new_h - target) // 2, (new_w - target) // 2


def preprocess_frame_u8(frame: np.ndarray, target: int = 288, resize_mode: str = "center_crop") -> np.ndarray:
    """One RGB uint8 frame -> uint8 [target, target, 3] (video_utils.py:74-84 after the colour conversion)."""
    h, w = frame.shape[:2]
    new_h, new_w, y0, x0 = resized_shape(h, w, target, resize_mode)
    out = resize_linear_u8(frame, new_w, new_h)
    return out[y0:y0 + target, x0:x0 + target]


def preprocess_frames(frames: np.ndarray, target: int = 288, resize_mode: str = "center_crop") -> np.ndarray:
    """uint8 [T, H, W, 3] RGB -> float32 [T, target, target, 3] in [0, 1] (video_utils.py:89-93)."""
    out = np.stack([preprocess_frame_u8(f, target, resize_mode) for f in frames], axis=0)
    return out.astype(np.float32) / 255.0
