"""Drop-in for `videoprism/video_utils.py` (load_video :20-94, _center_crop_resize :97-127, load_video_batch :130-152)
with the per-frame resize / centre crop / normalisation on the GPU.

Decoding stays on the host (cv2.VideoCapture, as in the reference); the decoded uint8 frames go to the device once
(1 byte per sample), are resized there by `vp_resize_frames_u8` bit-exactly as `cv2.resize` would, and can be fed to
the encoder as uint8 (`model(frames_u8)` normalises on the device) without ever materialising the float32 clip on the
host.  `load_video` itself returns what the reference returns: float32 `[T, S, S, 3]` in [0, 1].
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_MODES = {"center_crop": 0, "resize": 1}


def preprocess_frames(frames, target_size: int = 288, resize_mode: str = "center_crop"):
    """uint8 RGB frames `[T, H, W, 3]` (numpy, or a CUDA torch tensor) -> CUDA uint8 tensor `[T, S, S, 3]`.

    Equals, bit for bit, `np.stack([_center_crop_resize(f, S) for f in frames])` (or `cv2.resize(f, (S, S))` for
    `resize_mode="resize"`) of the reference."""
    import torch
    if resize_mode not in _MODES:
        raise ValueError(f"Unknown resize_mode: {resize_mode}")   # video_utils.py:83-84
    if isinstance(frames, np.ndarray):
        if frames.dtype != np.uint8:
            raise ValueError("frames must be uint8 (decoded RGB)")
        frames = torch.from_numpy(np.ascontiguousarray(frames)).cuda()
    if frames.dtype != torch.uint8 or frames.ndim != 4 or frames.shape[-1] != 3 or not frames.is_cuda:
        raise ValueError("frames must be a uint8 [T, H, W, 3] array")
    frames = frames.contiguous()
    t, h, w, _ = (int(s) for s in frames.shape)
    out = torch.empty((t, target_size, target_size, 3), dtype=torch.uint8, device=frames.device)
    with torch.cuda.device(frames.device):
        stream = int(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().vp_resize_frames_u8(frames.data_ptr(), t, h, w, out.data_ptr(), int(target_size),
                                                  _MODES[resize_mode], stream))
    return out


def read_frames(video_path: str, num_frames: int = 16) -> np.ndarray:
    """The host part of load_video (video_utils.py:44-79): uniformly sampled RGB uint8 frames `[T, H, W, 3]`."""
    try:
        import cv2
    except ImportError as e:
        raise ImportError("OpenCV is required for video loading. Install it with: pip install opencv-python") from e
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        raise ValueError(f"Could not open video file: {video_path}")
    total_frames = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    if total_frames < num_frames:
        raise ValueError(f"Video has only {total_frames} frames, but {num_frames} requested")
    frames = []
    for frame_idx in np.linspace(0, total_frames - 1, num_frames, dtype=int):
        cap.set(cv2.CAP_PROP_POS_FRAMES, frame_idx)
        ret, frame = cap.read()
        if not ret:
            raise ValueError(f"Could not read frame {frame_idx} from {video_path}")
        frames.append(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
    cap.release()
    return np.stack(frames, axis=0)


def load_video(video_path: str, num_frames: int = 16, target_size: int = 288, resize_mode: str = "center_crop") -> np.ndarray:
    """Same signature and result as the reference's load_video: float32 `[num_frames, S, S, 3]` in [0.0, 1.0]."""
    if resize_mode not in _MODES:
        raise ValueError(f"Unknown resize_mode: {resize_mode}")
    frames = preprocess_frames(read_frames(video_path, num_frames), target_size, resize_mode)
    # the reference's own expression (video_utils.py:91) on the 1-byte-per-sample result: an IEEE float32 division, which a
    # reciprocal multiply on the device would not reproduce bit for bit
    return frames.cpu().numpy().astype(np.float32) / 255.0


def load_video_u8(video_path: str, num_frames: int = 16, target_size: int = 288, resize_mode: str = "center_crop"):
    """load_video without the float32 detour: a CUDA uint8 `[num_frames, S, S, 3]` tensor, to be passed (with a leading
    batch axis) to the encoder, which normalises on the device."""
    return preprocess_frames(read_frames(video_path, num_frames), target_size, resize_mode)


def load_video_batch(video_paths, num_frames: int = 16, target_size: int = 288, resize_mode: str = "center_crop") -> np.ndarray:
    """float32 `[B, num_frames, S, S, 3]` (video_utils.py:130-152)."""
    return np.stack([load_video(p, num_frames, target_size, resize_mode) for p in video_paths], axis=0)
