"""Generates tests/golden/ingest.npz with cv2 ITSELF (the arithmetic the reference's video_utils.load_video delegates to).

    python tests/golden/make_golden_ingest.py          # needs cv2 and /root/reference (this container only)

Contents: a window of a frame of the reference's own fixture (assets/water_bottle_drumming.mp4) and random images,
each with the output of the reference's `_center_crop_resize` / plain `cv2.resize` for several source sizes
(down-scaling, the exact-2x INTER_AREA shortcut, up-scaling, non-square crops).  Small on purpose (< 1 MB).
"""
import os
import sys

import cv2
import numpy as np

REF = "/root/reference"
sys.path.insert(0, REF)
from videoprism import video_utils as ref_utils   # noqa: E402  (pure numpy + cv2 part of the reference)

out = {}
cap = cv2.VideoCapture(os.path.join(REF, "videoprism", "assets", "water_bottle_drumming.mp4"))
cap.set(cv2.CAP_PROP_POS_FRAMES, 57)
ok, fr = cap.read()
assert ok
cap.release()
real = cv2.cvtColor(fr, cv2.COLOR_BGR2RGB)[100:400, 60:460]                      # a 300 x 400 window of a real frame
out["mp4_window"] = real
out["mp4_window_center_crop_288"] = ref_utils._center_crop_resize(real, 288)
out["mp4_window_resize_288"] = cv2.resize(real, (288, 288))
rng = np.random.default_rng(7)
cases = [(90, 160, 72, "center_crop"), (180, 101, 72, "center_crop"), (144, 144, 72, "center_crop"), (45, 60, 72, "center_crop"),
         (111, 173, 72, "resize"), (125, 125, 36, "center_crop"), (64, 48, 72, "resize")]
for i, (h, w, target, mode) in enumerate(cases):
    img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    res = ref_utils._center_crop_resize(img, target) if mode == "center_crop" else cv2.resize(img, (target, target))
    out[f"case{i}_src"] = img
    out[f"case{i}_dst"] = res
    out[f"case{i}_meta"] = np.array([target, 0 if mode == "center_crop" else 1])
out["cv2_version"] = np.array(cv2.__version__)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ingest.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
