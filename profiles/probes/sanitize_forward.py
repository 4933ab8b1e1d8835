"""One pass over every kernel family, written for compute-sanitizer (memcheck): tiny encoder with frame paddings, tiny
video-text model with ragged text, tiny classifier, one full-size base clip (tcgen05 GEMM pair path, S=256 tcgen05
attention, folded LayerNorm), one full-size video-text clip (S=4096 key-loop attention, pooler), frame ingest.

    compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/probes/sanitize_forward.py

compute-sanitizer is closed on this GPU pool (gpurun refuses it: runs under it have left GPUs needing a reset), so in
round 1 this ran plainly, as an all-kernels smoke; out-of-bounds protection rests on the alignment / buffer-end / ragged
shape tests in tests/ (test_gpu_ingest_alignment_and_window_edge_cases, test_base_encoder_ragged_shapes, test_attention).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import torch

import videoprism_b200 as vp
import videoprism_oracle as O

rng = np.random.default_rng(0)
cfg = O.tiny_config("encoder")
m = vp.FactorizedEncoder(**{k: v for k, v in cfg.items() if k != "kind"})
fp = np.zeros((3, 4), np.float32); fp[1, 2:] = 1
out, outs = m.apply(O.make_synthetic_weights(cfg), O.make_video(3, 4, 16, seed=1, kind="normal"), train=False,
                    return_intermediate=True, frame_paddings=fp)
print("tiny encoder", out.shape, float(np.abs(out).max()))
cfg = O.tiny_config("clip")
m = vp.FactorizedVideoCLIP(**{k: v for k, v in cfg.items() if k != "kind"})
ids, pad = O.make_text(5, vocab=cfg["vocabulary_size"], max_len=8)
v, t, outs = m.apply(O.make_synthetic_weights(cfg), O.make_video(2, 4, 16, seed=2, kind="normal"), ids, pad, train=False,
                     return_intermediate=True)
print("tiny clip", v.shape, t.shape, sorted(outs))
cfg = O.tiny_config("classifier")
m = vp.FactorizedVideoClassifier(encoder_params={k: v for k, v in cfg.items() if k not in ("kind", "num_classes")}, num_classes=10)
lg, _ = m.apply(O.make_synthetic_weights(cfg), O.make_video(2, 4, 16, seed=3, kind="normal"), train=False)
print("tiny classifier", lg.shape)
if "--tiny-only" not in sys.argv:
    m = vp.get_model("videoprism_public_v1_base")
    m.load_state(vp.synthetic_state(m))
    vid = rng.random((2, 16, 288, 288, 3), dtype=np.float32)
    fp = np.zeros((2, 16), np.float32); fp[1, 10:] = 1
    out, _ = m(vid, frame_paddings=fp)
    out2, _ = m(torch.from_numpy((vid * 255).astype(np.uint8)).cuda())
    print("base encoder", out.shape, bool(np.isfinite(out).all()), out2.shape)
    del m
    m = vp.get_model("videoprism_lvt_public_v1_base")
    m.load_state(vp.synthetic_state(m))
    ids, pad = O.make_text(4)
    v, t, _ = m(vid[:1], ids, pad)
    print("lvt base", v.shape, t.shape, float((v @ t.T).max()))
    fr = torch.randint(0, 256, (4, 360, 640, 3), dtype=torch.uint8, device="cuda")
    u8 = vp.video_utils.preprocess_frames(fr)
    u8r = vp.video_utils.preprocess_frames(fr[:, :300, :301], 288, "resize")
    print("ingest", tuple(u8.shape), tuple(u8r.shape))
torch.cuda.synchronize()
print("done")
