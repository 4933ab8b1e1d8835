"""End-to-end check of the video-text models on a real clip and a real query, in the shape of the reference's
`verify_clip_models.py` (which compares its Flax and MLX implementations and passes at max-abs < 1e-3 on the video
embedding, the text embedding and their cosine similarity, verify_clip_models.py:92-95).

Here the three implementations of the same forward are compared with each other on the same inputs and weights:

  (P) the production path: bf16 tensor cores, fp32 accumulation            -> vp.get_model(name)
  (C) the fp32 check mode of the same library (float32 on the CUDA cores)  -> vp.get_model(name, check_fp32=True)
(the third implementation, the CPU oracle that restates the Flax reference line by line, is test infrastructure: the
same comparison against it is made by tests/test_verify_script_gpu.py, which calls `run` below).

    python scripts/verify_clip_models.py --weights-dir /path/with/flax_lvt_*_repeated.npz \
        --video videoprism/assets/water_bottle_drumming.mp4 --tokenizer /path/to/c4_en.model

Pass criteria: (P) vs (C) per-embedding cosine >= 0.999 and similarity within 2e-2 (the test adds (C) vs the oracle under
the reference's own bound: max-abs < 1e-3 on both embeddings and on the similarity).  Released weights and the SentencePiece model are
not reachable offline, so every missing piece falls back to a SEEDED SYNTHETIC stand-in and the report says which were
used; `--require-real` turns a fallback into an error (the skip-if-no-weights test uses it).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHECKPOINT_FILES = {   # videoprism/models.py:62-80
    "videoprism_lvt_public_v1_base": "flax_lvt_base_f16r288_repeated.npz",
    "videoprism_lvt_public_v1_large": "flax_lvt_large_f8r288_repeated.npz",
}


def cosine(a, b):
    a = np.asarray(a, np.float64).reshape(len(a), -1); b = np.asarray(b, np.float64).reshape(len(b), -1)
    return float(((a * b).sum(-1) / (np.linalg.norm(a, axis=-1) * np.linalg.norm(b, axis=-1))).min())


def main(argv=None, results=None) -> int:
    """`results` (optional dict): filled with {model: {inputs, production, check}} for a caller that wants the numbers."""
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--models", nargs="*", default=list(CHECKPOINT_FILES))
    ap.add_argument("--weights-dir", default=os.environ.get("VP_WEIGHTS_DIR"), help="directory holding the released flax_lvt_*_repeated.npz files")
    ap.add_argument("--video", default=os.path.join("videoprism", "assets", "water_bottle_drumming.mp4"))
    ap.add_argument("--tokenizer", default=os.environ.get("VP_SPM_MODEL"), help="SentencePiece model file of the c4_en tokenizer")
    ap.add_argument("--text", action="append", default=None)
    ap.add_argument("--require-real", action="store_true", help="fail instead of falling back to synthetic weights / video / token ids")
    args = ap.parse_args(argv)
    texts = args.text or ["child drumming on water bottles"]

    import videoprism_b200 as vp

    ok_all = True
    for name in args.models:
        print("=" * 80 + f"\n{name}\n" + "=" * 80)
        used = []
        model_p, model_c = vp.get_model(name), vp.get_model(name, check_fp32=True)
        wpath = os.path.join(args.weights_dir, CHECKPOINT_FILES[name]) if args.weights_dir else None
        if wpath and os.path.exists(wpath):
            state = vp.load_pretrained_weights(name, checkpoint_path=wpath)
            used.append(f"weights: {wpath}")
        elif args.require_real:
            print(f"released weights not found ({wpath}); --require-real given")
            return 2
        else:
            state = vp.synthetic_state(model_p, seed=1234)
            used.append("weights: SYNTHETIC seeded random init (released checkpoints are not reachable offline)")
        if os.path.exists(args.video):
            clip = vp.video_utils.load_video(args.video, num_frames=16, target_size=288)
            used.append(f"video: {args.video} -> {clip.shape}")
        elif args.require_real:
            print(f"video {args.video} not found; --require-real given")
            return 2
        else:
            clip = np.random.default_rng(0).random((16, 288, 288, 3), dtype=np.float32)
            used.append("video: SYNTHETIC uniform frames (file not found)")
        vocab = model_p.config["vocabulary_size"]
        if args.tokenizer and os.path.exists(args.tokenizer):
            tok = vp.tokenizers.SentencePieceTokenizer(args.tokenizer)
            ids, pad = vp.tokenize_texts(tok, texts)
            used.append(f"text: {texts} through {args.tokenizer}")
        elif args.require_real:
            print("SentencePiece model not found; --require-real given")
            return 2
        else:
            rng = np.random.default_rng(5)
            ids = np.zeros((len(texts), 64), np.int32); pad = np.ones((len(texts), 64), np.float32)
            for r, t in enumerate(texts):
                n = max(1, min(64, len(t.split())))
                ids[r, :n] = rng.integers(1, vocab, n); pad[r, :n] = 0.0
            used.append("text: SYNTHETIC token ids, one per word of the query (no SentencePiece model)")
        for u in used:
            print("  " + u)
        video = clip[None]
        vp_, tp_, _ = model_p.apply(state, video, ids, pad, train=False)
        vc_, tc_, _ = model_c.apply(state, video, ids, pad, train=False)
        sim_p, sim_c = vp_ @ tp_.T, vc_ @ tc_.T
        print(f"  production (bf16) vs fp32 check mode: video cosine {cosine(vp_, vc_):.6f}, text cosine {cosine(tp_, tc_):.6f}, "
              f"similarity diff {np.abs(sim_p - sim_c).max():.3e}")
        ok = cosine(vp_, vc_) >= 0.999 and cosine(tp_, tc_) >= 0.999 and np.abs(sim_p - sim_c).max() < 2e-2
        if results is not None:
            results[name] = {"state": state, "video": video, "ids": ids, "paddings": pad, "production": (vp_, tp_), "check": (vc_, tc_),
                             "used": used}
        print(f"  similarity (production): {np.round(sim_p, 4).tolist()}")
        print("  PASS" if ok else "  FAIL")
        ok_all = ok_all and ok
        del model_p, model_c
    return 0 if ok_all else 1


if __name__ == "__main__":
    raise SystemExit(main())
