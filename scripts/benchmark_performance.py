"""Forward-pass latency of a video-text model on the B200 path, with the command line of the reference's
`scripts/benchmark_performance.py` (same model / video / tokenizer / runs / warm-up options, same report of per-run
wall-clock statistics and peak resident-set size), so the two can be run side by side:

    python scripts/benchmark_performance.py --runs 20 --warmup 3
    python scripts/benchmark_performance.py --video-path clip.mp4 --weights flax_lvt_base_f16r288_repeated.npz \
        --text-tokenizer /path/to/c4_en.model --text "a person drumming on water bottles" --text "a cat"

Every piece that is not available falls back to synthetic data and says so: no video file -> uniform random frames,
no weights file -> seeded random-init weights (no checkpoint is reachable offline), no SentencePiece model -> random
token ids with ragged lengths.  One run = host frames in -> video + text embeddings out (numpy), synchronised.
"""
from __future__ import annotations

import argparse
import os
import resource
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--model-name", default="videoprism_lvt_public_v1_base")
    ap.add_argument("--video-path", default="videoprism/assets/water_bottle_drumming.mp4")
    ap.add_argument("--num-frames", type=int, default=16)
    ap.add_argument("--target-size", type=int, default=288)
    ap.add_argument("--text-tokenizer", default="c4_en", help="tokenizer name (models.TEXT_TOKENIZERS) or a SentencePiece model file")
    ap.add_argument("--text", action="append", default=None, help="query text (repeatable)")
    ap.add_argument("--weights", default=None, help="Flax-layout .npz checkpoint; random-init weights when omitted")
    ap.add_argument("--batch", type=int, default=1, help="copies of the clip per forward")
    ap.add_argument("--runs", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--normalize", action="store_true", help="return l2-normalised embeddings")
    args = ap.parse_args()

    import torch
    import videoprism_b200 as vp

    model = vp.get_model(args.model_name)
    is_clip = isinstance(model, vp.FactorizedVideoCLIP)
    if args.weights:
        state = vp.load_pretrained_weights(args.model_name, checkpoint_path=args.weights)
        print(f"weights: {args.weights}")
    else:
        state = vp.synthetic_state(model, seed=1234)
        print("weights: seeded random init (no --weights given)")
    model.load_state(state)

    if os.path.exists(args.video_path):
        clip = vp.video_utils.load_video(args.video_path, num_frames=args.num_frames, target_size=args.target_size)
        print(f"video: {args.video_path} -> {clip.shape} (resize / centre crop on the device)")
    else:
        clip = np.random.default_rng(0).random((args.num_frames, args.target_size, args.target_size, 3), dtype=np.float32)
        print(f"video: synthetic uniform frames {clip.shape} ({args.video_path} not found)")
    video = np.ascontiguousarray(np.broadcast_to(np.asarray(clip, dtype=np.float32)[None], (args.batch,) + tuple(clip.shape)))

    ids = paddings = None
    if is_clip:
        texts = args.text or ["a person drumming on water bottles", "a dog running on a beach", "someone cooking pasta"]
        try:
            path = args.text_tokenizer
            tok = vp.tokenizers.SentencePieceTokenizer(path) if os.path.isfile(path) else vp.load_text_tokenizer(path)
            ids, paddings = vp.tokenize_texts(tok, texts)
            print(f"text: {len(texts)} queries tokenised with {args.text_tokenizer}")
        except (FileNotFoundError, ValueError) as e:
            rng = np.random.default_rng(2)
            ids = rng.integers(1, model.config["vocabulary_size"], (len(texts), vp.models.TEXT_MAX_LEN)).astype(np.int32)
            lens = rng.integers(4, 33, (len(texts),))
            paddings = (np.arange(vp.models.TEXT_MAX_LEN)[None, :] >= lens[:, None]).astype(np.float32)
            ids = np.where(paddings > 0, 0, ids).astype(np.int32)
            print(f"text: synthetic token ids for {len(texts)} queries ({type(e).__name__}: tokenizer model not available)")

    def run_once():
        if is_clip:
            out = model(video, ids, paddings, normalize=args.normalize)
        else:
            out = model(video)
        torch.cuda.synchronize()
        return out

    for _ in range(args.warmup):
        run_once()
    times = []
    for _ in range(args.runs):
        t0 = time.perf_counter()
        out = run_once()
        times.append(time.perf_counter() - t0)
    mean = statistics.mean(times)
    std = statistics.pstdev(times) if len(times) > 1 else 0.0
    print(f"\n=== B200 ({torch.cuda.get_device_name()}) {args.model_name}, batch {args.batch} ===")
    print(f"mean={mean:.4f}s  std={std:.4f}s  min={min(times):.4f}s  max={max(times):.4f}s  ({args.batch / mean:.1f} clips/s, "
          f"{model.kernel_launches // (args.runs + args.warmup)} kernel launches per run)")
    print(f"peak RSS: {resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024 ** 2:.2f} GiB")
    if is_clip:
        v, t = np.asarray(out[0]), np.asarray(out[1])
        print(f"video embeddings {v.shape}, text embeddings {t.shape}; similarity of clip 0 to the queries: "
              f"{np.array2string(v[0] @ t.T, precision=4)}")
    else:
        print(f"features {np.asarray(out[0]).shape}")


if __name__ == "__main__":
    main()
