"""Asynchronous host-buffer entry points (vp_encoder_forward_host_async / vp_clip_video_forward_host_async / vp_wait):
several calls in flight, different batch sizes, float32 / uint8 frames, bfloat16 features; every result must be exactly what
the blocking call returns (the same kernels run on the same chunks; only the overlap of copies and forwards differs)."""
import numpy as np
import pytest
import torch

import videoprism_oracle as O

pytestmark = pytest.mark.gpu


def _encoder():
    import videoprism_b200 as vp
    cfg = O.tiny_config("encoder")
    m = vp.FactorizedEncoder(**{k: v for k, v in cfg.items() if k != "kind"})
    m.load_state(O.make_synthetic_weights(cfg))
    return m


def test_three_calls_in_flight_with_different_batch_sizes_match_blocking_calls():
    import videoprism_b200 as vp
    m = _encoder()
    clips = [O.make_video(b, 8, 32, seed=40 + b, kind="normal") for b in (2, 9, 1, 9, 3)]   # 9 clips: larger chunks, the slots move
    want = [m(c)[0].copy() for c in clips]                                                     # blocking calls, one at a time
    m = _encoder()             # a fresh handle: its staging slots start small and have to grow while calls are in flight
    pinned = []
    for c in clips:
        p = vp.pinned_empty(c.shape)
        p[...] = c
        pinned.append(p)
    tickets, outs = [], []
    for p in pinned[:3]:                                   # three calls in flight
        t, o, _ = m.forward_async(p)
        tickets.append(t); outs.append(o)
    for p in pinned[3:]:                                   # ... and two more behind them before anything is waited for
        t, o, _ = m.forward_async(p)
        tickets.append(t); outs.append(o)
    for t in reversed(tickets):                            # waiting out of order is allowed
        m.wait(t)
    for i, (o, w) in enumerate(zip(outs, want)):
        assert np.array_equal(o, w), f"call {i} (batch {clips[i].shape[0]}) differs from the blocking call"


def test_async_uint8_frames_frame_paddings_and_bf16_features():
    m = _encoder()
    rng = np.random.default_rng(7)
    u8 = rng.integers(0, 256, (3, 8, 32, 32, 3), dtype=np.uint8)
    fp = np.zeros((3, 8), np.float32); fp[1, 5:] = 1.0; fp[2, :] = 1.0
    want, _ = m(u8.astype(np.float32) / np.float32(255.0), frame_paddings=fp)
    t, got, _ = m.forward_async(u8, frame_paddings=fp)
    m.wait(t)
    assert np.array_equal(got, want)                       # uint8 ingest is bitwise the float path
    t, got16, _ = m.forward_async(u8, frame_paddings=fp, bf16_features=True)
    m.wait(t)
    assert got16.dtype == np.uint16
    as_f32 = torch.from_numpy(got16.view(np.int16)).view(torch.bfloat16).float().numpy()
    assert np.abs(as_f32 - want).max() <= 2.0 ** -8 * np.abs(want).max() + 1e-6     # one bfloat16 rounding of the same features


def test_async_video_text_embeddings_match_the_blocking_call():
    import videoprism_b200 as vp
    cfg = O.tiny_config("clip")
    W = O.make_synthetic_weights(cfg)
    m = vp.FactorizedVideoCLIP(**{k: v for k, v in cfg.items() if k != "kind"})
    m.load_state(W)
    a = O.make_video(5, 4, 16, seed=51, kind="normal")
    b = O.make_video(2, 4, 16, seed=52, kind="normal")
    t, want_a = m.embed_video_async(a); m.wait(t)          # one at a time (same chunk schedule as below: 5 clips in one chunk)
    t, want_b = m.embed_video_async(b); m.wait(t)
    blocking, _, _ = m(a)                                   # the blocking call chunks 2 / 1 / 2: same numbers up to bf16-level noise
    assert np.abs(blocking - want_a).max() < 2e-3
    ta, got_a = m.embed_video_async(a)
    tb, got_b = m.embed_video_async(b)                     # second call enqueued while the first is in flight
    m.wait(tb); m.wait(ta)
    assert np.array_equal(got_a, want_a) and np.array_equal(got_b, want_b)
    fp = np.zeros((5, 4), np.float32); fp[0, 2:] = 1.0
    dev, _, _ = m(torch.from_numpy(a).cuda(), frame_paddings=torch.from_numpy(fp).cuda())
    t, host = m.embed_video_async(a, frame_paddings=fp)    # host path with frame paddings (new this round)
    m.wait(t)
    assert np.abs(host - dev.cpu().numpy()).max() < 2e-3
