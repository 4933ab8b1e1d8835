"""Frame ingest (video_utils.load_video's resize / centre crop / normalise): the numpy oracle against fixtures made by
cv2 itself and against live cv2 (CPU), and the CUDA kernel against the oracle, bit for bit (GPU)."""
import os

import numpy as np
import pytest

import video_ingest_oracle as VO

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest.npz")


def _cases():
    g = np.load(G)
    out = [("mp4_window center_crop", g["mp4_window"], g["mp4_window_center_crop_288"], 288, "center_crop"),
           ("mp4_window resize", g["mp4_window"], g["mp4_window_resize_288"], 288, "resize")]
    i = 0
    while f"case{i}_src" in g:
        target, mode = (int(v) for v in g[f"case{i}_meta"])
        out.append((f"case{i}", g[f"case{i}_src"], g[f"case{i}_dst"], target, "resize" if mode else "center_crop"))
        i += 1
    return out


def test_oracle_matches_cv2_generated_fixtures_bit_for_bit():
    for tag, src, want, target, mode in _cases():
        got = VO.preprocess_frame_u8(src, target, mode)
        assert got.shape == want.shape and np.array_equal(got, want), tag


def test_oracle_matches_live_cv2_on_random_sizes():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for _ in range(60):
        sh, sw = (int(v) for v in rng.integers(20, 500, 2))
        dh, dw = (int(v) for v in rng.integers(16, 400, 2))
        img = rng.integers(0, 256, (sh, sw, 3)).astype(np.uint8)
        assert np.array_equal(VO.resize_linear_u8(img, dw, dh), cv2.resize(img, (dw, dh))), (sh, sw, dh, dw)
    for sh, sw in ((144, 256), (200, 120), (576, 576), (97, 301)):   # the reference's own helper, incl. the exact-2x shortcut
        img = rng.integers(0, 256, (sh, sw, 3)).astype(np.uint8)
        h, w = img.shape[:2]
        new_h, new_w = (288, int(w * (288 / h))) if h < w else (int(h * (288 / w)), 288)
        ref = cv2.resize(img, (new_w, new_h))
        y0, x0 = (new_h - 288) // 2, (new_w - 288) // 2
        assert np.array_equal(VO.preprocess_frame_u8(img, 288, "center_crop"), ref[y0:y0 + 288, x0:x0 + 288])


REF_MP4 = "/root/reference/videoprism/assets/water_bottle_drumming.mp4"


@pytest.mark.skipif(not os.path.exists(REF_MP4), reason="reference mp4 fixture not on this box")
@pytest.mark.parametrize("mode", ["center_crop", "resize"])
def test_whole_host_chain_equals_the_reference_load_video_on_its_own_fixture(mode):
    """The reference's video_utils.load_video, imported from where it lies and run UNMODIFIED on its own mp4 fixture,
    against this repo's frame sampling (video_utils.read_frames) followed by the ingest oracle: float32 clip, bit for bit.
    (The device resize is held to the same oracle in the GPU tests, so the three agree.)"""
    pytest.importorskip("cv2")
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_video_utils", "/root/reference/videoprism/video_utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    want = ref.load_video(REF_MP4, num_frames=16, target_size=288, resize_mode=mode)
    from videoprism_b200 import video_utils
    frames = video_utils.read_frames(REF_MP4, 16)
    assert frames.dtype == np.uint8 and frames.shape[0] == 16 and frames.shape[-1] == 3
    got = VO.preprocess_frames(frames, 288, mode)
    assert got.dtype == want.dtype == np.float32 and got.shape == want.shape == (16, 288, 288, 3)
    assert np.array_equal(got, want)
    with pytest.raises(ValueError, match="frames"):
        video_utils.read_frames(REF_MP4, 10 ** 6)


def test_normalisation_and_errors():
    frames = np.random.default_rng(0).integers(0, 256, (3, 40, 64, 3)).astype(np.uint8)
    out = VO.preprocess_frames(frames, 36, "center_crop")
    assert out.shape == (3, 36, 36, 3) and out.dtype == np.float32 and 0.0 <= out.min() and out.max() <= 1.0
    with pytest.raises(ValueError, match="Unknown resize_mode"):
        VO.preprocess_frames(frames, 36, "stretch")


@pytest.mark.gpu
def test_gpu_ingest_is_bit_exact_with_the_oracle():
    import torch
    import videoprism_b200 as vp
    for tag, src, want, target, mode in _cases():
        got = vp.video_utils.preprocess_frames(src[None], target, mode)
        assert got.dtype == torch.uint8 and tuple(got.shape) == (1, target, target, 3)
        assert np.array_equal(got[0].cpu().numpy(), want), tag
    rng = np.random.default_rng(5)
    for (t, h, w, target, mode) in [(16, 360, 640, 288, "center_crop"), (4, 576, 1024, 288, "center_crop"), (3, 720, 406, 288, "center_crop"),
                                    (2, 100, 100, 288, "center_crop"), (5, 333, 517, 288, "resize"), (2, 576, 576, 288, "resize"),
                                    (7, 241, 319, 144, "center_crop")]:
        frames = rng.integers(0, 256, (t, h, w, 3)).astype(np.uint8)
        got = vp.video_utils.preprocess_frames(frames, target, mode).cpu().numpy()
        want = np.stack([VO.preprocess_frame_u8(f, target, mode) for f in frames])
        assert np.array_equal(got, want), (t, h, w, target, mode)
        as_float = vp.video_utils.preprocess_frames(torch.from_numpy(frames).cuda(), target, mode).cpu().numpy().astype(np.float32) / 255.0
        assert np.array_equal(as_float, VO.preprocess_frames(frames, target, mode))
    with pytest.raises(ValueError, match="Unknown resize_mode"):
        vp.video_utils.preprocess_frames(frames, 288, "stretch")


@pytest.mark.gpu
def test_gpu_ingest_alignment_and_window_edge_cases():
    """Frames whose base address and row pitch are not aligned (odd widths, a tensor that starts 1..15 bytes into its
    allocation and ends at its very last byte: the 32-bit output stores and the gathers must not assume alignment or read
    past the end), a single frame, strong down-scaling and up-scaling, the identity size: all bit-exact with the oracle."""
    import torch
    import videoprism_b200 as vp
    rng = np.random.default_rng(6)
    for (t, h, w, target, mode) in [(1, 289, 431, 288, "center_crop"), (2, 361, 643, 288, "resize"), (1, 1080, 1920, 288, "center_crop"),
                                    (2, 97, 131, 288, "resize"), (3, 300, 299, 144, "center_crop"), (1, 288, 288, 288, "resize")]:
        frames = rng.integers(0, 256, (t, h, w, 3)).astype(np.uint8)
        want = np.stack([VO.preprocess_frame_u8(f, target, mode) for f in frames])
        for offset in (0, 1, 7, 13):
            flat = torch.zeros(frames.size + offset, dtype=torch.uint8, device="cuda")   # the frames end at the allocation's last byte
            flat[offset:] = torch.from_numpy(frames.reshape(-1)).cuda()
            view = flat[offset:].view(t, h, w, 3)
            assert view.data_ptr() % 16 == (flat.data_ptr() + offset) % 16
            got = vp.video_utils.preprocess_frames(view, target, mode).cpu().numpy()
            assert np.array_equal(got, want), (t, h, w, target, mode, offset)


@pytest.mark.gpu
def test_gpu_ingest_feeds_the_encoder():
    """decoded frames -> device resize -> uint8 encoder entry == the float path on the oracle-preprocessed clip, bitwise."""
    import torch
    import videoprism_b200 as vp
    import videoprism_oracle as O
    cfg = O.CONFIGS["videoprism_public_v1_base"]
    m = vp.get_model("videoprism_public_v1_base")
    m.load_state(O.make_synthetic_weights(cfg))
    frames = np.random.default_rng(9).integers(0, 256, (16, 300, 400, 3)).astype(np.uint8)
    u8 = vp.video_utils.preprocess_frames(frames)[None]                     # [1, 16, 288, 288, 3] uint8 on the device
    feats_u8, _ = m(u8)
    clip = VO.preprocess_frames(frames)[None]                                # what the reference's load_video returns
    feats_f32, _ = m(torch.from_numpy(clip).cuda())
    assert torch.equal(feats_u8, feats_f32)
