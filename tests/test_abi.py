"""CPU checks of the C-ABI boundary: the library builds, loads, exports every symbol the header declares,
and fails loudly (no fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "videoprism_b200.h")


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"VP_API[^;(]*?\b(vp_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_surface():
    names = _declared()
    for required in ("vp_create", "vp_set_weight", "vp_finalize", "vp_encoder_forward", "vp_encoder_forward_host",
                     "vp_clip_video_forward", "vp_clip_text_forward", "vp_gemm_bf16", "vp_attention", "vp_destroy"):
        assert required in names


def test_library_exports_every_declared_symbol(built_lib):
    handle = C.CDLL(built_lib)
    for name in _declared():
        assert hasattr(handle, name), name


def test_python_binding_covers_the_header(built_lib):
    import videoprism_b200._lib as L
    assert sorted(L.EXPORTED_SYMBOLS) == _declared()
    L.lib()  # sets argtypes for every symbol


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import videoprism_b200 as vp
    m = vp.get_model("videoprism_public_v1_base")
    with pytest.raises(Exception) as ei:
        m.param_shapes()
    assert "CUDA" in str(ei.value) or "device" in str(ei.value)
    import videoprism_b200._lib as L
    assert L.lib().vp_device_sm_count() < 0
