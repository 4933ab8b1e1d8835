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


def test_header_is_valid_c_and_a_c_caller_links_and_runs(built_lib, tmp_path):
    """examples/abi_smoke.c: the header compiles as C99 and a plain C program links against the library and gets the
    documented behaviour (no CUDA device here: vp_create_ex fails loudly with VP_ERR_CUDA)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(built_lib)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "abi_smoke.c"),
                        "-L" + libdir, "-lvideoprism_b200", "-Wl,-rpath," + libdir, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "failed loudly" in r.stdout or "parameter leaves" in r.stdout
