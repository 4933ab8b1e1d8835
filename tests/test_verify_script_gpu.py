"""scripts/verify_clip_models.py (the reference's verify_clip_models.py, reshaped for this library): production path vs
fp32 check mode on a clip + a query, and the check mode against the CPU oracle under the reference's own criterion
(max-abs < 1e-3 on the video embedding, the text embedding and the similarity; verify_clip_models.py:92-95)."""
import importlib.util
import os

import numpy as np
import pytest

import videoprism_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _script():
    spec = importlib.util.spec_from_file_location("verify_clip_models", os.path.join(ROOT, "scripts", "verify_clip_models.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _against_oracle(name, r):
    import videoprism_b200 as vp
    state = r["state"]
    flat = dict(vp.utils.tree_flatten_with_names(state)) if "params" in state else state
    vo, to, _ = O.run_clip(O.CONFIGS[name], {k: np.asarray(v, np.float32) for k, v in flat.items()}, r["video"], r["ids"], r["paddings"])
    vc, tc = r["check"]
    dv, dt, ds = np.abs(vc - vo).max(), np.abs(tc - to).max(), np.abs(vc @ tc.T - vo @ to.T).max()
    print(f"[verify] {name}: fp32 check mode vs CPU oracle: video {dv:.3e}, text {dt:.3e}, similarity {ds:.3e} (bound 1e-3)")
    assert dv < 1e-3 and dt < 1e-3 and ds < 1e-3


def test_verify_script_with_synthetic_stand_ins():
    """No released weights offline: seeded random-init weights, the mp4 fixture if present (else synthetic frames)."""
    results = {}
    name = "videoprism_lvt_public_v1_base"
    rc = _script().main(["--models", name, "--text", "child drumming on water bottles", "--text", "a cat sleeping"], results=results)
    assert rc == 0
    assert any("SYNTHETIC" in u for u in results[name]["used"])
    _against_oracle(name, results[name])


@pytest.mark.skipif(not os.environ.get("VP_WEIGHTS_DIR") or not os.environ.get("VP_SPM_MODEL"),
                    reason="released checkpoints (VP_WEIGHTS_DIR) and the c4_en SentencePiece model (VP_SPM_MODEL) are not reachable offline")
def test_verify_script_with_released_weights():
    """The reference's check as it is meant to be run: released weights, the repository's mp4 clip, the real tokenizer."""
    results = {}
    video = os.environ.get("VP_VERIFY_VIDEO", os.path.join("videoprism", "assets", "water_bottle_drumming.mp4"))
    rc = _script().main(["--require-real", "--video", video], results=results)
    assert rc == 0
    for name, r in results.items():
        _against_oracle(name, r)
