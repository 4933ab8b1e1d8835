"""CPU checks of the reference-facing Python surface (names, registry, errors, checkpoint tree I/O)."""
import numpy as np
import pytest

import videoprism_b200 as vp
from videoprism_b200 import models


def test_registry_matches_reference():
    # models.py:224-233, models_test.py:28-34
    assert set(models.MODELS) == {"videoprism_public_v1_base", "videoprism_public_v1_large",
                                  "videoprism_lvt_public_v1_base", "videoprism_lvt_public_v1_large"}
    assert vp.has_model("videoprism_public_v1_base")
    assert vp.has_model("google/videoprism-lvt-large-f8r288")
    assert not vp.has_model("videoprism_public_v1_giant")
    assert not vp.has_model("google/unknown")


def test_get_model_configs_and_errors():
    m = vp.get_model("videoprism_public_v1_large")
    assert isinstance(m, vp.FactorizedEncoder)
    assert m.config["model_dim"] == 1024 and m.config["pos_emb_shape"] == (8, 16, 16) and m.config["num_spatial_layers"] == 24
    c = vp.get_model("google/videoprism-lvt-base-f16r288")
    assert isinstance(c, vp.FactorizedVideoCLIP) and c.config["vocabulary_size"] == 32000
    with pytest.raises(ValueError):
        vp.get_model("nope")
    with pytest.raises(ValueError):
        vp.get_model("google/nope")


def test_mlx_style_loader_errors():
    with pytest.raises(ValueError):       # models_mlx.py:169-174
        vp.load_video_encoder("videoprism_lvt_public_v1_base")
    with pytest.raises(ValueError):
        vp.load_model("unknown")
    with pytest.raises(FileNotFoundError):  # models_mlx.py:191-196
        vp.load_video_encoder("videoprism_public_v1_base", weights_path="/nonexistent.npz")


def test_checkpoint_tree_roundtrip(tmp_path):
    flat = {"params/a/b": np.arange(6, dtype=np.float32).reshape(2, 3), "params/a/c": np.ones(4, np.float32), "params/d": np.zeros(1, np.float32)}
    p = tmp_path / "ckpt.npz"
    np.savez(p, **flat)
    tree = models.load_pretrained_weights(None, checkpoint_path=str(p))
    assert set(tree) == {"params"} and set(tree["params"]["a"]) == {"b", "c"}
    back = models._flatten(tree)
    assert set(back) == set(flat) and all(np.array_equal(back[k], flat[k]) for k in flat)
