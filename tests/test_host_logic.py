"""CPU tests of host-side logic that needs no device: the uploaded-state fingerprint, refused configurations."""
import numpy as np
import pytest

import videoprism_b200 as vp


def test_state_fingerprint_sees_replaced_leaves_in_a_mutated_dict():
    """`_bind` used to key the uploaded state on id(variables) alone: a dict mutated in place was silently not uploaded."""
    fp = vp.models._Module._fingerprint
    state = {"params/a/w": np.zeros((2, 3), np.float32), "params/a/b": np.zeros((3,), np.float32)}
    f0 = fp(state)
    assert fp(state) == f0                                           # stable while nothing changes
    state["params/a/b"] = np.ones((3,), np.float32)                  # a leaf replaced in the SAME dict object
    assert fp(state) != f0
    nested = {"params": {"a": {"w": np.zeros((2, 3), np.float32)}}}
    f1 = fp(nested)
    nested["params"]["a"]["w"] = np.zeros((2, 3), np.float32)        # same values, new array: re-upload (cheap, and safe)
    assert fp(nested) != f1


def test_non_causal_text_tower_is_refused_loudly():
    cfg = dict(vp.CONFIGS["videoprism_lvt_v1_base"])
    cfg["vocabulary_size"] = 32000
    cfg["enable_causal_atten"] = False
    m = vp.FactorizedVideoCLIP(**cfg)
    with pytest.raises(NotImplementedError):
        m._ensure_handle()


def test_check_mode_switches():
    m = vp.get_model("videoprism_public_v1_base")
    assert m.check_fp32 is False or m.check_fp32 is True             # follows VP_CHECK_FP32
    assert vp.get_model("videoprism_public_v1_base", check_fp32=True).check_fp32 is True
    assert vp.get_model("videoprism_public_v1_base", check_fp32=False).check_fp32 is False


def test_async_api_refuses_device_tensors_and_bad_buffers():
    import torch
    m = vp.get_model("videoprism_public_v1_base")
    with pytest.raises(ValueError):
        m.forward_async(torch.zeros((1, 16, 288, 288, 3)))           # host numpy buffers only
