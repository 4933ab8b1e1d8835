"""Host-side text path: tokenizer / tokenize_texts / canonicalize_text / checkpoint trees, held to the reference's own
known-answer tests (tokenizers_test.py:57-73, models_test.py:93-116, utils_test.py:23-26) on the reference's own
SentencePiece fixture.  The fixture is read where it lies under /root/reference (it is a binary asset of the reference,
not copied into this repo), so these tests skip on a box without it."""
import os

import numpy as np
import pytest

import videoprism_b200 as vp
from videoprism_b200 import models, tokenizers, utils

SPM = "/root/reference/videoprism/assets/testdata/test_spm.model"
needs_spm = pytest.mark.skipif(not os.path.exists(SPM), reason="reference SentencePiece fixture not on this box")


def test_canonicalize_text_known_answers():
    assert utils.canonicalize_text("Hello, World!") == "hello world."
    assert utils.canonicalize_text("Hello,World..") == "hello world."
    assert utils.canonicalize_text("  Hello   WORLD") == "hello world."
    assert utils.canonicalize_text("") == "."


@needs_spm
def test_sentencepiece_tokenizer_known_answers():
    tok = tokenizers.SentencePieceTokenizer(SPM)
    assert tok.vocab_size == 1000
    bos, eos = tok.bos_token, tok.eos_token
    assert (bos, eos) == (1, 2)
    assert tok.to_int("blah") == [80, 180, 60]
    assert tok.to_int("blah", bos=True) == [bos, 80, 180, 60]
    assert tok.to_int("blah", eos=True) == [80, 180, 60, eos]
    assert tok.to_int("blah", bos=True, eos=True) == [bos, 80, 180, 60, eos]
    assert tok.to_int(["blah", "blah blah"]) == [[80, 180, 60], [80, 180, 60, 80, 180, 60]]


@needs_spm
def test_tokenize_texts_known_answers():
    tok = tokenizers.SentencePieceTokenizer(SPM)
    ids, paddings = models.tokenize_texts(tok, ["blah", "blah blah", "blah blah blah"], max_length=6, add_bos=False,
                                          canonicalize=False)
    np.testing.assert_array_equal(ids, [[80, 180, 60, 0, 0, 0], [80, 180, 60, 80, 180, 60], [80, 180, 60, 80, 180, 60]])
    np.testing.assert_array_equal(paddings, [[0, 0, 0, 1, 1, 1], [0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0]])
    assert ids.dtype == np.int32 and paddings.dtype == np.float32
    # default arguments: canonicalised, bos added because the fixture has one, TEXT_MAX_LEN columns
    ids, paddings = models.tokenize_texts(tok, ["Blah!"])
    assert ids.shape == (1, models.TEXT_MAX_LEN) and ids[0, 0] == tok.bos_token
    assert paddings[0].sum() == models.TEXT_MAX_LEN - (ids[0] != 0).sum()
    ids0, pad0 = models.tokenize_texts(tok, [])
    assert ids0.shape == (0, models.TEXT_MAX_LEN) and pad0.shape == (0, models.TEXT_MAX_LEN)


def test_text_tokenizer_registry_errors():
    with pytest.raises(ValueError):
        models.load_text_tokenizer("no_such_tokenizer")
    with pytest.raises(FileNotFoundError):       # offline and not a local file: a clear error, not a hang
        os.environ.pop("VIDEOPRISM_SPM_MODEL", None)
        os.environ["HF_HUB_OFFLINE"] = "1"
        tokenizers.SentencePieceTokenizer("definitely_missing.model")


def test_checkpoint_tree_round_trip(tmp_path):
    flat = {"params/a/b/w": np.arange(6, dtype=np.float32).reshape(2, 3), "params/a/c": np.ones(2, np.float32),
            "params/z": np.zeros(1, np.float32)}
    path = str(tmp_path / "ckpt.npz")
    np.savez(path, **flat)
    tree = utils.load_checkpoint(path)
    assert set(tree) == {"params"} and set(tree["params"]) == {"a", "z"} and set(tree["params"]["a"]) == {"b", "c"}
    names = [n for n, _ in utils.tree_flatten_with_names(tree)]
    assert names == sorted(flat)
    again = utils.recover_tree(*zip(*utils.traverse_with_names(tree)))
    np.testing.assert_array_equal(again["params"]["a"]["b"]["w"], flat["params/a/b/w"])
    inner = [n for n, _ in utils.traverse_with_names(tree, with_inner_nodes=True)]
    assert inner[-1] == "" and "params/a" in inner and inner.index("params/a/b/w") < inner.index("params/a/b")
    assert models.load_checkpoint(path)["params"]["z"].shape == (1,)
    with pytest.raises(FileNotFoundError):
        utils.load_checkpoint(str(tmp_path / "missing.npz"))


def test_classifier_registry_and_config_surface():
    assert models.K400_NUM_CLASSES == 400 and models.SSV2_NUM_CLASSES == 174
    m = models.videoprism_vc_v1_base(num_classes=models.K400_NUM_CLASSES)
    assert isinstance(m, vp.FactorizedVideoClassifier) and m.num_classes == 400 and m.config["model_dim"] == 768
    assert models.videoprism_vc_v1_large(num_classes=7).config["num_spatial_layers"] == 24
    with pytest.raises(ValueError):
        vp.FactorizedVideoClassifier(encoder_params=models.CONFIGS["videoprism_v1_base"], num_classes=0)
    giant = models.videoprism_v1_giant()
    assert giant.config["model_dim"] == 1408 and giant.config["model_dim"] // giant.config["num_heads"] == 88
    assert models.videoprism_vc_v1_giant(num_classes=3).config["num_spatial_layers"] == 40
    assert models.videoprism_lvt_v1_giant().config["norm_policy"] == "primer_hybrid"
    cfg = models.get_model_config("videoprism_lvt_public_v1_base")
    assert cfg["num_auxiliary_layers"] == 2 and cfg["vocabulary_size"] == 32000
    with pytest.raises(ValueError):
        models.get_model_config("nope")
