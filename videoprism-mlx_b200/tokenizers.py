"""Text tokenizers with the reference's interface (videoprism/tokenizers.py:29-196).

Host-side: produces the int32 ids / float paddings that `vp_clip_text_forward` consumes.  SentencePiece itself is the
third-party library the reference uses too; nothing here touches the device.
"""
from __future__ import annotations

import os
from typing import List, Protocol, Sequence, Union


class Tokenizer(Protocol):
    """tokenizers.py:29-77."""

    def to_int(self, text: Union[str, Sequence[str]], *, bos: bool = False, eos: bool = False): ...

    @property
    def pad_token(self) -> int: ...

    @property
    def eos_token(self) -> int: ...

    @property
    def bos_token(self) -> int: ...

    @property
    def vocab_size(self) -> int: ...


class SentencePieceTokenizer:
    """tokenizers.py:80-196.  `model_path` is a local SentencePiece model file; anything else (the reference's default
    'c4_en.model', or a legacy gs:// path, :91-93) is fetched from the reference's HuggingFace repo when the hub is
    reachable, and raises FileNotFoundError with the path to provide otherwise (the GPU boxes have no network)."""

    HF_REPO = "tom-moroney/videoprism-mlx"   # tokenizers.py:95-98

    def __init__(self, model_path: str = "c4_en.model"):
        from sentencepiece import SentencePieceProcessor
        if model_path.startswith("gs://"):
            model_path = "c4_en.model"
        local = model_path if os.path.isfile(model_path) else os.environ.get("VIDEOPRISM_SPM_MODEL", "")
        if not os.path.isfile(local):
            try:
                from huggingface_hub import hf_hub_download
                local = hf_hub_download(repo_id=self.HF_REPO, filename=os.path.basename(model_path))
            except Exception as e:  # offline: say what to do instead of failing somewhere inside the hub client
                raise FileNotFoundError(
                    f"SentencePiece model '{model_path}' is not a local file and could not be downloaded from "
                    f"{self.HF_REPO} ({type(e).__name__}); pass a local path or set VIDEOPRISM_SPM_MODEL") from e
        self._model = SentencePieceProcessor()
        self._model.Load(local)

    def to_int(self, text: Union[str, Sequence[str]], *, bos: bool = False, eos: bool = False):
        """tokenizers.py:102-126: a str -> list[int]; a sequence of str -> list[list[int]]."""
        head: List[int] = [self.bos_token] if bos else []
        tail: List[int] = [self.eos_token] if eos else []
        if isinstance(text, str):
            return head + self._model.EncodeAsIds(text) + tail
        return [head + self._model.EncodeAsIds(s) + tail for s in text]

    def to_int_tf_op(self, text, *, bos: bool = False, eos: bool = False):
        """tokenizers.py:128-173 (TensorFlow data pipelines): not part of this path."""
        raise ImportError("TensorFlow is required for to_int_tf_op(); use to_int() (tokenizers.py:146-151)")

    @property
    def pad_token(self) -> int:
        return self._model.pad_id()

    @property
    def eos_token(self) -> int:
        return self._model.eos_id()

    @property
    def bos_token(self) -> int:
        return self._model.bos_id()

    @property
    def vocab_size(self) -> int:
        return self._model.GetPieceSize()
