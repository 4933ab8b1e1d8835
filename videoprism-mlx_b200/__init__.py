"""videoprism-mlx_b200: B200-native forward path of VideoPrism's FactorizedEncoder / VideoCLIP.

Python surface mirroring the reference (`videoprism/models.py`, `videoprism/models_mlx.py`) on top of
a C-ABI shared library of hand-written sm_100a CUDA kernels (`csrc/`, `include/videoprism_b200.h`).
There is no CPU fallback: importing is cheap, but creating a model without the built library or
without a B200 raises.
"""
from . import models  # noqa: F401
from . import video_utils  # noqa: F401
from . import tokenizers  # noqa: F401
from . import utils  # noqa: F401
from .models import (  # noqa: F401
    CONFIGS, MODELS, get_model, has_model, load_pretrained_weights, load_video_encoder, load_model,
    FactorizedEncoder, FactorizedVideoCLIP, FactorizedVideoClassifier, load_classifier, get_model_config, synthetic_state, pinned_empty, load_text_tokenizer, tokenize_texts, compute_similarity_matrix,
)

__version__ = "0.1.0"
