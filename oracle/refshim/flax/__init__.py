"""Stand-in for flax (apply-only linen subset); see ../README.md."""
from . import linen  # noqa: F401
