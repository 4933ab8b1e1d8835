"""Stand-in for einshape.jax_einshape on numpy arrays (single-letter axes, e.g. '(bt)nd->(bn)td')."""
import einops as _einops
import numpy as _np


def _spaced(side: str) -> str:
    return " ".join(ch if ch in "()" else ch for ch in side).replace("( ", "(").replace(" )", ")")


def jax_einshape(equation: str, value, **sizes):
    lhs, rhs = equation.split("->")
    out = _einops.rearrange(_np.asarray(value), f"{_spaced(lhs)} -> {_spaced(rhs)}", **sizes)
    from jax.numpy import Arr
    return _np.ascontiguousarray(out).view(Arr)
