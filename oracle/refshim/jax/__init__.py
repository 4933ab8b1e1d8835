"""numpy-backed subset of the jax API used by the reference (see ../README.md)."""
import math as _math
import types as _types

import numpy as _np
from scipy import special as _special

from . import numpy  # noqa: F401  (jax.numpy)
from .numpy import Arr, _w

Array = _np.ndarray
typing = _types.SimpleNamespace(DTypeLike=object, ArrayLike=object)
checkpoint_policies = _types.SimpleNamespace()


# ---- jax.nn
def _softmax(x, axis=-1):
    m = _np.max(x, axis=axis, keepdims=True)
    e = _np.exp(x - m)
    return _w(e / _np.sum(e, axis=axis, keepdims=True, dtype=x.dtype))


def _gelu(x, approximate=True):
    if approximate:
        return _w(0.5 * x * (1.0 + _np.tanh(_np.float32(_math.sqrt(2 / _math.pi)) * (x + _np.float32(0.044715) * x ** 3))))
    return _w((x * (_special.erf(x / _np.sqrt(_np.asarray(2.0, x.dtype))) + 1) / 2).astype(x.dtype))


def _softplus(x):
    return _w(_np.logaddexp(x, _np.zeros((), x.dtype)).astype(x.dtype))


def _one_hot(ids, num_classes, dtype=_np.float32):
    return _w((_np.asarray(ids)[..., None] == _np.arange(num_classes)).astype(dtype))


nn = _types.SimpleNamespace(softmax=_softmax, gelu=_gelu, softplus=_softplus, one_hot=_one_hot,
                            relu=lambda x: _w(_np.maximum(x, 0)))


# ---- jax.lax
def _slice_in_dim(x, start, limit, stride=1, axis=0):
    idx = [slice(None)] * x.ndim
    idx[axis] = slice(start, limit, stride)
    return _w(x[tuple(idx)])


lax = _types.SimpleNamespace(rsqrt=lambda x: _w((1.0 / _np.sqrt(x)).astype(x.dtype)), slice_in_dim=_slice_in_dim)


# ---- jax.image.resize(method='bilinear'): scale_and_translate with the triangle kernel, antialias=True
def _weights(n_in, n_out, dtype):
    scale = n_out / n_in
    inv_scale = 1.0 / scale
    kernel_scale = max(inv_scale, 1.0)
    sample_f = (_np.arange(n_out, dtype=dtype) + 0.5) * inv_scale - 0.5
    x = _np.abs(sample_f[None, :] - _np.arange(n_in, dtype=dtype)[:, None]) / kernel_scale
    w = _np.maximum(0, 1 - _np.abs(x))
    tot = _np.sum(w, axis=0, keepdims=True)
    w = _np.where(_np.abs(tot) > 1000.0 * float(_np.finfo(_np.float32).eps), _np.divide(w, _np.where(tot != 0, tot, 1)), 0)
    ok = _np.logical_and(sample_f >= -0.5, sample_f <= n_in - 0.5)[None, :]
    return _np.where(ok, w, 0).astype(dtype)


def _resize(image, shape, method="bilinear", antialias=True):
    assert method == "bilinear" and antialias
    out = _np.asarray(image)
    for d, (n_in, n_out) in enumerate(zip(image.shape, shape)):
        if n_in == n_out:
            continue
        w = _weights(n_in, n_out, out.dtype)
        out = _np.moveaxis(_np.tensordot(_np.moveaxis(out, d, -1), w, axes=([-1], [0])), -1, d)
    return _w(out.astype(image.dtype))


image = _types.SimpleNamespace(resize=_resize)


# ---- jax.tree_util
def _tree_map(f, tree, *rest):
    if isinstance(tree, dict):
        return {k: _tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(_tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    if tree is None:
        return None
    return f(tree, *rest)


def _tree_flatten(tree):
    leaves = []
    _tree_map(lambda x: leaves.append(x), tree)
    return leaves, None


tree_util = _types.SimpleNamespace(tree_map=_tree_map, tree_flatten=_tree_flatten)
