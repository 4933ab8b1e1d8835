"""numpy-backed subset of jax.numpy used by the reference (float32 semantics, immutable-style arrays)."""
import numpy as _np

newaxis = None
nan = _np.nan
float32 = _np.float32
float64 = _np.float64
int32 = _np.int32
bfloat16 = _np.float32  # not exercised (fp32 goldens only)
inexact = _np.inexact
integer = _np.integer
floating = _np.floating
dtype = _np.dtype
finfo = _np.finfo
iinfo = _np.iinfo
issubdtype = _np.issubdtype
ndarray = _np.ndarray


class Arr(_np.ndarray):
    """ndarray whose augmented assignments REBIND (jax arrays are immutable; `x *= y` in the reference must
    not write through views such as `paddings[:, None, None, :]`)."""
    __iadd__ = lambda self, o: _np.add(self, o)
    __isub__ = lambda self, o: _np.subtract(self, o)
    __imul__ = lambda self, o: _np.multiply(self, o)
    __itruediv__ = lambda self, o: _np.true_divide(self, o)

    def astype(self, dt, *a, **k):
        return _np.ndarray.astype(self, dt, *a, **k).view(Arr)


def _w(x):
    return x.view(Arr) if isinstance(x, _np.ndarray) else x


def _axis(a):
    return tuple(a) if isinstance(a, list) else a


def asarray(x, dtype=None):
    return _w(_np.asarray(x, dtype=dtype))


array = asarray


def arange(*a, dtype=None):
    return _w(_np.arange(*a, dtype=dtype))


def zeros(shape, dtype=float32):
    return _w(_np.zeros(shape, dtype=dtype))


def mean(x, axis=None, keepdims=False):
    return _w(_np.mean(x, axis=_axis(axis), keepdims=keepdims, dtype=x.dtype))


def sum(x, axis=None, keepdims=False):  # noqa: A001
    return _w(_np.sum(x, axis=_axis(axis), keepdims=keepdims, dtype=x.dtype))


def _lift(fn):
    def g(*a, **k):
        return _w(fn(*a, **k))
    g.__name__ = fn.__name__
    return g


square = _lift(_np.square)
sqrt = _lift(_np.sqrt)
exp = _lift(_np.exp)
sin = _lift(_np.sin)
cos = _lift(_np.cos)
tanh = _lift(_np.tanh)
where = _lift(_np.where)
tile = _lift(_np.tile)
minimum = _lift(_np.minimum)
maximum = _lift(_np.maximum)
multiply = _lift(_np.multiply)
transpose = _lift(_np.transpose)
reshape = _lift(_np.reshape)
squeeze = _lift(_np.squeeze)
expand_dims = _lift(_np.expand_dims)
concatenate = _lift(_np.concatenate)
repeat = _lift(_np.repeat)
pad = _lift(_np.pad)


def einsum(eqn, *ops):
    # float32 in, float32 out; accumulate in float64 then round once (XLA's CPU dot accumulates in f32 with a
    # blocked order: both are within 1e-6 relative of the exact sum)
    dt = _np.result_type(*[o.dtype for o in ops])
    return _w(_np.einsum(eqn, *[_np.asarray(o, dtype=_np.float64) for o in ops], optimize=True).astype(dt))
