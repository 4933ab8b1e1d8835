"""Apply-only subset of flax.linen on numpy arrays: Module (auto-dataclass, name/parent scoping), compact,
nowrap, param, Dense, Dropout, scan, remat, initializers, relu.  Parameters are looked up in the nested
`{'params': ...}` tree passed to `Module.apply`; nothing is ever initialised."""
import dataclasses as _dc
import functools as _ft
import types as _types

import numpy as _np

from jax.numpy import Arr, _w

_stack = []          # modules whose compact method is executing (innermost last)
_MISSING = object()


def compact(fn):
    @_ft.wraps(fn)
    def wrapped(self, *a, **k):
        _stack.append(self)
        try:
            return fn(self, *a, **k)
        finally:
            _stack.pop()
    return wrapped


def nowrap(fn):
    return fn


class Module:
    name = None
    parent = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        # like flax: user fields stay positional-or-keyword, `parent` / `name` are keyword-only and last
        ann = dict(cls.__dict__.get("__annotations__", {}))
        ann.pop("name", None); ann.pop("parent", None)
        ann["parent"] = object
        ann["name"] = object
        cls.__annotations__ = ann
        cls.parent = _dc.field(default=None, kw_only=True)
        cls.name = _dc.field(default=None, kw_only=True)
        _dc.dataclass(cls, eq=False, repr=False)

    def __post_init__(self):
        object.__setattr__(self, "_bound", None)
        if self.parent is None and _stack:
            object.__setattr__(self, "parent", _stack[-1])

    # ---- scopes
    def _scope(self):
        if self._bound is not None:
            return self._bound
        if self.parent is None:
            raise RuntimeError(f"module {type(self).__name__} is not bound to parameters")
        if self.name is None:
            return {}
        ps = self.parent._scope()
        return ps.get(self.name, {})

    def param(self, name, init_fn, shape=None, dtype=None, *a, **k):
        sc = self._scope()
        if name not in sc:
            raise KeyError(f"parameter '{name}' missing under module '{self.name}' ({type(self).__name__})")
        v = _np.asarray(sc[name])
        if shape is not None and tuple(v.shape) != tuple(int(s) for s in shape):
            raise ValueError(f"parameter '{name}' of '{self.name}': shape {v.shape} != {tuple(shape)}")
        return _w(v)

    def apply(self, variables, *args, method=None, rngs=None, mutable=False, **kwargs):
        object.__setattr__(self, "_bound", variables["params"])
        try:
            fn = self.__call__ if method is None else _ft.partial(method, self)
            return fn(*args, **kwargs)
        finally:
            object.__setattr__(self, "_bound", None)


class Dense(Module):
    features: int = 0
    use_bias: bool = True
    kernel_init: object = None
    bias_init: object = None
    param_dtype: object = _np.float32
    promote_dtype: object = None
    dtype: object = None

    def __call__(self, inputs):
        kernel = self.param("kernel", self.kernel_init, (inputs.shape[-1], self.features), self.param_dtype)
        bias = self.param("bias", self.bias_init, (self.features,), self.param_dtype) if self.use_bias else None
        if self.promote_dtype is not None:
            inputs, kernel, bias = self.promote_dtype(inputs, kernel, bias, dtype=self.dtype)
        y = _w((_np.asarray(inputs, _np.float64) @ _np.asarray(kernel, _np.float64)).astype(inputs.dtype))
        if bias is not None:
            y = y + _np.reshape(bias, (1,) * (y.ndim - 1) + (-1,))
        return y


class Dropout(Module):
    rate: float = 0.0

    def __init__(self, rate=0.0, name=None, parent=None, **kw):
        object.__setattr__(self, "rate", rate)
        object.__setattr__(self, "name", name)
        object.__setattr__(self, "parent", parent)
        self.__post_init__()

    def __call__(self, inputs, deterministic=True):
        assert deterministic or not self.rate, "the shim only implements inference (dropout disabled)"
        return inputs


def remat(fn, prevent_cse=True, policy=None, **kw):
    return fn


def scan(body_fn, variable_axes=None, split_rngs=None, length=0, **kw):
    """nn.scan over a module whose parameters carry a leading [length] axis (variable_axes={'params': 0})."""
    def run(module, carry, *xs):
        parent_scope = module.parent._scope()
        stacked = parent_scope[module.name]
        from jax import tree_util
        ys = []
        for i in range(length):
            sliced = tree_util.tree_map(lambda v: _np.asarray(v)[i], stacked)
            object.__setattr__(module, "_bound", sliced)
            try:
                carry, y = body_fn(module, carry, *xs)
            finally:
                object.__setattr__(module, "_bound", None)
            ys.append(y)
        return carry, ys
    return run


def relu(x):
    return _w(_np.maximum(x, 0))


def _init(*a, **k):
    return lambda *aa, **kk: None


initializers = _types.SimpleNamespace(lecun_normal=_init, zeros_init=_init, constant=_init, normal=_init, Initializer=object)
module = _types.SimpleNamespace(VariableDict=dict)
