"""Importable alias of the product package, which lives in the directory `videoprism-mlx_b200/`
(a hyphen is not importable with the `import` statement).  `import videoprism_b200` executes that
package's `__init__` with this module as the package object, so `videoprism_b200.models` etc. are
the files under `videoprism-mlx_b200/`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "videoprism-mlx_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
