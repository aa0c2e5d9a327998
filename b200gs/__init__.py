"""Importable alias of the product package.

The product lives in ``gaussian-splatting-lightning_b200/`` (the directory name the build contract asks for); a hyphen
is not importable, so ``import b200gs`` resolves to that directory: this shim points ``__path__`` at it and executes
its ``__init__``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "gaussian-splatting-lightning_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
