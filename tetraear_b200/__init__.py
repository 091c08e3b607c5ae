"""Importable alias of the product package.

The package lives in ``tetraear-bladerf_b200/`` (the name the build contract fixes); a
hyphen is not a legal Python identifier, so this stub points ``tetraear_b200.__path__`` at
that directory and executes its ``__init__.py``.  ``import tetraear_b200.signal`` therefore
loads ``tetraear-bladerf_b200/signal/__init__.py``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "tetraear-bladerf_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
