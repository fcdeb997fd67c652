"""Importable alias of the `diff-sampler_b200/` package directory (a hyphen is not a valid module name)."""
import os as _os

_real = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '..', 'diff-sampler_b200'))
__path__.append(_real)
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
