"""Importable alias of the `new-vit_b200/` package directory.

The product package lives in `new-vit_b200/` (the name the build contract fixes); a hyphen is
not a legal Python identifier, so this shim makes `import new_vit_b200` resolve to that
directory by executing its `__init__.py` in this module's namespace.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "new-vit_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
