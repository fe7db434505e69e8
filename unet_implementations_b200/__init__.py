"""Importable alias of the package directory `unet-implementations_b200/` (a hyphen is not a valid module name).

`import unet_implementations_b200` resolves sub-modules from ../unet-implementations_b200/.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "unet-implementations_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
