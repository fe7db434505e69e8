"""Drop-in shim for Ulixes-8/UNet-Implementations: put this directory FIRST on PYTHONPATH (followed by this repository's
root and the reference's `Our_UNet/`), and the reference's unmodified `src/main.py`, `src/train.py`, `src/evaluate.py`
and `utils/visualize.py` pick up the B200 implementation:

    PYTHONPATH=<repo>/dropin:<repo>:<reference>/Our_UNet  python <reference>/Our_UNet/src/main.py ...

Python imports `sitecustomize` at start-up; it installs a meta-path finder that resolves
    models.unet, models.losses                 (train.py:28-29)            -> unet_implementations_b200.models.*
    src.models.unet, src.models.losses         (evaluate.py:32)            -> the same
    src.utils, src.utils.metrics, src.utils.visualize (evaluate.py:33-41)  -> the reference's own utils.* modules
(the `src.models` / `src.utils` spelling is a layout the reference does not ship).  No file of the
reference is edited."""
import importlib
import importlib.abc
import importlib.util
import sys
import types

_ALIASES = {
    "models.unet": "unet_implementations_b200.models.unet",
    "models.losses": "unet_implementations_b200.models.losses",
    "src.models.unet": "unet_implementations_b200.models.unet",
    "src.models.losses": "unet_implementations_b200.models.losses",
    "src.utils": "utils",
    "src.utils.metrics": "utils.metrics",
    "src.utils.visualize": "utils.visualize",
}
_PACKAGES = {"src.models"}  # empty namespace-like packages that only exist to carry the aliases


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        if self.target is None:
            m = types.ModuleType(spec.name)
            m.__path__ = []
            return m
        return importlib.import_module(self.target)

    def exec_module(self, module):
        return None


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path=None, target=None):
        if name in _ALIASES:
            return importlib.util.spec_from_loader(name, _AliasLoader(_ALIASES[name]), is_package=name == "src.utils")
        if name in _PACKAGES:
            return importlib.util.spec_from_loader(name, _AliasLoader(None), is_package=True)
        return None


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
