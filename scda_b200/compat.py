"""Make the reference's top-level module names resolve to this package.

The reference driver and its own modules import `extensions`, `functions`,
`models` and `utils` as top-level packages (e.g. functions/rpn_proposal.py:2-4,
models/faster_rcnn/vgg_adver_expansion_cluster.py:5-7,
tools/faster_rcnn_train_val.py:31-38).  `install()` puts a finder on
sys.meta_path that answers `import extensions.<x>` with the module object of
`scda_b200.extensions.<x>` (same object, not a second copy), so code written
against the reference imports the B200 implementations unchanged.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import sys

_TOP = ("extensions", "functions", "models", "utils")


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self):
        self.names = set()

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".", 1)[0] not in self.names:
            return None
        try:
            real = importlib.import_module("scda_b200." + fullname)
        except ModuleNotFoundError as e:
            if e.name == "scda_b200." + fullname:
                return None
            raise
        spec = importlib.machinery.ModuleSpec(fullname, self, is_package=hasattr(real, "__path__"))
        spec._scda_real = real
        return spec

    def create_module(self, spec):
        real = spec._scda_real
        spec._scda_real_spec = real.__spec__
        return real

    def exec_module(self, module):
        # the import machinery re-pointed __spec__ at the alias; put the real one back so
        # relative imports inside the module keep resolving against scda_b200.*
        alias = module.__spec__
        real_spec = getattr(alias, "_scda_real_spec", None)
        if real_spec is not None:
            module.__spec__ = real_spec


_FINDER = _AliasFinder()


def install(names=_TOP, override: bool = False) -> list[str]:
    done = []
    for name in names:
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__name__", "").startswith("scda_b200."):
            if not override:
                raise ImportError(
                    "a different top-level module %r is already imported; pass override=True "
                    "to replace it" % name)
            for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                del sys.modules[k]
        _FINDER.names.add(name)
        done.append(name)
    if _FINDER not in sys.meta_path:
        sys.meta_path.insert(0, _FINDER)
    for name in done:
        importlib.import_module(name)
    return done


def uninstall() -> None:
    if _FINDER in sys.meta_path:
        sys.meta_path.remove(_FINDER)
    for name in list(_FINDER.names):
        for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
            if getattr(sys.modules[k], "__name__", "").startswith("scda_b200."):
                del sys.modules[k]
    _FINDER.names.clear()
