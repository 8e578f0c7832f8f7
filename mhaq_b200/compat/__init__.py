"""Import-path compatibility with the reference checkout.

The reference's scripts, callbacks and Trainer import the quantization package as ``src.*``
(``from src.quantization.quantizer import Quantizer``,
``from src.quantization.gdnsq.layers.gdnsq_act import NoisyAct``, ``src.aux.types.QScheme`` ...).
``install_src_alias()`` makes those exact import statements resolve to this package:

    import mhaq_b200.compat as compat
    compat.install_src_alias()
    from src.quantization.quantizer import Quantizer          # -> mhaq_b200.quantization.quantizer
    qmodel = Quantizer(config)().quantize(lmodel, in_place=True)

Every ``mhaq_b200.quantization[.*]`` and ``mhaq_b200.aux[.*]`` module is registered under the
corresponding ``src.`` name (the SAME module objects, so ``isinstance`` checks agree whichever
name a caller used).  It refuses to shadow a real ``src`` package that is already imported —
inside the reference checkout, edit the imports instead.
"""
from __future__ import annotations

import importlib
import pkgutil
import sys
import types

_ROOTS = ("quantization", "aux")


def _walk(root: str):
    pkg = importlib.import_module(f"mhaq_b200.{root}")
    yield root, pkg
    for info in pkgutil.walk_packages(pkg.__path__, prefix=f"mhaq_b200.{root}."):
        leaf = info.name.rsplit(".", 1)[-1]
        if leaf.startswith("_"):
            continue
        yield info.name[len("mhaq_b200."):], importlib.import_module(info.name)


def install_src_alias() -> list:
    """Register the aliases; returns the sorted list of ``src.*`` names now importable."""
    cur = sys.modules.get("src")
    if cur is not None and not getattr(cur, "__mhaq_alias__", False):
        raise RuntimeError("a real `src` package is already imported in this process "
                           f"({getattr(cur, '__path__', cur)}); refusing to shadow it")
    if cur is None:
        cur = types.ModuleType("src")
        cur.__path__ = []
        cur.__mhaq_alias__ = True
        cur.__doc__ = "alias of mhaq_b200 (mhaq_b200.compat.install_src_alias)"
        sys.modules["src"] = cur
    names = []
    for root in _ROOTS:
        for rel, mod in _walk(root):
            alias = f"src.{rel}"
            sys.modules[alias] = mod
            parent_name, _, leaf = alias.rpartition(".")
            setattr(sys.modules[parent_name], leaf, mod)
            names.append(alias)
    return sorted(names)


def remove_src_alias() -> None:
    cur = sys.modules.get("src")
    if cur is None or not getattr(cur, "__mhaq_alias__", False):
        return
    for name in [n for n in sys.modules if n == "src" or n.startswith("src.")]:
        del sys.modules[name]
