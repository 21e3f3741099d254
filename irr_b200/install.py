"""Swap the irr_b200 kernels in UNDER the reference's own model classes.

``install(models_pkg)`` rebinds the module-level names the reference's forward() bodies call
(``compute_cost_volume``, ``upsample2d_as``, ``WarpingLayer`` …) in ``models.pwc_modules`` and in every PWC model
module that imported them by name, so an unmodified ``models.PWCNet_irr(args)`` etc. runs on the new kernels.
Note models/__init__.py:34-35 rebinds ``models.IRR_PWC`` from the submodule to the class, so submodules are reached
through sys.modules (SURVEY.md §1 gotcha)."""
from __future__ import annotations

import sys

from . import pwc_modules as P

_FUNCS = ["compute_cost_volume", "upsample2d_as"]


def install(models_pkg=None, prefix: str = "models"):
    patched = []
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name == prefix or name.startswith(prefix + ".")):
            continue
        for fn in _FUNCS:
            if hasattr(mod, fn):
                setattr(mod, fn, getattr(P, fn))
                patched.append(f"{name}.{fn}")
        if hasattr(mod, "WarpingLayer"):
            setattr(mod, "WarpingLayer", P.WarpingLayer)
            patched.append(f"{name}.WarpingLayer")
    return patched
