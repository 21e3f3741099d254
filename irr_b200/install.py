"""Swap the irr_b200 kernels in UNDER the reference's own model classes.

``install(models_pkg)`` rebinds the module-level names the reference's forward() bodies call
(``compute_cost_volume``, ``upsample2d_as``, ``WarpingLayer``) in ``models.pwc_modules`` and in every PWC model module
that imported them by name, so an unmodified ``models.PWCNet_irr(args)`` / ``models.IRR_PWC(args)`` runs its cost
volumes, warps and bilinear resizes on the new kernels (the convolutions stay ``nn.Conv2d`` — to replace those too, use
the ``irr_b200`` model classes, which load the same checkpoints).  Note models/__init__.py:34-35 rebinds
``models.IRR_PWC`` from the submodule to the class, so submodules are reached through sys.modules (SURVEY.md §1 gotcha).

Ordering: ``WarpingLayer`` is a CLASS the reference instantiates in each model's ``__init__``
(models/IRR_PWC.py:24), so rebinding the name only reaches models constructed AFTER ``install()``.  For a model that
already exists call ``install(models_pkg, model=that_model)`` (or ``patch_instances(model)``): every sub-module whose
class is named ``WarpingLayer`` gets the new forward bound on the instance.

Autograd: the installed functions check ``torch.is_grad_enabled()`` and the inputs' ``requires_grad``: when a gradient
is wanted, the cost volume goes through ``CorrelationFunction`` (``irr_correlation_bwd``), the warp through
``WarpFunction`` (``irr_warp_bwd``), and the bilinear resize is handed back to the reference's ORIGINAL implementation
(kept at install time), so a reference model in train mode keeps back-propagating.
``uninstall()`` restores every patched name.
"""
from __future__ import annotations

import sys
import types

import torch

from . import pwc_modules as P
from .correlation import CorrelationFunction

_FUNCS = ["compute_cost_volume", "upsample2d_as"]
_saved = []        # (module, attribute name, original object)
_originals = {}    # attribute name -> first original seen (the reference's own implementation)
_patched_instances = []  # (module instance, had_own_forward, previous forward)


def _wants_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def compute_cost_volume(feat1, feat2, param_dict):
    """models/pwc_modules.py:42-62 on the irr_b200 kernel; differentiable through irr_correlation_bwd."""
    if _wants_grad(feat1, feat2):
        if param_dict["max_disp"] != 4:
            return _originals["compute_cost_volume"](feat1, feat2, param_dict)
        return CorrelationFunction.apply(feat1, feat2)
    return P.compute_cost_volume(feat1.contiguous(), feat2.contiguous(), param_dict)


def upsample2d_as(inputs, target_as, mode="bilinear"):
    """models/pwc_modules.py:65-67; falls back to the reference's own code when a gradient is required."""
    if _wants_grad(inputs) or mode != "bilinear":
        return _originals["upsample2d_as"](inputs, target_as, mode)
    return P.upsample2d_as(inputs.contiguous(), target_as, mode)


def _warp_forward(self, x, flow, height_im, width_im, div_flow):
    if _wants_grad(x, flow):
        return P.WarpFunction.apply(x, flow, height_im, width_im, div_flow)   # irr_warp_fwd / irr_warp_bwd
    return P.ops.warp(x.contiguous(), flow.contiguous(), height_im, width_im, div_flow)


class WarpingLayer(P.WarpingLayer):
    """models/pwc_modules.py:115-133 on the irr_b200 warp kernel (bit-exact hard mask)."""
    forward = _warp_forward


_NEW = {"compute_cost_volume": compute_cost_volume, "upsample2d_as": upsample2d_as, "WarpingLayer": WarpingLayer}


def patch_instances(model):
    """Bind the new warp forward on every already-constructed ``WarpingLayer`` instance inside ``model``."""
    n = 0
    for mod in model.modules():
        if type(mod).__name__ == "WarpingLayer" and not isinstance(mod, P.WarpingLayer):
            _originals.setdefault("WarpingLayer", type(mod))
            _patched_instances.append((mod, "forward" in mod.__dict__, mod.__dict__.get("forward")))
            mod.forward = types.MethodType(_warp_forward, mod)
            n += 1
    return n


def install(models_pkg=None, prefix: str = "models", model=None):
    """Patch the reference package (``models_pkg`` — the imported ``models`` module — or every ``sys.modules`` entry
    under ``prefix``).  Returns the list of patched ``module.attribute`` names.  Idempotent."""
    if models_pkg is not None:
        prefix = models_pkg.__name__
    patched = []
    for name, mod in list(sys.modules.items()):
        if mod is None or not (name == prefix or name.startswith(prefix + ".")):
            continue
        for attr, new in _NEW.items():
            cur = mod.__dict__.get(attr)
            if cur is None or cur is new:
                continue
            _originals.setdefault(attr, cur)
            _saved.append((mod, attr, cur))
            setattr(mod, attr, new)
            patched.append(f"{name}.{attr}")
    if model is not None:
        k = patch_instances(model)
        patched.append(f"<{k} WarpingLayer instance(s) of the given model>")
    return patched


def uninstall():
    """Undo every ``install()`` / ``patch_instances()`` since the last ``uninstall()``."""
    n = len(_saved) + len(_patched_instances)
    while _saved:
        mod, attr, orig = _saved.pop()
        setattr(mod, attr, orig)
    while _patched_instances:
        mod, had, prev = _patched_instances.pop()
        if had:
            mod.forward = prev
        else:
            mod.__dict__.pop("forward", None)
    _originals.clear()
    return n
