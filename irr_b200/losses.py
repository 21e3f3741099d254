"""Eval-mode branches of the reference's losses on the irr_b200 metrics kernel.

Same class names, constructor (``args``) and ``forward(output_dict, target_dict) -> loss_dict`` as
``/root/reference/losses.py``, so the reference's ``ModelAndLoss`` / ``EvaluationEpoch`` (configuration.py:45-62,
runtime.py:354-469) can drive the new model unchanged.  Only the evaluation branch exists (training is out of scope,
DESIGN.md §7): one deterministic launch per batch produces the per-image sums, the few scalar operations that finish
the metric stay on the device (no ``.item()`` here — the harness decides when to synchronise).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def _check_eval(mod):
    if mod.training:
        raise RuntimeError(f"irr_b200.losses.{type(mod).__name__}: only the eval-mode branch is implemented")


class MultiScaleEPE_PWC_Bi_Occ_upsample(nn.Module):
    """losses.py:515-636 (eval branch :634-636): ``epe`` = mean end-point error, ``F1`` = f1_score(target_occ1,
    round(sigmoid(occ))) — per-image F1 (sums over H, W), averaged over the batch (losses.py:24-37)."""

    def __init__(self, args=None):
        super().__init__()
        self._args = args
        self._batch_size = getattr(args, "batch_size", None)
        self._weights = [0.32, 0.08, 0.02, 0.01, 0.005, 0.00125, 0.0003125]
        self.occ_activ = nn.Sigmoid()

    def forward(self, output_dict, target_dict):
        _check_eval(self)
        flow, target = output_dict["flow"], target_dict["target1"]
        B, _, H, W = flow.shape
        s = ops.eval_metrics(flow.float(), target.float(), None, output_dict["occ"].float(),
                             target_dict["target_occ1"].float())
        eps = 1e-8
        tp, npred, ntrue = s[:, 3], s[:, 4], s[:, 5]
        precision = tp / (npred + eps)
        recall = tp / (ntrue + eps)
        f1 = (precision * recall / (precision + recall + eps) * 2.0).mean()
        return {"epe": (s[:, 0].sum() / (B * H * W)).float(), "F1": f1.float()}


class MultiScaleEPE_PWC_Bi_Occ_upsample_Sintel(MultiScaleEPE_PWC_Bi_Occ_upsample):
    """losses.py:579-… — the Sintel fine-tuning loss evaluates exactly like its parent."""


class MultiScaleEPE_PWC_Bi_Occ_upsample_KITTI(nn.Module):
    """losses.py:638-699 (eval branch :688-697): per-image masked ``epe`` and KITTI ``outlier`` rate (> 3 px and > 5 %)."""

    def __init__(self, args=None):
        super().__init__()
        self._args = args
        self._batch_size = getattr(args, "batch_size", None)
        self._weights = [0.001, 0.001, 0.001, 0.002, 0.004, 0.004, 0.004]
        self.occ_activ = nn.Sigmoid()

    def forward(self, output_dict, target_dict):
        _check_eval(self)
        s = ops.eval_metrics(output_dict["flow"].float(), target_dict["target1"].float(),
                             target_dict["input_valid"].float())
        return {"epe": (s[:, 0] / s[:, 1]).mean().float(), "outlier": (s[:, 2] / s[:, 1]).mean().float()}


class MultiScaleEPE_PWC(nn.Module):
    """losses.py (PWCNet family, eval branch): ``epe`` = mean end-point error of ``output_dict['flow']``."""

    def __init__(self, args=None):
        super().__init__()
        self._args = args
        self._batch_size = getattr(args, "batch_size", None)
        self._weights = [0.32, 0.08, 0.02, 0.01, 0.005]

    def forward(self, output_dict, target_dict):
        _check_eval(self)
        flow = output_dict["flow"]
        B, _, H, W = flow.shape
        s = ops.eval_metrics(flow.float(), target_dict["target1"].float())
        return {"epe": (s[:, 0].sum() / (B * H * W)).float()}
