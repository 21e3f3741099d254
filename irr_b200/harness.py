"""The evaluation harness either side of the hot path: the reference's ``ModelAndLoss``
(configuration.py:23-62) and ``EvaluationEpoch`` (runtime.py:354-469) contracts, without their per-step host syncs.

* ``ModelAndLoss(args, model, training_loss, evaluation_loss)`` — same constructor and ``forward(example_dict) ->
  (loss_dict, output_dict)``; inputs and targets live in ONE dictionary (configuration.py:42-44).
* ``evaluate(model_and_loss, loader)`` — what ``EvaluationEpoch.run`` computes (the batch-size-weighted moving average
  of every loss key, runtime.py:417-431 + tools.MovingAverage), but: host batches are staged through pinned memory and
  copied with ``non_blocking=True`` on a side stream one batch ahead (the reference does a blocking ``.cuda()`` per key,
  runtime.py:365-368), and the running sums stay on the device — one ``.item()`` per key at the end instead of one per
  key per step (runtime.py:427).
"""
from __future__ import annotations

from typing import Dict, Iterable

import torch
import torch.nn as nn


class ModelAndLoss(nn.Module):
    def __init__(self, args, model, training_loss=None, evaluation_loss=None):
        super().__init__()
        self._model = model
        self._training_loss = training_loss
        self._evaluation_loss = evaluation_loss

    @property
    def training_loss(self):
        return self._training_loss

    @property
    def evaluation_loss(self):
        return self._evaluation_loss

    @property
    def model(self):
        return self._model

    def num_parameters(self):
        return sum(p.data.nelement() if p.requires_grad else 0 for p in self.parameters())

    def forward(self, example_dict):
        output_dict = self._model(example_dict)
        if self.training:
            if self._training_loss is None:
                raise RuntimeError("irr_b200.harness.ModelAndLoss: training is out of scope (eval-mode hot path only)")
            loss_dict = self._training_loss(output_dict, example_dict)
        else:
            loss_dict = self._evaluation_loss(output_dict, example_dict)
        return loss_dict, output_dict


def _is_tensor_key(k: str) -> bool:  # runtime.py:359-361
    return "input" in k or "target" in k


def _stage(example: Dict, device, stream) -> Dict:
    out = dict(example)
    with torch.cuda.stream(stream):
        for k, v in example.items():
            if _is_tensor_key(k) and torch.is_tensor(v):
                if v.device.type == "cpu" and not v.is_pinned():
                    v = v.pin_memory()
                out[k] = v.to(device, non_blocking=True)
    return out


def evaluate(model_and_loss: ModelAndLoss, loader: Iterable[Dict], device=None) -> Dict[str, float]:
    """Average of every evaluation-loss key over the loader, weighted by batch size (tools.MovingAverage semantics)."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    model_and_loss.eval()
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)
    sums: Dict[str, torch.Tensor] = {}
    count = 0
    it = iter(loader)
    try:
        nxt = _stage(next(it), device, copy_stream)
    except StopIteration:
        return {}
    with torch.no_grad():
        while nxt is not None:
            main.wait_stream(copy_stream)
            cur = nxt
            for v in cur.values():
                if torch.is_tensor(v) and v.is_cuda:
                    v.record_stream(main)
            try:
                nxt = _stage(next(it), device, copy_stream)  # next batch's H2D overlaps this batch's forward
            except StopIteration:
                nxt = None
            loss_dict, _ = model_and_loss(cur)
            bs = cur["input1"].size(0)  # runtime.py:380
            for k, v in loss_dict.items():
                v = v.detach().double() * bs
                sums[k] = v if k not in sums else sums[k] + v
            count += bs
    return {k: (v / count).item() for k, v in sums.items()}
