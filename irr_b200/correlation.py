"""``Correlation`` module and ``CorrelationFunction`` with the reference's constructor and forward signature
(models/correlation_package/correlation.py:1-61).  The reference class is orphaned (no model imports it) and
un-buildable on a modern stack (F3); this one runs.  The PWC parameters (pad 4, k 1, md 4, stride 1/1) take the tuned
tiled kernel and are differentiable (``irr_correlation_bwd`` replaces ``correlation_cuda.backward``); anything else
takes the generic forward kernel (inference only)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class CorrelationFunction(torch.autograd.Function):
    """correlation.py:6-45 as a modern static autograd Function (forward saves the inputs, backward returns
    (grad_input1, grad_input2)); PWC parameters only."""

    @staticmethod
    def forward(ctx, input1, input2):
        input1, input2 = input1.contiguous(), input2.contiguous()
        ctx.save_for_backward(input1, input2)
        return ops.correlation(input1, input2, max_disp=4)

    @staticmethod
    def backward(ctx, grad_output):
        input1, input2 = ctx.saved_tensors
        g1, g2 = ops.correlation_backward(input1, input2, grad_output.contiguous(), ctx.needs_input_grad[0],
                                          ctx.needs_input_grad[1])
        return g1, g2


class Correlation(nn.Module):
    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super().__init__()
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply  # accepted and ignored, like correlation_cuda.cc:14

    def forward(self, input1, input2):
        if (self.kernel_size == 1 and self.stride1 == 1 and self.stride2 == 1 and self.max_displacement == 4
                and self.pad_size == 4):
            if torch.is_grad_enabled() and (input1.requires_grad or input2.requires_grad):
                return CorrelationFunction.apply(input1, input2)
            return ops.correlation(input1.contiguous(), input2.contiguous(), max_disp=4)
        if torch.is_grad_enabled() and (input1.requires_grad or input2.requires_grad):
            raise NotImplementedError("irr_b200.Correlation: backward is implemented for the PWC parameters "
                                      "(pad 4, kernel 1, max_displacement 4, strides 1/1) only")
        return ops.correlation_generic(input1, input2, self.pad_size, self.kernel_size, self.max_displacement,
                                       self.stride1, self.stride2)
