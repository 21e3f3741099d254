"""``Correlation`` module with the reference's constructor and forward signature
(models/correlation_package/correlation.py:47-61).  The reference class is orphaned (no model imports it)
and un-buildable on a modern stack (F3); this one runs.  The PWC parameters (pad 4, k 1, md 4, stride 1/1) take the
tuned tiled kernel; anything else takes the generic kernel.  Inference only (no autograd)."""
from __future__ import annotations

import torch.nn as nn

from . import ops


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
            return ops.correlation(input1.contiguous(), input2.contiguous(), max_disp=4)
        return ops.correlation_generic(input1, input2, self.pad_size, self.kernel_size, self.max_displacement,
                                       self.stride1, self.stride2)
