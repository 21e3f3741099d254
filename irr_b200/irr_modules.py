"""Host-side mirror of the reference's ``models/irr_modules.py`` (inference path): OccUpsampleNetwork, RefineFlow,
RefineOcc with identical parameter names and forward() signatures, executed by the irr_b200 kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .pwc_modules import conv


def upsample_factor2(inputs, target_as):
    """models/irr_modules.py:21-27."""
    _, _, h, w = target_as.size()
    return ops.upsample_nearest2x(inputs, h, w)


class OccUpsampleNetwork(nn.Module):
    """models/irr_modules.py:30-56.  The residual adds / *0.1 / skip connections are conv epilogues
    (y = addend + alpha*act(conv+b)); the input concat is a pre-allocated 11-channel buffer."""

    def __init__(self, ch_in, ch_out):
        super().__init__()
        self.feat_dim = 32
        self.ch_in = ch_in
        self.init_conv = conv(ch_in, self.feat_dim)
        self.res_convs = nn.Sequential(conv(self.feat_dim, self.feat_dim), conv(self.feat_dim, self.feat_dim, isReLU=False))
        self.res_end_conv = conv(self.feat_dim, self.feat_dim)
        self.mul_const = 0.1
        self.out_convs = conv(self.feat_dim, ch_out)

    def forward_into(self, x_in):
        """x_in: (B, ch_in, H, W) buffer whose channel 0 already holds the x2-upsampled occlusion map."""
        x_init = self.init_conv(x_in)
        r = x_init
        for _ in range(3):  # shared weights, irr_modules.py:51-53
            t = self.res_convs[0](r)
            r = self.res_convs[1](t, addend=r, alpha=self.mul_const)
        x_init2 = self.res_end_conv(r, addend=x_init)
        return self.out_convs(x_init2, addend=x_in[:, 0:1])

    def forward(self, occ, x):
        occ, x = ops.pitched(occ), ops.pitched(x)
        B, C, H, W = x.shape
        x_in = ops.empty(B, C + 1, H, W, x.device)
        ops.upsample_nearest2x(occ, H, W, out=x_in[:, 0:1])
        ops.scale_channels(x, out=x_in[:, 1:])
        return self.forward_into(x_in)


def subtract_mean(input):
    """models/irr_modules.py:59-60."""
    return ops.sub_spatial_mean(input)


class _Refine(nn.Module):
    def __init__(self, ch_in):
        super().__init__()
        self.kernel_size = 3
        self.pad_size = 1
        self.ch_in = ch_in
        self.convs = nn.Sequential(
            conv(ch_in, 128, 3, 1, 1), conv(128, 128, 3, 1, 1), conv(128, 64, 3, 1, 1), conv(64, 64, 3, 1, 1),
            conv(64, 32, 3, 1, 1), conv(32, 32, 3, 1, 1), conv(32, self.kernel_size * self.kernel_size, 3, 1, 1))

    def gather(self, x_in, src):
        """convs -> softmax(-f^2) -> weighted 3x3 replicate-padded gather of ``src``."""
        f = x_in
        for c in self.convs:
            f = c(f)
        return ops.refine_gather(f, src)


class RefineFlow(_Refine):
    """models/irr_modules.py:63-104."""

    def forward(self, flow, diff_img, feature):
        flow, diff_img, feature = ops.pitched(flow), ops.pitched(diff_img), ops.pitched(feature)
        B, _, H, W = flow.shape
        x_in = ops.empty(B, self.ch_in, H, W, flow.device)
        ops.sub_spatial_mean(flow, out=x_in[:, 0:2])
        ops.channel_l2norm(diff_img, out=x_in[:, 2:3])
        ops.scale_channels(feature, out=x_in[:, 3:])
        return self.gather(x_in, flow)


class RefineOcc(_Refine):
    """models/irr_modules.py:107-138."""

    def forward(self, occ, feat1, feat2):
        occ, feat1, feat2 = ops.pitched(occ), ops.pitched(feat1), ops.pitched(feat2)
        B, _, H, W = occ.shape
        c1 = feat1.shape[1]
        x_in = ops.empty(B, self.ch_in, H, W, occ.device)
        ops.scale_channels(occ, out=x_in[:, 0:1])
        ops.scale_channels(feat1, out=x_in[:, 1:1 + c1])
        ops.scale_channels(feat2, out=x_in[:, 1 + c1:])
        return self.gather(x_in, occ)
