"""PWC-Net + IRR + occlusion + bi-directional, without the refinement / occlusion-upsampling nets — eval forward.

Drop-in for the reference's ``models/pwcnet_irr_occ_bi.py`` ``PWCNet`` (BASELINE config 4): same constructor, parameter
names (five ``conv_1x1`` blocks, shared ``flow_estimators`` / ``occ_estimators`` / context nets) and
``forward({'input1','input2'}) -> {'flow','occ'}`` (pwcnet_irr_occ_bi.py:43-133).  Both directions run as one 2B batch."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .pwc_modules import (ContextNetwork, FeatureExtractor, FlowEstimatorDense, OccContextNetwork, OccEstimatorDense,
                          WarpingLayer, conv, flow_scales, initialize_msra)


class PWCNet(nn.Module):
    def __init__(self, args=None, div_flow=0.05):
        super().__init__()
        self.args = args
        self._div_flow = div_flow
        self.search_range = 4
        self.num_chs = [3, 16, 32, 64, 96, 128, 196]
        self.output_level = 4
        self.num_levels = 7
        self.leakyRELU = nn.LeakyReLU(0.1, inplace=True)
        self.feature_pyramid_extractor = FeatureExtractor(self.num_chs)
        self.warping_layer = WarpingLayer()
        self.dim_corr = (self.search_range * 2 + 1) ** 2
        self.num_ch_in_flo = self.dim_corr + 32 + 2
        self.num_ch_in_occ = self.dim_corr + 32 + 1
        self.flow_estimators = FlowEstimatorDense(self.num_ch_in_flo)
        self.context_networks = ContextNetwork(self.num_ch_in_flo + 448 + 2)
        self.occ_estimators = OccEstimatorDense(self.num_ch_in_occ)
        self.occ_context_networks = OccContextNetwork(self.num_ch_in_occ + 448 + 1)
        self.conv_1x1 = nn.ModuleList([conv(196, 32, kernel_size=1, stride=1, dilation=1),
                                       conv(128, 32, kernel_size=1, stride=1, dilation=1),
                                       conv(96, 32, kernel_size=1, stride=1, dilation=1),
                                       conv(64, 32, kernel_size=1, stride=1, dilation=1),
                                       conv(32, 32, kernel_size=1, stride=1, dilation=1)])
        self.corr_params = {"pad_size": self.search_range, "kernel_size": 1, "max_disp": self.search_range,
                            "stride1": 1, "stride2": 1, "corr_multiply": 1}
        initialize_msra(self.modules())

    def estimator_level(self, l, feat, flow_up, occ_up, height_im, width_im, record=None):
        """One pass of pwcnet_irr_occ_bi.py:66-121 for pyramid level l on the 2B batch.

        feat   : (2B, C_l, h, w) rows [0,B) = x1 features, rows [B,2B) = x2 features
        flow_up: (2B, 2, h, w) flow in GLOBAL units already resized to this level (zeros at l == 0); rows [0,B) forward
        occ_up : (2B, 1, h, w)
        returns (flow in GLOBAL units, occ) of this level on the 2B batch."""
        feat, flow_up, occ_up = ops.pitched(feat), ops.pitched(flow_up), ops.pitched(occ_up)
        B2, C, h, w = feat.shape
        B = B2 // 2
        dev = feat.device
        df = self._div_flow
        nf, no = self.num_ch_in_flo, self.num_ch_in_occ
        buf_f = ops.empty(B2, 448 + nf + 2, h, w, dev)
        buf_o = ops.empty(B2, 448 + no + 1, h, w, dev)
        corr = buf_f[:, 448:529]
        if l == 0:  # pwcnet_irr_occ_bi.py:68-70,82-85
            ops.correlation(feat, feat, out=corr, shift=B, slope=0.1)
        else:       # :72-85
            ops.warp_correlation(feat, feat, flow_up, height_im, width_im, df, out=corr, shift=B, slope=0.1)
        self.conv_1x1[l](feat, out=buf_f[:, 529:561])  # :91-92
        ops.scale_channels(buf_f[:, 448:561], out=buf_o[:, 448:561])
        su_l, sv_l = flow_scales(h, w, df, width_im, height_im, True)
        su_g, sv_g = flow_scales(h, w, df, width_im, height_im, False)
        ops.scale_channels(flow_up, out=buf_f[:, 561:563], s_even=su_l, s_odd=sv_l)  # :88-89
        ops.scale_channels(occ_up, out=buf_o[:, 561:562])
        # flow (:93-104)
        self.flow_estimators.forward_into(buf_f, out=buf_f[:, 563:565], addend=buf_f[:, 561:563])
        flow = self.context_networks(buf_f, addend=buf_f[:, 563:565])
        ops.scale_channels(flow, out=flow, s_even=su_g, s_odd=sv_g)
        # occlusion (:109-117)
        self.occ_estimators.forward_into(buf_o, out=buf_o[:, 562:563], addend=buf_o[:, 561:562])
        occ = self.occ_context_networks(buf_o, addend=buf_o[:, 562:563])
        if record is not None:
            record.update({"corr": corr.clone(), "x_1by1": buf_f[:, 529:561].clone(), "flow": flow.clone(),
                           "occ": occ.clone(), "flow_up": flow_up.clone(), "occ_up": occ_up.clone()})
        return flow, occ

    def forward(self, input_dict, record=None):
        if self.training:
            raise RuntimeError("irr_b200.pwcnet_irr_occ_bi: only the eval-mode forward is implemented (call .eval())")
        x1_raw, x2_raw = input_dict['input1'], input_dict['input2']
        B, _, height_im, width_im = x1_raw.shape
        B2 = 2 * B
        df = self._div_flow
        # the kernels launch on the CURRENT device / stream: make the input's device current, as the reference's
        # correlation.py:21 does (a model on cuda:1 must work while cuda:0 is current)
        with torch.cuda.device_of(x1_raw), torch.no_grad():
            imgs = ops.stack_pair(x1_raw, x2_raw)  # (2B, 3, H, W), rows pitched when W % 4 != 0
            dev = imgs.device
            pyramid = self.feature_pyramid_extractor(imgs)
            flow = occ = None
            for l, feat in enumerate(pyramid[:self.output_level + 1]):
                _, C, h, w = feat.shape
                if l == 0:
                    flow_up = ops.empty(B2, 2, h, w, dev, zero=True)
                    occ_up = ops.empty(B2, 1, h, w, dev, zero=True)
                else:
                    flow_up = ops.resize_ac(flow, h, w)
                    occ_up = ops.resize_ac(occ, h, w)
                rec_l = None
                if record is not None:
                    rec_l = record[l] = {}
                flow, occ = self.estimator_level(l, feat, flow_up, occ_up, height_im, width_im, rec_l)
            out_flow = ops.resize_ac(flow[:B], height_im, width_im, s_even=1.0 / df, s_odd=1.0 / df, pitched=False)  # :130
            out_occ = ops.resize_ac(occ[:B], height_im, width_im, pitched=False)  # :131
        return {'flow': out_flow, 'occ': out_occ}
