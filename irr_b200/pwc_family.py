"""The six remaining classes of the reference's PWC family — eval-mode forwards on the irr_b200
kernels: ``pwcnet_bi`` / ``pwcnet_occ`` / ``pwcnet_occ_bi`` (per-level estimators, context at the output level only) and
``pwcnet_irr`` / ``pwcnet_irr_bi`` / ``pwcnet_irr_occ`` (shared estimators, 1x1 convs, ``rescale_flow``, context at every
level).  The reference writes them as six files that differ in three switches (models/pwcnet_bi.py:41-109,
pwcnet_occ.py:49-117, pwcnet_occ_bi.py:49-132, pwcnet_irr.py:43-97, pwcnet_irr_bi.py:43-111, pwcnet_irr_occ.py:47-112);
here one forward parameterised by (IRR, BI, OCC) serves six thin classes with the reference's constructor, module /
parameter names and ``forward({'input1','input2'}) -> {'flow'[, 'occ']}``.

Data movement as in the benchmarked models: the pyramid runs once on the stacked pair; bi-directional variants process
both directions as ONE 2B batch ("the other image" is a batch rotation inside the warp / correlation kernels);
warp + mask + cost volume + LeakyReLU is one kernel writing into the estimator's input buffer; dense blocks and context
inputs never concatenate.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .pwc_modules import (ContextNetwork, FeatureExtractor, FlowEstimatorDense, OccContextNetwork, OccEstimatorDense,
                          WarpingLayer, conv, flow_scales, initialize_msra)


class PWCFamily(nn.Module):
    IRR = False
    BI = False
    OCC = False

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
        if self.IRR:
            self.num_ch_in = self.dim_corr + 32 + 2
            self.flow_estimators = FlowEstimatorDense(self.num_ch_in)
            self.context_networks = ContextNetwork(self.num_ch_in + 448 + 2)
            if self.OCC:
                self.num_ch_in_occ = self.dim_corr + 32 + 1
                self.occ_estimators = OccEstimatorDense(self.num_ch_in_occ)
                self.occ_context_networks = OccContextNetwork(self.num_ch_in_occ + 448 + 1)
            self.conv_1x1 = nn.ModuleList([conv(c, 32, kernel_size=1, stride=1, dilation=1)
                                           for c in (196, 128, 96, 64, 32)])
        else:
            self.flow_estimators = nn.ModuleList()
            if self.OCC:
                self.occ_estimators = nn.ModuleList()
            for l, ch in enumerate(self.num_chs[::-1]):
                if l > self.output_level:
                    break
                self.flow_estimators.append(FlowEstimatorDense(self.dim_corr if l == 0 else self.dim_corr + ch + 2))
                if self.OCC:
                    self.occ_estimators.append(OccEstimatorDense(self.dim_corr if l == 0 else self.dim_corr + ch + 1))
            self.context_networks = ContextNetwork(self.dim_corr + 32 + 2 + 448 + 2)
            if self.OCC:
                self.context_networks_occ = OccContextNetwork(self.dim_corr + 32 + 1 + 448 + 1)
        self.corr_params = {"pad_size": self.search_range, "kernel_size": 1, "max_disp": self.search_range,
                            "stride1": 1, "stride2": 1, "corr_multiply": 1}
        initialize_msra(self.modules())

    # -------------------------------------------------------------------------------------------------------------
    def _cost_volume(self, l, feat, B, flow_up, out, height_im, width_im):
        """warp + cost volume + LeakyReLU of level l into ``out`` (uni: x1 vs x2; bi: both directions, 2B rows)."""
        if self.BI:
            f1, f2, shift = feat, feat, B
        else:
            f1, f2, shift = feat[:B], feat[B:], 0
        if l == 0:
            ops.correlation(f1, f2, out=out, shift=shift, slope=0.1)
        else:
            ops.warp_correlation(f1, f2, flow_up, height_im, width_im, self._div_flow, out=out, shift=shift, slope=0.1)

    def forward(self, input_dict, record=None):
        if self.training:
            raise RuntimeError(f"irr_b200.{type(self).__module__}: only the eval-mode forward is implemented (call .eval())")
        x1_raw, x2_raw = input_dict['input1'], input_dict['input2']
        B, _, height_im, width_im = x1_raw.shape
        NB = 2 * B if self.BI else B
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
                    flow_up = ops.empty(NB, 2, h, w, dev, zero=True)
                    occ_up = ops.empty(NB, 1, h, w, dev, zero=True) if self.OCC else None
                else:
                    flow_up = ops.resize_ac(flow, h, w)
                    occ_up = ops.resize_ac(occ, h, w) if self.OCC else None
                flow, occ = self.estimator_level(l, feat, flow_up, occ_up, height_im, width_im)
                if record is not None:
                    record[l] = {"flow": flow.clone()}
                    if self.OCC:
                        record[l]["occ"] = occ.clone()
            out = {'flow': ops.resize_ac(flow[:B], height_im, width_im, s_even=1.0 / df, s_odd=1.0 / df, pitched=False)}
            if self.OCC:
                out['occ'] = ops.resize_ac(occ[:B], height_im, width_im, pitched=False)
        return out

    def estimator_level(self, l, feat, flow_up, occ_up, height_im, width_im):
        """One pyramid level (the loop bodies cited above) given the level's inputs — the stage entry point the
        teacher-forced parity tests drive.

        feat   : (2B, C_l, h, w) rows [0,B) = x1 features, rows [B,2B) = x2 features
        flow_up: (NB, 2, h, w) previous flow resized to this level (zeros at l == 0); NB = 2B for the *_bi classes
                 (rows [0,B) forward, [B,2B) backward), else B
        occ_up : (NB, 1, h, w) or None (classes without the occlusion branch)
        returns (flow, occ-or-None) of this level."""
        feat, flow_up, occ_up = ops.pitched(feat), ops.pitched(flow_up), ops.pitched(occ_up)
        B = feat.shape[0] // 2
        NB = 2 * B if self.BI else B
        C = feat.shape[1]
        x = feat if self.BI else feat[:B]          # the feature each row's estimator sees (x1 | x2)
        if self.IRR:
            return self._level_irr(l, feat, x, B, NB, flow_up, occ_up, height_im, width_im)
        return self._level_plain(l, feat, x, B, NB, C, flow_up, occ_up, l == self.output_level, height_im, width_im)

    # pwcnet_irr.py:73-84, pwcnet_irr_bi.py:79-98, pwcnet_irr_occ.py:79-97
    def _level_irr(self, l, feat, x, B, NB, flow_up, occ_up, height_im, width_im):
        _, _, h, w = feat.shape
        dev = feat.device
        df = self._div_flow
        nf = self.num_ch_in
        buf_f = ops.empty(NB, 448 + nf + 2, h, w, dev)
        self._cost_volume(l, feat, B, flow_up, buf_f[:, 448:529], height_im, width_im)
        self.conv_1x1[l](x, out=buf_f[:, 529:561])
        su_l, sv_l = flow_scales(h, w, df, width_im, height_im, True)
        su_g, sv_g = flow_scales(h, w, df, width_im, height_im, False)
        ops.scale_channels(flow_up, out=buf_f[:, 561:563], s_even=su_l, s_odd=sv_l)
        occ = None
        if self.OCC:
            no = self.num_ch_in_occ
            buf_o = ops.empty(NB, 448 + no + 1, h, w, dev)
            ops.scale_channels(buf_f[:, 448:561], out=buf_o[:, 448:561])
            ops.scale_channels(occ_up, out=buf_o[:, 561:562])
        self.flow_estimators.forward_into(buf_f, out=buf_f[:, 563:565], addend=buf_f[:, 561:563])
        flow = self.context_networks(buf_f, addend=buf_f[:, 563:565])
        ops.scale_channels(flow, out=flow, s_even=su_g, s_odd=sv_g)
        if self.OCC:
            self.occ_estimators.forward_into(buf_o, out=buf_o[:, 562:563], addend=buf_o[:, 561:562])
            occ = self.occ_context_networks(buf_o, addend=buf_o[:, 562:563])
        return flow, occ

    # pwcnet_bi.py:82-98, pwcnet_occ.py:86-105, pwcnet_occ_bi.py:93-121
    def _level_plain(self, l, feat, x, B, NB, C, flow_up, occ_up, last, height_im, width_im):
        _, _, h, w = feat.shape
        dev = feat.device
        est = self.flow_estimators[l]
        buf_f = ops.empty(NB, est.total_ch + (2 if last else 0), h, w, dev)
        corr = buf_f[:, 448:529]
        self._cost_volume(l, feat, B, flow_up, corr, height_im, width_im)
        if l > 0:  # cat[corr, x, flow]
            ops.scale_channels(x, out=buf_f[:, 529:529 + C])
            ops.scale_channels(flow_up, out=buf_f[:, 529 + C:531 + C])
        occ = None
        if self.OCC:
            oest = self.occ_estimators[l]
            buf_o = ops.empty(NB, oest.total_ch + (1 if last else 0), h, w, dev)
            ops.scale_channels(corr, out=buf_o[:, 448:529])
            if l > 0:  # cat[corr, x1, occ] — x1 for BOTH directions (pwcnet_occ_bi.py:102-103, as written)
                ops.scale_channels(feat[:B], out=buf_o[:B, 529:529 + C])
                if self.BI:
                    ops.scale_channels(feat[:B], out=buf_o[B:, 529:529 + C])
                ops.scale_channels(occ_up, out=buf_o[:, 529 + C:530 + C])
        if not last:
            flow = est.forward_into(buf_f)
            if self.OCC:
                occ = oest.forward_into(buf_o)
        else:  # flow + context(cat[x_intm, flow])
            tail = buf_f[:, est.total_ch:est.total_ch + 2]
            est.forward_into(buf_f, out=tail)
            flow = self.context_networks(buf_f, addend=tail)
            if self.OCC:
                otail = buf_o[:, oest.total_ch:oest.total_ch + 1]
                oest.forward_into(buf_o, out=otail)
                occ = self.context_networks_occ(buf_o, addend=otail)
        return flow, occ
