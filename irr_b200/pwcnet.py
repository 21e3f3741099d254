"""Plain PWC-Net (per-level dense estimators, uni-directional) — eval-mode forward.

Drop-in for the reference's ``models/pwcnet.py`` ``PWCNet`` (BASELINE config 2): same constructor, parameter names
(``flow_estimators.{0-4}.*``, ``context_networks.*``) and ``forward({'input1','input2'}) -> {'flow'}`` (pwcnet.py:43-99).
The feature pyramid runs once on the stacked pair; warp+mask+cost volume+LeakyReLU is one kernel writing into the
estimator's input buffer; the dense block and the context input never concatenate."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .pwc_modules import ContextNetwork, FeatureExtractor, FlowEstimatorDense, WarpingLayer, initialize_msra


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
        self.flow_estimators = nn.ModuleList()
        self.dim_corr = (self.search_range * 2 + 1) ** 2
        for l, ch in enumerate(self.num_chs[::-1]):
            if l > self.output_level:
                break
            num_ch_in = self.dim_corr if l == 0 else self.dim_corr + ch + 2
            self.flow_estimators.append(FlowEstimatorDense(num_ch_in))
        self.context_networks = ContextNetwork(self.dim_corr + 32 + 2 + 448 + 2)
        self.corr_params = {"pad_size": self.search_range, "kernel_size": 1, "max_disp": self.search_range,
                            "stride1": 1, "stride2": 1, "corr_multiply": 1}
        initialize_msra(self.modules())

    def estimator_level(self, l, feat, flow_up, height_im, width_im, record=None):
        """One pass of pwcnet.py:62-92 for pyramid level l on the stacked pair.

        feat   : (2B, C_l, h, w) rows [0,B) = x1 features, rows [B,2B) = x2 features
        flow_up: (B, 2, h, w) the previous level's flow already resized to this level (ignored at l == 0)
        returns this level's flow (B, 2, h, w) — after the context network at the output level."""
        feat, flow_up = ops.pitched(feat), ops.pitched(flow_up)
        B = feat.shape[0] // 2
        df = self._div_flow
        x1, x2 = feat[:B], feat[B:]
        _, C, h, w = x1.shape
        est = self.flow_estimators[l]
        last = l == self.output_level
        buf = ops.empty(B, est.total_ch + (2 if last else 0), h, w, feat.device)
        corr = buf[:, 448:529]
        if l == 0:  # pwcnet.py:66-68,73-74
            ops.correlation(x1, x2, out=corr, slope=0.1)
        else:       # pwcnet.py:69-74
            ops.warp_correlation(x1, x2, flow_up, height_im, width_im, df, out=corr, slope=0.1)
            ops.scale_channels(x1, out=buf[:, 529:529 + C])          # pwcnet.py:80 cat[corr, x1, flow]
            ops.scale_channels(flow_up, out=buf[:, 529 + C:531 + C])
        if record is not None:
            record["corr"] = corr.clone()
        if not last:
            flow = est.forward_into(buf)
        else:  # pwcnet.py:85-88: flow + context(cat[x_intm, flow])
            tail = buf[:, est.total_ch:est.total_ch + 2]
            est.forward_into(buf, out=tail)
            if record is not None:
                record["flow_est"] = tail.clone()
            flow = self.context_networks(buf, addend=tail)
        if record is not None:
            record["flow"] = flow.clone()
        return flow

    def forward(self, input_dict, record=None):
        if self.training:
            raise RuntimeError("irr_b200.pwcnet: only the eval-mode forward is implemented (call .eval())")
        x1_raw, x2_raw = input_dict['input1'], input_dict['input2']
        B, _, height_im, width_im = x1_raw.shape
        df = self._div_flow
        # the kernels launch on the CURRENT device / stream: make the input's device current, as the reference's
        # correlation.py:21 does (a model on cuda:1 must work while cuda:0 is current)
        with torch.cuda.device_of(x1_raw), torch.no_grad():
            imgs = ops.stack_pair(x1_raw, x2_raw)  # (2B, 3, H, W), rows pitched when W % 4 != 0
            pyramid = self.feature_pyramid_extractor(imgs)
            flow = None
            for l, feat in enumerate(pyramid[:self.output_level + 1]):
                _, _, h, w = feat.shape
                if l > 0:
                    flow = ops.resize_ac(flow, h, w)
                rec_l = None
                if record is not None:
                    rec_l = record[l] = {}
                flow = self.estimator_level(l, feat, flow, height_im, width_im, rec_l)
            out = ops.resize_ac(flow, height_im, width_im, s_even=1.0 / df, s_odd=1.0 / df, pitched=False)  # pwcnet.py:97
        return {'flow': out}
