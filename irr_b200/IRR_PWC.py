"""IRR-PWC (bi-directional flow + occlusion, shared-weight iterative residual refinement) — eval-mode forward.

Drop-in for the reference's ``models/IRR_PWC.py`` ``PWCNet``: same constructor ``(args, div_flow=0.05)``, same
sub-module / parameter names (so ``saved_check_point/pwcnet/IRR-PWC_*`` load unchanged), same
``forward({'input1','input2'}) -> {'flow','occ'}`` contract (IRR_PWC.py:51-184, eval branch :176-184).

What is different is HOW the level loop (IRR_PWC.py:73-174) is executed:
  * the forward and backward directions are one 2B batch ([x1;x2] paired with [x2;x1] through a batch rotation inside
    the warp / correlation kernels) — half the launches, twice the parallelism at the 7x16 ... 28x64 levels;
  * the feature pyramid runs once on the 2B batch (the reference runs the extractor twice, IRR_PWC.py:58-59);
  * warp + mask + cost volume + LeakyReLU are ONE kernel writing straight into the estimator's input buffer;
  * every torch.cat of the dense blocks / context / refinement inputs is a channel slice of a pre-allocated buffer;
  * residual adds, the *0.1 of the occlusion up-sampler and the flow/occ skip connections are conv epilogues;
  * the as-executed flow scaling (rescale_flow mutates its argument at IRR_PWC.py:128-129) is
    written out explicitly: RefineFlow consumes ``flow_cont`` in GLOBAL units.
Training mode (per-level output lists) is out of scope.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .irr_modules import OccUpsampleNetwork, RefineFlow, RefineOcc
from .pwc_modules import (ContextNetwork, FeatureExtractor, FlowEstimatorDense, OccContextNetwork, OccEstimatorDense,
                          WarpingLayer, conv, flow_scales, initialize_msra)


import os

_SIDE_STREAMS = {}
_USE_SIDE_STREAM = os.environ.get("IRR_NO_SIDE_STREAM", "0") != "1"


def set_side_stream(enabled: bool) -> None:
    """Run the flow and occlusion branches of a level on two streams (default) or serially (per-kernel timing)."""
    global _USE_SIDE_STREAM
    _USE_SIDE_STREAM = bool(enabled)


def _side_stream(device):
    key = str(device)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device)
        _SIDE_STREAMS[key] = st
    return st


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
        self.occ_shuffle_upsample = OccUpsampleNetwork(11, 1)

        self.conv_1x1 = nn.ModuleList([conv(196, 32, kernel_size=1, stride=1, dilation=1),
                                       conv(128, 32, kernel_size=1, stride=1, dilation=1),
                                       conv(96, 32, kernel_size=1, stride=1, dilation=1),
                                       conv(64, 32, kernel_size=1, stride=1, dilation=1)])
        self.conv_1x1_1 = conv(16, 3, kernel_size=1, stride=1, dilation=1)

        self.refine_flow = RefineFlow(2 + 1 + 32)
        self.refine_occ = RefineOcc(1 + 32 + 32)
        self.corr_params = {"pad_size": self.search_range, "kernel_size": 1, "max_disp": self.search_range,
                            "stride1": 1, "stride2": 1, "corr_multiply": 1}
        # BASELINE config 5 ("mixed bf16 features"): "bf16" rounds the feature pyramid to bf16 values before the warp /
        # correlation / 1x1 convs (a capability the fp32-only reference does not have; default = the reference's fp32)
        self.feature_dtype = "fp32"
        self.bf16_storage = True   # with feature_dtype "bf16": the cost volumes read a packed bf16 copy of the features
        # Dead-branch elimination for the eval forward (off by default = the reference's executed work, layer for
        # layer).  In eval mode the reference returns only flow_f and occ_f (IRR_PWC.py:176-184); nothing that feeds
        # them reads the BACKWARD occlusion chain: occ_b enters only its own estimator / context / refinement /
        # up-sampling at the next level (:117-123, :141-145, :172).  With ``eval_prune_dead`` the occlusion branch runs
        # on the forward rows only (the backward FLOW chain stays: flow_b is warped into the level-5/6 occlusion
        # up-sampler, :157) — same outputs, ~30 % fewer conv FLOPs.
        self.eval_prune_dead = False
        initialize_msra(self.modules())

    def set_feature_dtype(self, dtype: str):
        if dtype not in ("fp32", "bf16"):
            raise ValueError("feature_dtype must be 'fp32' or 'bf16'")
        self.feature_dtype = dtype
        return self

    # ------------------------------------------------------------------------------------------------ stages
    def estimator_level(self, l, feat, flow_up, occ_up, imgs, height_im, width_im, record=None, feat16=None):
        """One pass of IRR_PWC.py:75-147 for pyramid level l <= 4 on the 2B batch.

        feat   : (2B, C_l, h, w)  rows [0,B) = x1 features, rows [B,2B) = x2 features
        flow_up: (2B, 2, h, w) flow in GLOBAL units already resized to this level (zeros at l == 0); rows [0,B) forward
        occ_up : (2B, 1, h, w)
        imgs   : (2B, 3, H, W)
        feat16 : optional packed bf16 copy of ``feat`` (feature_dtype "bf16"): the cost volume reads it instead — half the
                 f1 / f2 bytes, bit-identical results (ops.round_bf16_store)
        returns (flow, occ) for this level (flow in GLOBAL units), each on the 2B batch."""
        feat, flow_up, occ_up, imgs = ops.pitched(feat), ops.pitched(flow_up), ops.pitched(occ_up), ops.pitched(imgs)
        B2, C, h, w = feat.shape
        B = B2 // 2
        dev = feat.device
        df = self._div_flow
        nf, no = self.num_ch_in_flo, self.num_ch_in_occ  # 115, 114
        rec = (lambda k, v: record.__setitem__(k, v.clone())) if record is not None else (lambda k, v: None)

        # estimator input buffers: [448 dense outputs | corr 81 | x_1by1 32 | flow 2 or occ 1 | est 2 or 1]
        buf_f = ops.empty(B2, 448 + nf + 2, h, w, dev)
        BO = B if self.eval_prune_dead else B2   # rows the occlusion branch runs on
        buf_o = ops.empty(BO, 448 + no + 1, h, w, dev)
        corr = buf_f[:, 448:529]
        cf = feat if feat16 is None else feat16
        if l == 0:  # IRR_PWC.py:78-80,90-95 — no warp at the coarsest level
            ops.correlation(cf, cf, out=corr, shift=B, slope=0.1)
        else:       # :86-95 fused
            ops.warp_correlation(cf, cf, flow_up, height_im, width_im, df, out=corr, shift=B, slope=0.1)
        x1by1 = buf_f[:, 529:561]
        if l != self.output_level:  # :97-102
            self.conv_1x1[l](feat, out=x1by1)
        else:
            ops.scale_channels(feat, out=x1by1)
        ops.scale_channels(buf_f[:BO, 448:561], out=buf_o[:, 448:561])  # shared [corr | x_1by1] block
        su_l, sv_l = flow_scales(h, w, df, width_im, height_im, True)
        ops.scale_channels(flow_up, out=buf_f[:, 561:563], s_even=su_l, s_odd=sv_l)  # :105-106 to_local
        ops.scale_channels(occ_up[:BO], out=buf_o[:, 561:562])
        rec("corr", corr); rec("x_1by1", x1by1)

        # The flow branch (:108-114) and the occlusion branch (:117-123) are independent until the refinement: they run
        # on two streams (fork/join, also inside a captured CUDA graph) so that at the coarse levels — 7..56 persistent
        # CTAs per conv — two convs share the 148 SMs instead of running back to back.
        main = torch.cuda.current_stream()
        side = _side_stream(dev) if _USE_SIDE_STREAM else main
        side.wait_stream(main)
        with torch.cuda.stream(side):
            self.occ_estimators.forward_into(buf_o, out=buf_o[:, 562:563], addend=buf_o[:, 561:562])
            occ_cont = self.occ_context_networks(buf_o, addend=buf_o[:, 562:563])
        self.flow_estimators.forward_into(buf_f, out=buf_f[:, 563:565], addend=buf_f[:, 561:563])
        rec("flow_est", buf_f[:, 563:565])
        flow_cont = self.context_networks(buf_f, addend=buf_f[:, 563:565])
        rec("flow_cont", flow_cont)
        flow = self.refine_flow_stage(flow_cont, x1by1, imgs, height_im, width_im)
        rec("flow", flow)
        main.wait_stream(side)  # join; occ_cont (allocated on `side`) is only reused by `side` after the next fork
        rec("occ_cont", occ_cont)
        occ = self.refine_occ_stage(occ_cont, x1by1, flow, height_im, width_im)
        rec("occ", occ)
        return flow, occ

    def refine_flow_stage(self, flow_cont, x1by1, imgs, height_im, width_im):
        """IRR_PWC.py:126-138.  ``flow_cont`` arrives in LOCAL units and is converted IN PLACE to global units first —
        exactly what the un-rebound rescale_flow call at :128-129 leaves behind, so RefineFlow (:132-133)
        sees global units; its output is then scaled to global once more (:137-138)."""
        flow_cont_arg = flow_cont
        flow_cont, x1by1, imgs = ops.pitched(flow_cont), ops.pitched(x1by1), ops.pitched(imgs)
        B2, _, h, w = flow_cont.shape
        B = B2 // 2
        df = self._div_flow
        su_g, sv_g = flow_scales(h, w, df, width_im, height_im, False)
        img_r = ops.resize_ac(imgs, h, w)  # :126-127
        ops.scale_channels(flow_cont, out=flow_cont, s_even=su_g, s_odd=sv_g)
        rf_in = ops.empty(B2, 35, h, w, flow_cont.device)
        diff = ops.warp(img_r, flow_cont, height_im, width_im, df, minuend=img_r, shift=B)  # img_a - warp(img_b)
        ops.sub_spatial_mean(flow_cont, out=rf_in[:, 0:2])
        ops.channel_l2norm(diff, out=rf_in[:, 2:3])
        ops.scale_channels(x1by1, out=rf_in[:, 3:35])
        flow = self.refine_flow.gather(rf_in, flow_cont)
        ops.scale_channels(flow, out=flow, s_even=su_g, s_odd=sv_g)
        if flow_cont is not flow_cont_arg:   # a dense argument was re-pitched: keep the in-place side effect visible (F6)
            ops.scale_channels(flow_cont, out=flow_cont_arg)
        return flow

    def refine_occ_stage(self, occ_cont, x1by1, flow, height_im, width_im):
        """IRR_PWC.py:141-145: occ = RefineOcc(occ_cont, x_1by1, x_1by1 - warp(other x_1by1, flow))."""
        occ_cont, x1by1, flow = ops.pitched(occ_cont), ops.pitched(x1by1), ops.pitched(flow)
        BO, _, h, w = occ_cont.shape
        B = x1by1.shape[0] // 2
        ro_in = ops.empty(BO, 65, h, w, occ_cont.device)
        ops.scale_channels(occ_cont, out=ro_in[:, 0:1])
        ops.scale_channels(x1by1[:BO], out=ro_in[:, 1:33])
        if BO == 2 * B:
            ops.warp(x1by1, flow, height_im, width_im, self._div_flow, minuend=x1by1, shift=B, out=ro_in[:, 33:65])
        else:  # forward rows only (eval_prune_dead): x1_1by1 - warp(x2_1by1, flow_f)
            ops.warp(x1by1[B:], flow[:B], height_im, width_im, self._div_flow, minuend=x1by1[:B], out=ro_in[:, 33:65])
        return self.refine_occ.gather(ro_in, occ_cont)

    def upsample_level(self, l, feat, flow, occ_prev, height_im, width_im, record=None):
        """IRR_PWC.py:150-174 for l in {5, 6}: occlusion up-sampling guided by warped features / flows."""
        feat, flow, occ_prev = ops.pitched(feat), ops.pitched(flow), ops.pitched(occ_prev)
        B2, C, h, w = feat.shape
        B = B2 // 2
        df = self._div_flow
        if occ_prev.shape[0] == B:  # eval_prune_dead: forward rows only (IRR_PWC.py:155,157,172 for occ_f)
            x_in = ops.empty(B, 11, h, w, feat.device)
            ops.upsample_nearest2x(occ_prev, h, w, out=x_in[:, 0:1])
            if l != self.num_levels - 1:
                self.conv_1x1_1(feat[:B], out=x_in[:, 1:4])
                xw = ops.warp(feat[B:], flow[:B], height_im, width_im, df)
                self.conv_1x1_1(xw, out=x_in[:, 4:7])
            else:
                ops.scale_channels(feat[:B], out=x_in[:, 1:4])
                ops.warp(feat[B:], flow[:B], height_im, width_im, df, out=x_in[:, 4:7])
            ops.scale_channels(flow[:B], out=x_in[:, 7:9])
            ops.warp(flow[B:], flow[:B], height_im, width_im, df, out=x_in[:, 9:11])  # flow_b warped by flow_f
        else:
            x_in = ops.empty(B2, 11, h, w, feat.device)
            ops.upsample_nearest2x(occ_prev, h, w, out=x_in[:, 0:1])
            if l != self.num_levels - 1:  # :160-164
                self.conv_1x1_1(feat, out=x_in[:, 1:4])
                xw = ops.warp(feat, flow, height_im, width_im, df, shift=B)
                self.conv_1x1_1(xw, out=x_in[:, 4:7])
            else:
                ops.scale_channels(feat, out=x_in[:, 1:4])
                ops.warp(feat, flow, height_im, width_im, df, shift=B, out=x_in[:, 4:7])
            ops.scale_channels(flow, out=x_in[:, 7:9])
            ops.warp(flow, flow, height_im, width_im, df, shift=B, out=x_in[:, 9:11])  # flow_b warped by flow_f, and v.v.
        occ = self.occ_shuffle_upsample.forward_into(x_in)
        if record is not None:
            record["occ"] = occ.clone()
        return occ

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, input_dict, record=None):
        if self.training:
            raise RuntimeError("irr_b200.IRR_PWC: only the eval-mode forward is implemented (call .eval())")
        x1_raw = input_dict['input1']
        x2_raw = input_dict['input2']
        B, _, height_im, width_im = x1_raw.shape
        # the kernels launch on the CURRENT device / stream: make the input's device current, as the reference's
        # correlation.py:21 does (a model on cuda:1 must work while cuda:0 is current)
        with torch.cuda.device_of(x1_raw), torch.no_grad():
            imgs = ops.stack_pair(x1_raw, x2_raw)  # (2B, 3, H, W), rows pitched when W % 4 != 0
            pyramid = self.feature_pyramid_extractor(imgs)
            pyr16 = [None] * (len(pyramid) + 1)
            if self.feature_dtype == "bf16":
                # bf16 VALUES in the fp32 layout for the convs / warps; for the levels that feed a cost volume also a packed
                # bf16 copy (STORAGE), written in the same pass, which the correlation kernels read
                for l, f in enumerate(pyramid):
                    if l <= self.output_level and self.bf16_storage:
                        pyr16[l] = ops.round_bf16_store(f, round_in_place=True)
                    else:
                        ops.round_bf16(f, out=f)
            pyramid = pyramid + [imgs]
            flow = occ = None
            for l, feat in enumerate(pyramid):
                _, _, h, w = feat.shape
                rec_l = None
                if record is not None:
                    rec_l = {}
                    record[l] = rec_l
                if l <= self.output_level:
                    if l == 0:
                        flow_up = ops.empty(2 * B, 2, h, w, imgs.device, zero=True)
                        occ_up = ops.empty(2 * B, 1, h, w, imgs.device, zero=True)
                    else:
                        flow_up = ops.resize_ac(flow, h, w)
                        occ_up = ops.resize_ac(occ, h, w)
                    if rec_l is not None:
                        rec_l["feat"] = feat.clone(); rec_l["flow_up"] = flow_up.clone(); rec_l["occ_up"] = occ_up.clone()
                    flow, occ = self.estimator_level(l, feat, flow_up, occ_up, imgs, height_im, width_im, rec_l, pyr16[l])
                else:
                    flow = ops.resize_ac(flow, h, w)
                    if rec_l is not None:
                        rec_l["feat"] = feat.clone(); rec_l["flow_up"] = flow.clone(); rec_l["occ_in"] = occ.clone()
                    occ = self.upsample_level(l, feat, flow, occ, height_im, width_im, rec_l)
            out_flow = ops.resize_ac(flow[:B], height_im, width_im, s_even=1.0 / self._div_flow,
                                     s_odd=1.0 / self._div_flow, pitched=False)  # :176 (user-facing: dense rows)
            out_occ = ops.resize_ac(occ[:B], height_im, width_im, pitched=False)  # :177
        return {'flow': out_flow, 'occ': out_occ}
