"""CPU/torch restatement of the visinf/irr PWC-family inference path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``irr_b200/`` may import this file; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl
reference`` legs of ``bench.py`` as the *checker* (and as the timed CPU reference),
never as the product path.

Parity status: the reference repo ships no tests ("parity unpinned" by reference-owned
fixtures).  This restatement is pinned instead against the reference's
own Python implementation run in the build container (``oracle/gen_golden.py`` imports
``/root/reference`` and writes ``tests/golden/*.npz``; ``tests/test_oracle_vs_reference.py``
re-checks live whenever ``/root/reference`` is mounted).

Everything is written functionally over a flat ``{name: tensor}`` parameter dict that
uses the reference's ``state_dict`` names (``models/IRR_PWC.py:26-45``), in plain torch
fp32 library ops, device agnostic (runs on CPU, or on the GPU as the "reference on the
same GPU" arm).  Each function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

LEAKY = 0.1  # models/pwc_modules.py:13
PYR_CHS = [3, 16, 32, 64, 96, 128, 196]  # models/IRR_PWC.py:20
SEARCH = 4  # models/IRR_PWC.py:19
OUT_LEVEL = 4  # models/IRR_PWC.py:21


# --------------------------------------------------------------------------- blocks
def conv_block(p: Params, name: str, x, stride=1, dilation=1, relu=True):
    """``conv()`` of models/pwc_modules.py:8-19: Conv2d(pad=((k-1)*dil)//2) [+ LeakyReLU(0.1)]."""
    w = p[name + ".0.weight"]
    b = p[name + ".0.bias"]
    k = w.shape[-1]
    y = F.conv2d(x, w, b, stride=stride, padding=((k - 1) * dilation) // 2, dilation=dilation)
    return F.leaky_relu(y, LEAKY) if relu else y


def feature_pyramid(p: Params, x, prefix="feature_pyramid_extractor") -> List[torch.Tensor]:
    """models/pwc_modules.py:85-104 — six [3x3 s2, 3x3 s1] pairs, returned coarse -> fine."""
    out = []
    for l in range(len(PYR_CHS) - 1):
        x = conv_block(p, f"{prefix}.convs.{l}.0", x, stride=2)
        x = conv_block(p, f"{prefix}.convs.{l}.1", x)
        out.append(x)
    return out[::-1]


def cost_volume(f1, f2, max_disp=SEARCH):
    """models/pwc_modules.py:42-62: out[:, (dy+d)*(2d+1)+(dx+d)] = mean_c f1 * shift(f2, dy, dx), zero padded."""
    B, C, H, W = f1.shape
    f2p = F.pad(f2, (max_disp,) * 4)
    planes = []
    for i in range(2 * max_disp + 1):
        for j in range(2 * max_disp + 1):
            planes.append((f1 * f2p[:, :, i:i + H, j:j + W]).mean(dim=1, keepdim=True))
    return torch.cat(planes, dim=1)


def host_linspace(n: int) -> torch.Tensor:
    """models/pwc_modules.py:108-109: the reference builds the base grid with CPU torch.linspace."""
    return torch.linspace(-1.0, 1.0, n)


def sampling_grid(flow, height_im, width_im, div_flow, lin_x=None, lin_y=None):
    """models/pwc_modules.py:107-126 — base grid + flow*2/max(dim-1,1)/div_flow, as N x H x W x 2."""
    B, _, H, W = flow.shape
    lx = (host_linspace(W) if lin_x is None else lin_x).to(flow.device).view(1, 1, W)
    ly = (host_linspace(H) if lin_y is None else lin_y).to(flow.device).view(1, H, 1)
    gx = lx + flow[:, 0] * 2 / max(width_im - 1, 1) / div_flow
    gy = ly + flow[:, 1] * 2 / max(height_im - 1, 1) / div_flow
    return torch.stack([gx, gy], dim=-1)


def warp(x, flow, height_im, width_im, div_flow, lin_x=None, lin_y=None):
    """models/pwc_modules.py:119-133 — bilinear grid_sample (align_corners) times the hard validity mask."""
    grid = sampling_grid(flow, height_im, width_im, div_flow, lin_x, lin_y)
    xw = F.grid_sample(x, grid, align_corners=True)
    m = F.grid_sample(torch.ones_like(x), grid, align_corners=True)
    return xw * (m >= 1.0).to(x.dtype)


def resize_ac(x, like):
    """models/pwc_modules.py:65-67 — bilinear, align_corners=True, to the H x W of ``like``."""
    return F.interpolate(x, size=list(like.shape[2:]), mode="bilinear", align_corners=True)


def flow_scales(level_hw, height_im, width_im, div_flow, to_local):
    """The two Python floats of models/pwc_modules.py:71-76 (computed in double, cast by the mul)."""
    h, w = level_hw
    if to_local:
        return float(w / width_im / div_flow), float(h / height_im / div_flow)
    return float(width_im * div_flow / w), float(height_im * div_flow / h)


def scale_flow(flow, su, sv):
    """Value semantics of rescale_flow (models/pwc_modules.py:78-82); aliasing handled by callers."""
    return torch.cat([flow[:, 0:1] * su, flow[:, 1:2] * sv], dim=1)


def dense_estimator(p: Params, prefix: str, x):
    """FlowEstimatorDense / OccEstimatorDense, models/pwc_modules.py:153-170,190-207."""
    for i in range(1, 6):
        x = torch.cat([conv_block(p, f"{prefix}.conv{i}", x), x], dim=1)
    return x, conv_block(p, f"{prefix}.conv_last", x, relu=False)


CTX_DIL = [1, 2, 4, 8, 16, 1, 1]  # models/pwc_modules.py:215-221


def context_net(p: Params, prefix: str, x):
    """ContextNetwork / OccContextNetwork, models/pwc_modules.py:210-243."""
    for i, d in enumerate(CTX_DIL):
        x = conv_block(p, f"{prefix}.convs.{i}", x, dilation=d, relu=(i != 6))
    return x


def _kernel_gather(src, logits):
    """irr_modules.py:86-104 — softmax(-feat^2) over 9 taps applied to the replicate-padded 3x3 window."""
    B, C, H, W = src.shape
    k = torch.softmax(-logits ** 2, dim=1).reshape(B, 9, H * W)
    sp = F.pad(src, (1, 1, 1, 1), mode="replicate")
    outs = []
    for c in range(C):  # per channel, reduce a (B, 9, H*W) stack over dim 1 exactly like the reference does
        win = F.unfold(sp[:, c:c + 1], kernel_size=3)
        outs.append(torch.sum(win * k, dim=1).view(B, 1, H, W))
    return torch.cat(outs, dim=1)


def _refine_convs(p: Params, prefix: str, x):
    for i in range(7):  # irr_modules.py:72-80 — all seven convs keep the LeakyReLU
        x = conv_block(p, f"{prefix}.convs.{i}", x)
    return x


def refine_flow(p: Params, flow, diff_img, feature, prefix="refine_flow"):
    """RefineFlow.forward, models/irr_modules.py:82-104."""
    flow_m = flow - flow.mean(2).mean(2)[:, :, None, None]
    nrm = torch.norm(diff_img, p=2, dim=1, keepdim=True)
    logits = _refine_convs(p, prefix, torch.cat([flow_m, nrm, feature], dim=1))
    return _kernel_gather(flow, logits)


def refine_occ(p: Params, occ, feat1, feat2, prefix="refine_occ"):
    """RefineOcc.forward, models/irr_modules.py:126-138."""
    logits = _refine_convs(p, prefix, torch.cat([occ, feat1, feat2], dim=1))
    return _kernel_gather(occ, logits)


def upsample_x2(x, like):
    """upsample_factor2, models/irr_modules.py:21-27."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    if tuple(x.shape[2:]) != tuple(like.shape[2:]):
        x = F.interpolate(x, size=list(like.shape[2:]), mode="bilinear", align_corners=False)
    return x


def occ_upsample(p: Params, occ, x, prefix="occ_shuffle_upsample"):
    """OccUpsampleNetwork.forward, models/irr_modules.py:46-56 (res_convs weights shared 3x)."""
    occ = upsample_x2(occ, x)
    x_init = conv_block(p, f"{prefix}.init_conv", torch.cat([occ, x], dim=1))
    r = x_init
    for _ in range(3):
        t = conv_block(p, f"{prefix}.res_convs.0", r)
        t = conv_block(p, f"{prefix}.res_convs.1", t, relu=False)
        r = r + t * 0.1
    x_init = x_init + conv_block(p, f"{prefix}.res_end_conv", r)
    return conv_block(p, f"{prefix}.out_convs", x_init) + occ


# --------------------------------------------------------------------------- models
def irr_pwc_forward(p: Params, img1, img2, div_flow=0.05, record: Optional[dict] = None, feature_bf16: bool = False):
    """Eval-mode ``IRR_PWC.PWCNet.forward`` (models/IRR_PWC.py:51-184), as executed.

    Note F6: models/IRR_PWC.py:128-129 call rescale_flow on ``flow_cont_*``
    without rebinding; the in-place ``u *= u_scale`` (pwc_modules.py:78-80) leaves the
    tensor in global units for lines 132-133.  That is written out explicitly below.
    ``record`` (optional dict) receives every stage tensor for teacher-forced tests.
    """
    B, _, Him, Wim = img1.shape
    pyr1 = feature_pyramid(p, img1)
    pyr2 = feature_pyramid(p, img2)
    if feature_bf16:  # BASELINE config 5: the feature pyramid is cast to bf16 before the warp / correlation (not in the
        # reference, which is fp32 only): bf16 values carried in fp32 tensors, everything downstream unchanged
        pyr1 = [f.bfloat16().float() for f in pyr1]
        pyr2 = [f.bfloat16().float() for f in pyr2]
    pyr1, pyr2 = pyr1 + [img1], pyr2 + [img2]
    h0, w0 = pyr1[0].shape[2:]
    z = lambda c: torch.zeros(B, c, h0, w0, dtype=img1.dtype, device=img1.device)
    flow_f, flow_b, occ_f, occ_b = z(2), z(2), z(1), z(1)
    rec = (lambda k, v: record.__setitem__(k, v.clone())) if record is not None else (lambda k, v: None)

    for l, (x1, x2) in enumerate(zip(pyr1, pyr2)):
        rec(f"l{l}.x1", x1); rec(f"l{l}.x2", x2)
        if l <= OUT_LEVEL:
            if l == 0:
                x2w, x1w = x2, x1
            else:  # IRR_PWC.py:82-87
                flow_f, flow_b = resize_ac(flow_f, x1), resize_ac(flow_b, x2)
                occ_f, occ_b = resize_ac(occ_f, x1), resize_ac(occ_b, x2)
                x2w = warp(x2, flow_f, Him, Wim, div_flow)
                x1w = warp(x1, flow_b, Him, Wim, div_flow)
            rec(f"l{l}.flow_up_f", flow_f); rec(f"l{l}.flow_up_b", flow_b)
            rec(f"l{l}.occ_up_f", occ_f); rec(f"l{l}.occ_up_b", occ_b)
            corr_f = F.leaky_relu(cost_volume(x1, x2w), LEAKY)  # IRR_PWC.py:90-95
            corr_b = F.leaky_relu(cost_volume(x2, x1w), LEAKY)
            rec(f"l{l}.corr_f", corr_f); rec(f"l{l}.corr_b", corr_b)
            if l != OUT_LEVEL:  # IRR_PWC.py:97-102
                x1_1, x2_1 = conv_block(p, f"conv_1x1.{l}", x1), conv_block(p, f"conv_1x1.{l}", x2)
            else:
                x1_1, x2_1 = x1, x2
            rec(f"l{l}.x1_1by1", x1_1); rec(f"l{l}.x2_1by1", x2_1)
            hw = x1.shape[2:]
            su_l, sv_l = flow_scales(hw, Him, Wim, div_flow, True)
            su_g, sv_g = flow_scales(hw, Him, Wim, div_flow, False)
            flow_f, flow_b = scale_flow(flow_f, su_l, sv_l), scale_flow(flow_b, su_l, sv_l)  # :105-106

            xi_f, res_f = dense_estimator(p, "flow_estimators", torch.cat([corr_f, x1_1, flow_f], 1))
            xi_b, res_b = dense_estimator(p, "flow_estimators", torch.cat([corr_b, x2_1, flow_b], 1))
            est_f, est_b = flow_f + res_f, flow_b + res_b  # :110-111
            rec(f"l{l}.flow_est_f", est_f); rec(f"l{l}.flow_est_b", est_b)
            cont_f = est_f + context_net(p, "context_networks", torch.cat([xi_f, est_f], 1))  # :113-114
            cont_b = est_b + context_net(p, "context_networks", torch.cat([xi_b, est_b], 1))
            rec(f"l{l}.flow_cont_f", cont_f); rec(f"l{l}.flow_cont_b", cont_b)

            xo_f, ores_f = dense_estimator(p, "occ_estimators", torch.cat([corr_f, x1_1, occ_f], 1))  # :117-118
            xo_b, ores_b = dense_estimator(p, "occ_estimators", torch.cat([corr_b, x2_1, occ_b], 1))
            oest_f, oest_b = occ_f + ores_f, occ_b + ores_b
            ocont_f = oest_f + context_net(p, "occ_context_networks", torch.cat([xo_f, oest_f], 1))  # :122-123
            ocont_b = oest_b + context_net(p, "occ_context_networks", torch.cat([xo_b, oest_b], 1))
            rec(f"l{l}.occ_cont_f", ocont_f); rec(f"l{l}.occ_cont_b", ocont_b)

            img1_r, img2_r = resize_ac(img1, flow_f), resize_ac(img2, flow_b)  # :126-127
            # :128-129 — in-place rescale: cont_* are in GLOBAL units from here on (F6)
            cont_f, cont_b = scale_flow(cont_f, su_g, sv_g), scale_flow(cont_b, su_g, sv_g)
            img2_w = warp(img2_r, cont_f, Him, Wim, div_flow)
            img1_w = warp(img1_r, cont_b, Him, Wim, div_flow)
            rec(f"l{l}.flow_cont_glob_f", cont_f); rec(f"l{l}.flow_cont_glob_b", cont_b)
            flow_f = refine_flow(p, cont_f, img1_r - img2_w, x1_1)  # :132-133
            flow_b = refine_flow(p, cont_b, img2_r - img1_w, x2_1)
            # :135-136 rescale cont_* once more — only feeds the train-mode output list
            flow_f, flow_b = scale_flow(flow_f, su_g, sv_g), scale_flow(flow_b, su_g, sv_g)  # :137-138
            rec(f"l{l}.flow_f", flow_f); rec(f"l{l}.flow_b", flow_b)
            x2_1w = warp(x2_1, flow_f, Him, Wim, div_flow)  # :141-142
            x1_1w = warp(x1_1, flow_b, Him, Wim, div_flow)
            occ_f = refine_occ(p, ocont_f, x1_1, x1_1 - x2_1w)  # :144-145
            occ_b = refine_occ(p, ocont_b, x2_1, x2_1 - x1_1w)
            rec(f"l{l}.occ_f", occ_f); rec(f"l{l}.occ_b", occ_b)
        else:  # IRR_PWC.py:150-174
            flow_f, flow_b = resize_ac(flow_f, x1), resize_ac(flow_b, x2)
            x2w = warp(x2, flow_f, Him, Wim, div_flow)
            x1w = warp(x1, flow_b, Him, Wim, div_flow)
            fb_w = warp(flow_b, flow_f, Him, Wim, div_flow)
            ff_w = warp(flow_f, flow_b, Him, Wim, div_flow)
            if l != 6:
                c = lambda t: conv_block(p, "conv_1x1_1", t)
                x1_in, x2_in, x1w_in, x2w_in = c(x1), c(x2), c(x1w), c(x2w)
            else:
                x1_in, x2_in, x1w_in, x2w_in = x1, x2, x1w, x2w
            rec(f"l{l}.occ_in_f", occ_f); rec(f"l{l}.occ_in_b", occ_b)
            rec(f"l{l}.flow_up_f", flow_f); rec(f"l{l}.flow_up_b", flow_b)
            occ_f, occ_b = (occ_upsample(p, occ_f, torch.cat([x1_in, x2w_in, flow_f, fb_w], 1)),
                            occ_upsample(p, occ_b, torch.cat([x2_in, x1w_in, flow_b, ff_w], 1)))
            rec(f"l{l}.occ_f", occ_f); rec(f"l{l}.occ_b", occ_b)

    out = {"flow": resize_ac(flow_f, img1) * (1.0 / div_flow), "occ": resize_ac(occ_f, img1)}  # :176-177
    return out


def pwcnet_forward(p: Params, img1, img2, div_flow=0.05, record: Optional[dict] = None):
    """Eval-mode ``pwcnet.PWCNet.forward`` (models/pwcnet.py:43-99): per-level estimators, uni-directional."""
    B, _, Him, Wim = img1.shape
    pyr1 = feature_pyramid(p, img1) + [img1]
    pyr2 = feature_pyramid(p, img2) + [img2]
    h0, w0 = pyr1[0].shape[2:]
    flow = torch.zeros(B, 2, h0, w0, dtype=img1.dtype, device=img1.device)
    rec = (lambda k, v: record.__setitem__(k, v.clone())) if record is not None else (lambda k, v: None)
    for l, (x1, x2) in enumerate(zip(pyr1, pyr2)):
        if l == 0:
            x2w = x2
        else:
            flow = resize_ac(flow, x1)
            x2w = warp(x2, flow, Him, Wim, div_flow)
        rec(f"l{l}.x1", x1); rec(f"l{l}.x2", x2); rec(f"l{l}.flow_up", flow)
        corr = F.leaky_relu(cost_volume(x1, x2w), LEAKY)
        rec(f"l{l}.corr", corr)
        inp = corr if l == 0 else torch.cat([corr, x1, flow], 1)
        x_intm, flow = dense_estimator(p, f"flow_estimators.{l}", inp)
        if l == OUT_LEVEL:
            flow = flow + context_net(p, "context_networks", torch.cat([x_intm, flow], 1))
            rec(f"l{l}.flow", flow)
            break
        rec(f"l{l}.flow", flow)
    return {"flow": resize_ac(flow, img1) * (1.0 / div_flow)}


def pwcnet_irr_occ_bi_forward(p: Params, img1, img2, div_flow=0.05, record: Optional[dict] = None):
    """Eval-mode ``pwcnet_irr_occ_bi.PWCNet.forward`` (models/pwcnet_irr_occ_bi.py:43-133)."""
    B, _, Him, Wim = img1.shape
    pyr1 = feature_pyramid(p, img1) + [img1]
    pyr2 = feature_pyramid(p, img2) + [img2]
    h0, w0 = pyr1[0].shape[2:]
    z = lambda c: torch.zeros(B, c, h0, w0, dtype=img1.dtype, device=img1.device)
    flow_f, flow_b, occ_f, occ_b = z(2), z(2), z(1), z(1)
    rec = (lambda k, v: record.__setitem__(k, v.clone())) if record is not None else (lambda k, v: None)
    for l, (x1, x2) in enumerate(zip(pyr1, pyr2)):
        if l == 0:
            x2w, x1w = x2, x1
        else:
            flow_f, flow_b = resize_ac(flow_f, x1), resize_ac(flow_b, x2)
            occ_f, occ_b = resize_ac(occ_f, x1), resize_ac(occ_b, x2)
            x2w = warp(x2, flow_f, Him, Wim, div_flow)
            x1w = warp(x1, flow_b, Him, Wim, div_flow)
        rec(f"l{l}.x1", x1); rec(f"l{l}.x2", x2)
        rec(f"l{l}.flow_up_f", flow_f); rec(f"l{l}.flow_up_b", flow_b)
        rec(f"l{l}.occ_up_f", occ_f); rec(f"l{l}.occ_up_b", occ_b)
        corr_f = F.leaky_relu(cost_volume(x1, x2w), LEAKY)
        corr_b = F.leaky_relu(cost_volume(x2, x1w), LEAKY)
        hw = x1.shape[2:]
        su_l, sv_l = flow_scales(hw, Him, Wim, div_flow, True)
        su_g, sv_g = flow_scales(hw, Him, Wim, div_flow, False)
        flow_f, flow_b = scale_flow(flow_f, su_l, sv_l), scale_flow(flow_b, su_l, sv_l)
        x1_1, x2_1 = conv_block(p, f"conv_1x1.{l}", x1), conv_block(p, f"conv_1x1.{l}", x2)
        xi_f, res_f = dense_estimator(p, "flow_estimators", torch.cat([corr_f, x1_1, flow_f], 1))
        xi_b, res_b = dense_estimator(p, "flow_estimators", torch.cat([corr_b, x2_1, flow_b], 1))
        flow_f, flow_b = flow_f + res_f, flow_b + res_b
        flow_f = flow_f + context_net(p, "context_networks", torch.cat([xi_f, flow_f], 1))
        flow_b = flow_b + context_net(p, "context_networks", torch.cat([xi_b, flow_b], 1))
        flow_f, flow_b = scale_flow(flow_f, su_g, sv_g), scale_flow(flow_b, su_g, sv_g)
        xo_f, ores_f = dense_estimator(p, "occ_estimators", torch.cat([corr_f, x1_1, occ_f], 1))
        xo_b, ores_b = dense_estimator(p, "occ_estimators", torch.cat([corr_b, x2_1, occ_b], 1))
        occ_f, occ_b = occ_f + ores_f, occ_b + ores_b
        occ_f = occ_f + context_net(p, "occ_context_networks", torch.cat([xo_f, occ_f], 1))
        occ_b = occ_b + context_net(p, "occ_context_networks", torch.cat([xo_b, occ_b], 1))
        rec(f"l{l}.flow_f", flow_f); rec(f"l{l}.flow_b", flow_b)
        rec(f"l{l}.occ_f", occ_f); rec(f"l{l}.occ_b", occ_b)
        if l == OUT_LEVEL:
            break
    return {"flow": resize_ac(flow_f, img1) * (1.0 / div_flow), "occ": resize_ac(occ_f, img1)}



# The six remaining ablation classes of the PWC family.  One restatement parameterised by the three
# switches the reference spells out in six files:
#   irr : shared estimators + conv_1x1 + rescale_flow, context at every level      (pwcnet_irr*.py) vs per-level
#         estimators on cat[corr, x, flow], context only at the output level       (pwcnet.py, pwcnet_bi/occ/occ_bi.py)
#   bi  : forward and backward flow with the same weights                          (*_bi.py)
#   occ : an occlusion branch next to the flow branch                              (*_occ*.py)
FAMILY = {"PWCNet_bi": (False, True, False), "PWCNet_occ": (False, False, True), "PWCNet_occ_bi": (False, True, True),
          "PWCNet_irr": (True, False, False), "PWCNet_irr_bi": (True, True, False), "PWCNet_irr_occ": (True, False, True)}


def pwc_family_forward(model: str, p: Params, img1, img2, div_flow=0.05, record: Optional[dict] = None):
    """Eval-mode forwards of models/pwcnet_bi.py:41-109, pwcnet_occ.py:49-117, pwcnet_occ_bi.py:49-132,
    pwcnet_irr.py:43-97, pwcnet_irr_bi.py:43-111, pwcnet_irr_occ.py:47-112 — as executed, including
    pwcnet_occ_bi.py:103, which feeds x1 (not x2) to the BACKWARD occlusion estimator."""
    irr, bi, occ = FAMILY[model]
    B, _, Him, Wim = img1.shape
    pyr1 = feature_pyramid(p, img1) + [img1]
    pyr2 = feature_pyramid(p, img2) + [img2]
    h0, w0 = pyr1[0].shape[2:]
    z = lambda c: torch.zeros(B, c, h0, w0, dtype=img1.dtype, device=img1.device)
    flow_f, flow_b, occ_f, occ_b = z(2), z(2), z(1), z(1)
    rec = (lambda k, v: record.__setitem__(k, v.clone())) if record is not None else (lambda k, v: None)
    occ_ctx = "occ_context_networks" if irr else "context_networks_occ"
    for l, (x1, x2) in enumerate(zip(pyr1, pyr2)):
        if l == 0:
            x2w, x1w = x2, x1
        else:
            flow_f = resize_ac(flow_f, x1)
            x2w = warp(x2, flow_f, Him, Wim, div_flow)
            if bi:
                flow_b = resize_ac(flow_b, x2)
                x1w = warp(x1, flow_b, Him, Wim, div_flow)
            if occ:
                occ_f = resize_ac(occ_f, x1)
                if bi:
                    occ_b = resize_ac(occ_b, x2)
        rec(f"l{l}.x1", x1); rec(f"l{l}.x2", x2)
        rec(f"l{l}.flow_up_f", flow_f); rec(f"l{l}.flow_up_b", flow_b)
        rec(f"l{l}.occ_up_f", occ_f); rec(f"l{l}.occ_up_b", occ_b)
        corr_f = F.leaky_relu(cost_volume(x1, x2w), LEAKY)
        corr_b = F.leaky_relu(cost_volume(x2, x1w), LEAKY) if bi else None
        rec(f"l{l}.corr_f", corr_f)
        if irr:
            hw = x1.shape[2:]
            su_l, sv_l = flow_scales(hw, Him, Wim, div_flow, True)
            su_g, sv_g = flow_scales(hw, Him, Wim, div_flow, False)
            x1_1 = conv_block(p, f"conv_1x1.{l}", x1)
            flow_f = scale_flow(flow_f, su_l, sv_l)
            xi, res = dense_estimator(p, "flow_estimators", torch.cat([corr_f, x1_1, flow_f], 1))
            flow_f = flow_f + res
            flow_f = flow_f + context_net(p, "context_networks", torch.cat([xi, flow_f], 1))
            flow_f = scale_flow(flow_f, su_g, sv_g)
            if bi:
                x2_1 = conv_block(p, f"conv_1x1.{l}", x2)
                flow_b = scale_flow(flow_b, su_l, sv_l)
                xi, res = dense_estimator(p, "flow_estimators", torch.cat([corr_b, x2_1, flow_b], 1))
                flow_b = flow_b + res
                flow_b = flow_b + context_net(p, "context_networks", torch.cat([xi, flow_b], 1))
                flow_b = scale_flow(flow_b, su_g, sv_g)
            if occ:  # pwcnet_irr_occ.py:92-97 (uni-directional only in this family: irr+occ+bi is PWCNet_irr_occ_bi)
                xo, ores = dense_estimator(p, "occ_estimators", torch.cat([corr_f, x1_1, occ_f], 1))
                occ_f = occ_f + ores
                occ_f = occ_f + context_net(p, occ_ctx, torch.cat([xo, occ_f], 1))
        else:
            in_f = corr_f if l == 0 else torch.cat([corr_f, x1, flow_f], 1)
            xi_f, flow_f = dense_estimator(p, f"flow_estimators.{l}", in_f)
            if bi:
                in_b = corr_b if l == 0 else torch.cat([corr_b, x2, flow_b], 1)
                xi_b, flow_b = dense_estimator(p, f"flow_estimators.{l}", in_b)
            if occ:
                io_f = corr_f if l == 0 else torch.cat([corr_f, x1, occ_f], 1)
                xo_f, occ_f = dense_estimator(p, f"occ_estimators.{l}", io_f)
                if bi:  # pwcnet_occ_bi.py:103 — x1, as written
                    io_b = corr_b if l == 0 else torch.cat([corr_b, x1, occ_b], 1)
                    xo_b, occ_b = dense_estimator(p, f"occ_estimators.{l}", io_b)
            if l == OUT_LEVEL:
                flow_f = flow_f + context_net(p, "context_networks", torch.cat([xi_f, flow_f], 1))
                if bi:
                    flow_b = flow_b + context_net(p, "context_networks", torch.cat([xi_b, flow_b], 1))
                if occ:
                    occ_f = occ_f + context_net(p, occ_ctx, torch.cat([xo_f, occ_f], 1))
                    if bi:
                        occ_b = occ_b + context_net(p, occ_ctx, torch.cat([xo_b, occ_b], 1))
        rec(f"l{l}.flow_f", flow_f)
        if bi:
            rec(f"l{l}.flow_b", flow_b)
        if occ:
            rec(f"l{l}.occ_f", occ_f)
            if bi:
                rec(f"l{l}.occ_b", occ_b)
        if l == OUT_LEVEL:
            break
    out = {"flow": resize_ac(flow_f, img1) * (1.0 / div_flow)}
    if occ:
        out["occ"] = resize_ac(occ_f, img1)
    return out


FORWARDS = {"IRR_PWC": irr_pwc_forward, "PWCNet": pwcnet_forward, "PWCNet_irr_occ_bi": pwcnet_irr_occ_bi_forward}
for _name in FAMILY:
    FORWARDS[_name] = (lambda name: (lambda p, a, b, div_flow=0.05, record=None:
                                     pwc_family_forward(name, p, a, b, div_flow, record)))(_name)


# --------------------------------------------------------------------------- parameters / synthetic data
# Data generators (no model arithmetic) live with the product so that bench.py's timed path never touches oracle/.
from irr_b200.synthetic import epe, param_shapes, synthetic_pair, synthetic_params  # noqa: E402,F401
