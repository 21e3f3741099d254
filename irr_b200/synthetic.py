"""Deterministic synthetic weights and image pairs — data generators only, no model arithmetic.

Shared by the bench, the tests, the oracle and the profiling scripts so that the build container, the GPU box, the
golden fixtures and every arm of ``bench.py`` see bit-identical parameters and inputs.  Parameter names / OIHW shapes
follow the reference's ``state_dict``; values come from numpy's frozen ``RandomState`` stream.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

PYR_CHS = [3, 16, 32, 64, 96, 128, 196]  # models/IRR_PWC.py:20
SEARCH = 4  # models/IRR_PWC.py:19
OUT_LEVEL = 4  # models/IRR_PWC.py:21


def param_shapes(model: str) -> Dict[str, tuple]:
    """Parameter names and OIHW shapes of the three benchmarked classes."""
    s: Dict[str, tuple] = {}

    def add(name, cin, cout, k=3):
        s[name + ".0.weight"] = (cout, cin, k, k)
        s[name + ".0.bias"] = (cout,)

    for l in range(6):
        add(f"feature_pyramid_extractor.convs.{l}.0", PYR_CHS[l], PYR_CHS[l + 1])
        add(f"feature_pyramid_extractor.convs.{l}.1", PYR_CHS[l + 1], PYR_CHS[l + 1])

    def dense(prefix, cin, cout):
        for i, (extra, co) in enumerate([(0, 128), (128, 128), (256, 96), (352, 64), (416, 32)], start=1):
            add(f"{prefix}.conv{i}", cin + extra, co)
        add(f"{prefix}.conv_last", cin + 448, cout)

    def ctx(prefix, cin, cout):
        chs = [cin, 128, 128, 128, 96, 64, 32, cout]
        for i in range(7):
            add(f"{prefix}.convs.{i}", chs[i], chs[i + 1])

    def refine(prefix, cin):
        chs = [cin, 128, 128, 64, 64, 32, 32, 9]
        for i in range(7):
            add(f"{prefix}.convs.{i}", chs[i], chs[i + 1])

    dim_corr = (2 * SEARCH + 1) ** 2
    if model in ("PWCNet", "PWCNet_bi", "PWCNet_occ", "PWCNet_occ_bi"):  # per-level estimators (pwcnet.py:23-37 ...)
        for l, ch in enumerate(PYR_CHS[::-1][:OUT_LEVEL + 1]):
            dense(f"flow_estimators.{l}", dim_corr if l == 0 else dim_corr + ch + 2, 2)
            if "occ" in model:
                dense(f"occ_estimators.{l}", dim_corr if l == 0 else dim_corr + ch + 1, 1)
        ctx("context_networks", dim_corr + 32 + 2 + 448 + 2, 2)
        if "occ" in model:
            ctx("context_networks_occ", dim_corr + 32 + 1 + 448 + 1, 1)
        return s
    if model in ("PWCNet_irr", "PWCNet_irr_bi", "PWCNet_irr_occ"):  # shared estimators, five 1x1 convs
        dense("flow_estimators", dim_corr + 34, 2)
        ctx("context_networks", dim_corr + 34 + 448 + 2, 2)
        if model == "PWCNet_irr_occ":
            dense("occ_estimators", dim_corr + 33, 1)
            ctx("occ_context_networks", dim_corr + 33 + 448 + 1, 1)
        for l, c in enumerate([196, 128, 96, 64, 32]):
            add(f"conv_1x1.{l}", c, 32, 1)
        return s
    dense("flow_estimators", dim_corr + 34, 2)
    ctx("context_networks", dim_corr + 34 + 448 + 2, 2)
    dense("occ_estimators", dim_corr + 33, 1)
    ctx("occ_context_networks", dim_corr + 33 + 448 + 1, 1)
    chs_1x1 = [196, 128, 96, 64] + ([32] if model == "PWCNet_irr_occ_bi" else [])
    for l, c in enumerate(chs_1x1):
        add(f"conv_1x1.{l}", c, 32, 1)
    if model == "IRR_PWC":
        add("occ_shuffle_upsample.init_conv", 11, 32)
        add("occ_shuffle_upsample.res_convs.0", 32, 32)
        add("occ_shuffle_upsample.res_convs.1", 32, 32)
        add("occ_shuffle_upsample.res_end_conv", 32, 32)
        add("occ_shuffle_upsample.out_convs", 32, 1)
        add("conv_1x1_1", 16, 3, 1)
        refine("refine_flow", 35)
        refine("refine_occ", 65)
    return s


def synthetic_params(model: str, seed: int = 1234, gain: float = 1.0) -> Params:
    """Deterministic MSRA-like weights (cf. initialize_msra, pwc_modules.py:22-39) from numpy's frozen
    RandomState stream, so the container, the GPU box, fixtures and tests all agree bit for bit.
    Biases get a small non-zero value so the bias path is exercised."""
    import zlib
    import numpy as np
    out: Params = {}
    for name, shape in param_shapes(model).items():
        rs = np.random.RandomState((zlib.crc32(name.encode()) + seed) & 0x7FFFFFFF)
        if len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            a = rs.standard_normal(shape).astype("float32") * np.float32(gain * math.sqrt(2.0 / fan_in))
        else:
            a = (rs.standard_normal(shape) * 0.01).astype("float32")
        out[name] = torch.from_numpy(a)
    return out


def synthetic_pair(B: int, H: int, W: int, seed: int = 3, max_flow: float = 12.0):
    """Smooth image pair with a known flow (SURVEY.md §8(d) recipe): img1 = bicubic-upsampled low-res noise;
    img2 is img1 displaced by a smooth low-frequency field (so img2(x) = img1(x - u) to first order).
    Returns (img1, img2, flow_gt) as fp32 CPU tensors."""
    import numpy as np
    rs = np.random.RandomState(seed)
    lo = torch.from_numpy(rs.uniform(0, 1, (B, 3, max(H // 16, 2), max(W // 16, 2))).astype("float32"))
    img1 = F.interpolate(lo, size=[H, W], mode="bicubic", align_corners=True).clamp(0, 1)
    fl = torch.from_numpy(rs.uniform(-1, 1, (B, 2, 3, 4)).astype("float32")) * max_flow
    flow = F.interpolate(fl, size=[H, W], mode="bicubic", align_corners=True)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    gx = (xs[None] - flow[:, 0]) * 2 / max(W - 1, 1) - 1
    gy = (ys[None] - flow[:, 1]) * 2 / max(H - 1, 1) - 1
    img2 = F.grid_sample(img1, torch.stack([gx, gy], -1), mode="bilinear", padding_mode="border", align_corners=True)
    return img1.contiguous(), img2.contiguous(), flow.contiguous()


def epe(a, b):
    """losses.py:8-10 — per-pixel end-point error, averaged."""
    return torch.norm(a - b, p=2, dim=1).mean()
