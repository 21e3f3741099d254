"""Host-side mirror of the reference's ``models/pwc_modules.py`` for the inference path.

Same public names, constructor arguments, ``forward()`` signatures and parameter names (so the reference's
checkpoints load unchanged); the arithmetic is the hand-written sm_100a library behind
``irr_b200.ops``.  nn.Conv2d / nn.LeakyReLU objects are kept as *parameter holders* only — their library
forward is never executed.  There is no non-CUDA path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

_default_math = ops.MATH_TC_3XF16  # the product path; MATH_FP32_SIMT / MATH_TC_3XTF32 remain selectable


def set_conv_math(math: int) -> None:
    """Process-wide default arithmetic for conv blocks (ops.MATH_*); layers the tensor-core path cannot take
    fall back to the CUDA-core kernel *inside the native library's dispatch table*, never to PyTorch."""
    global _default_math
    _default_math = math
    ops.set_pitch_enabled(math == ops.MATH_TC_3XF16)   # only the 3xF16 conv path reads / writes row-pitched tensors


def get_conv_math() -> int:
    return _default_math


class ConvBlock(nn.Sequential):
    """``conv()`` of models/pwc_modules.py:8-19: [Conv2d(bias, pad=((k-1)*dil)//2), LeakyReLU(0.1)?].

    forward(x, out=None, addend=None, alpha=1.0) computes  addend + alpha * act(conv(x) + bias)  with our kernel,
    optionally straight into a channel slice ``out`` of a larger buffer."""

    def __init__(self, in_planes, out_planes, kernel_size=3, stride=1, dilation=1, isReLU=True):
        mods = [nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, dilation=dilation,
                          padding=((kernel_size - 1) * dilation) // 2, bias=True)]
        if isReLU:
            mods.append(nn.LeakyReLU(0.1, inplace=True))
        super().__init__(*mods)
        self.cin, self.cout, self.ks, self.stride, self.dil = in_planes, out_planes, kernel_size, stride, dilation
        self.slope = 0.1 if isReLU else 1.0
        self._packed = {}  # math -> (key, packed tensor)

    def _math(self):
        m = _default_math
        if m == ops.MATH_TC_3XF16 and ops.direct_supported(self.cout, self.cin, self.ks):
            return ops.MATH_FP32_SIMT   # thin HBM-bound layers (3 -> 16, 16 -> 3 1x1): the direct fp32 kernel, not a tile pipeline
        if m != ops.MATH_FP32_SIMT and not ops.tc_supported(self.cout, self.cin, self.ks, self.stride, self.dil, m):
            m = ops.MATH_FP32_SIMT
        return m

    def packed(self, math):
        w = self[0].weight
        key = (w.data_ptr(), w._version, str(w.device))
        hit = self._packed.get(math)
        if hit is None or hit[0] != key:
            hit = (key, ops.pack_weights(w, math))
            self._packed[math] = hit
        return hit[1]

    def forward(self, x, out=None, addend=None, alpha=1.0):
        math = self._math()
        w, b = self[0].weight, self[0].bias
        # Autograd: a gradient is wanted when the input carries one, or when the block is in train
        # mode with trainable parameters.  An .eval() model called outside torch.no_grad() — parameters still at
        # requires_grad=True — runs the inference kernels, as the stage-level entry points always have.
        if torch.is_grad_enabled() and (x.requires_grad or (self.training and (w.requires_grad or b.requires_grad))):
            if out is not None or addend is not None or alpha != 1.0:
                raise RuntimeError("irr_b200.conv: the fused out= / addend= / alpha forms are inference-only")
            return _ConvFunction.apply(x, w, b, self, math)
        return ops.conv2d(x, self.packed(math), b, self.cout, self.ks, self.stride, self.dil,
                          slope=self.slope, out=out, addend=addend, alpha=alpha, math=math)


class _ConvFunction(torch.autograd.Function):
    """conv() under autograd: the forward is the irr_b200 kernel; the backward (not on the inference
    hot path, no kernel of ours) goes through ATen's convolution_backward on the LeakyReLU-masked gradient — the sign of
    the output equals the sign of the pre-activation for any positive slope, so nothing but y has to be kept."""

    @staticmethod
    def forward(ctx, x, w, b, block, math):
        x = x.contiguous()
        y = ops.conv2d(x, block.packed(math), b, block.cout, block.ks, block.stride, block.dil, slope=block.slope, math=math)
        ctx.save_for_backward(x, w, y)
        ctx.block = block
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        blk = ctx.block
        g = gy if blk.slope == 1.0 else gy * torch.where(y > 0, 1.0, blk.slope).to(gy.dtype)
        pad = ((blk.ks - 1) * blk.dil) // 2
        gx, gw, gb = torch.ops.aten.convolution_backward(
            g.contiguous(), x, w, [blk.cout], [blk.stride] * 2, [pad] * 2, [blk.dil] * 2, False, [0, 0], 1,
            [ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]])
        return gx, gw, gb, None, None


def conv(in_planes, out_planes, kernel_size=3, stride=1, dilation=1, isReLU=True):
    return ConvBlock(in_planes, out_planes, kernel_size, stride, dilation, isReLU)


def initialize_msra(modules):
    """models/pwc_modules.py:22-39."""
    for layer in modules:
        if isinstance(layer, (nn.Conv2d, nn.ConvTranspose2d)):
            nn.init.kaiming_normal_(layer.weight)
            if layer.bias is not None:
                nn.init.constant_(layer.bias, 0)


def compute_cost_volume(feat1, feat2, param_dict):
    """models/pwc_modules.py:42-62 (kernel_size=1, stride1=stride2=1); only ``max_disp`` is read, as there."""
    return ops.correlation(feat1, feat2, max_disp=param_dict["max_disp"])


def upsample2d_as(inputs, target_as, mode="bilinear"):
    """models/pwc_modules.py:65-67."""
    assert mode == "bilinear"
    _, _, h, w = target_as.size()
    return ops.resize_ac(inputs, h, w, pitched=False)   # the reference-API function returns dense rows


def flow_scales(h, w, div_flow, width_im, height_im, to_local=True):
    """The Python-double scale factors of rescale_flow (models/pwc_modules.py:71-76)."""
    if to_local:
        return float(w / width_im / div_flow), float(h / height_im / div_flow)
    return float(width_im * div_flow / w), float(height_im * div_flow / h)


class _ScaleChannels(torch.autograd.Function):
    """y[:, c] = x[:, c] * (s_even if c even else s_odd) on the irr_b200 kernel, differentiable (the backward is the
    same scaling of the incoming gradient)."""

    @staticmethod
    def forward(ctx, x, s_even, s_odd):
        ctx.scales = (s_even, s_odd)
        return ops.scale_channels(x.contiguous(), s_even=s_even, s_odd=s_odd)

    @staticmethod
    def backward(ctx, g):
        return ops.scale_channels(g.contiguous(), s_even=ctx.scales[0], s_odd=ctx.scales[1]), None, None


def rescale_flow(flow, div_flow, width_im, height_im, to_local=True):
    """models/pwc_modules.py:70-82.

    Without autograd (inference) it keeps the reference's in-place side effect on ``flow``: the argument
    is scaled in place AND a new tensor with the same values is returned — IRR_PWC's eval forward depends on it
    (IRR_PWC.py:128-129).  When a gradient is required the reference's in-place ``u *= u_scale`` on a ``chunk`` view is
    exactly what torch >= 2 refuses to differentiate; there the AUTOGRAD-SAFE form runs: the argument is left
    untouched and the scaled flow is returned as a new differentiable tensor."""
    su, sv = flow_scales(flow.size(2), flow.size(3), div_flow, width_im, height_im, to_local)
    if torch.is_grad_enabled() and flow.requires_grad:
        return _ScaleChannels.apply(flow, su, sv)
    ops.scale_channels(flow, out=flow, s_even=su, s_odd=sv)
    return flow.clone()


class FeatureExtractor(nn.Module):
    """models/pwc_modules.py:85-104."""

    def __init__(self, num_chs):
        super().__init__()
        self.num_chs = num_chs
        self.convs = nn.ModuleList()
        for ch_in, ch_out in zip(num_chs[:-1], num_chs[1:]):
            self.convs.append(nn.Sequential(conv(ch_in, ch_out, stride=2), conv(ch_out, ch_out)))

    def forward(self, x):
        pyramid = []
        for pair in self.convs:
            x = pair[1](pair[0](x))
            pyramid.append(x)
        return pyramid[::-1]


def get_grid(x):
    """models/pwc_modules.py:107-112 — kept for API parity (the kernels take the two 1-D vectors instead)."""
    B, _, H, W = x.shape
    gx = ops.host_linspace(W, x.device).view(1, 1, 1, W).expand(B, 1, H, W)
    gy = ops.host_linspace(H, x.device).view(1, 1, H, 1).expand(B, 1, H, W)
    return torch.cat([gx, gy], 1)


class WarpFunction(torch.autograd.Function):
    """WarpingLayer.forward with a hand-written backward (``irr_warp_bwd``): gradients w.r.t. the warped tensor and the
    flow, the hard validity mask being a constant exactly as under autograd in the reference (pwc_modules.py:129-133)."""

    @staticmethod
    def forward(ctx, x, flow, height_im, width_im, div_flow):
        x, flow = x.contiguous(), flow.contiguous()
        ctx.save_for_backward(x, flow)
        ctx.geom = (height_im, width_im, div_flow)
        return ops.warp(x, flow, height_im, width_im, div_flow)

    @staticmethod
    def backward(ctx, grad_out):
        x, flow = ctx.saved_tensors
        gx, gf = ops.warp_backward(x, flow, grad_out.contiguous(), *ctx.geom, need_x=ctx.needs_input_grad[0],
                                   need_flow=ctx.needs_input_grad[1])
        return gx, gf, None, None, None


class WarpingLayer(nn.Module):
    """models/pwc_modules.py:115-133 (differentiable: ``WarpFunction`` when a gradient is required)."""

    def forward(self, x, flow, height_im, width_im, div_flow):
        if torch.is_grad_enabled() and (x.requires_grad or flow.requires_grad):
            return WarpFunction.apply(x, flow, height_im, width_im, div_flow)
        return ops.warp(x, flow, height_im, width_im, div_flow)


class _DenseEstimator(nn.Module):
    """FlowEstimatorDense / OccEstimatorDense (models/pwc_modules.py:153-170,190-207).

    The reference re-concatenates after every conv (x_{i} = cat[conv_i(x_{i-1}), x_{i-1}]).  Here the whole block lives
    in ONE buffer laid out as the final x5 = [conv5 | conv4 | conv3 | conv2 | conv1 | x]; conv_i reads the channel
    suffix that already exists and writes its slice in front of it, so no copy is ever made."""

    GROWTH = [128, 128, 96, 64, 32]

    def __init__(self, ch_in, ch_out):
        super().__init__()
        self.ch_in, self.ch_out = ch_in, ch_out
        self.conv1 = conv(ch_in, 128)
        self.conv2 = conv(ch_in + 128, 128)
        self.conv3 = conv(ch_in + 256, 96)
        self.conv4 = conv(ch_in + 352, 64)
        self.conv5 = conv(ch_in + 416, 32)
        self.conv_last = conv(ch_in + 448, ch_out, isReLU=False)

    @property
    def total_ch(self):
        return self.ch_in + 448

    def forward_into(self, buf, out=None, addend=None):
        """``buf``: (B, >= ch_in+448, H, W) whose channels [448 : 448+ch_in] already hold the block input.
        Fills channels [0:448]; returns conv_last(buf[:, :ch_in+448]) (+ addend) in ``out``."""
        hi = 448
        fuse = get_conv_math() == ops.MATH_TC_3XF16
        for c, g in zip([self.conv1, self.conv2, self.conv3, self.conv4, self.conv5], self.GROWTH):
            if fuse and c is self.conv4:
                break
            c(buf[:, hi:self.total_ch], out=buf[:, hi - g:hi])
            hi -= g
        if not fuse:
            return self.conv_last(buf[:, 0:self.total_ch], out=out, addend=addend)
        # Fused tail.  conv5 and conv_last read everything conv4 reads (plus conv4's / conv5's own outputs), and thin
        # layers cost the same producer time per input channel as fat ones, so the part of conv5 and of conv_last that
        # sees conv4's input rides along with conv4 as extra output columns (raw partial sums), one pass over the
        # 467/466 channels instead of three:
        #   pass B  x = [c3|c2|c1|in]    : conv4 (64, final) | conv5 partial (32) | conv_last partial (+ b_last + addend)
        #   pass C  x = c4 (64 ch)        : conv5 = lrelu(. + partial + b5) (32, final) | conv_last partial += .
        #   pass D  x = c5 (32 ch)        : conv_last = partial + .
        # Same sums, different association (<= 1e-6).
        (pB, bB), (pC, bC), (pD, bD) = self._fused_tail()
        B, _, H, W = buf.shape
        co = self.ch_out
        dev = buf.device
        if out is None:
            out = ops.new_like(buf, co)
        p5 = ops.new_like(buf, 32)
        pl = ops.new_like(buf, co)
        ops.conv2d_multi(buf[:, 96:self.total_ch], pB, bB, 96 + co, 3, [
            dict(n_begin=0, out=buf[:, 32:96], slope=0.1),
            dict(n_begin=64, out=p5, slope=1.0),
            dict(n_begin=96, out=pl, slope=1.0, addend=addend)])
        ops.conv2d_multi(buf[:, 32:96], pC, bC, 32 + co, 3, [
            dict(n_begin=0, out=buf[:, 0:32], slope=0.1, addend=p5, pre=True),
            dict(n_begin=32, out=pl, slope=1.0, addend=pl)])
        return ops.conv2d(buf[:, 0:32], pD, bD, co, 3, slope=1.0, out=out, addend=pl, math=ops.MATH_TC_3XF16)

    def _fused_tail(self):
        """Packed weights / biases of the three fused-tail passes (cached; rebuilt when a parameter changes)."""
        w4, b4 = self.conv4[0].weight, self.conv4[0].bias
        w5, b5 = self.conv5[0].weight, self.conv5[0].bias
        wl, bl = self.conv_last[0].weight, self.conv_last[0].bias
        key = tuple((t.data_ptr(), t._version, str(t.device)) for t in (w4, b4, w5, b5, wl, bl))
        hit = getattr(self, "_fused_cache", None)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                # channel order of the buffer: [c5 32 | c4 64 | c3 96 | c2 128 | c1 128 | in]; conv4 reads [96:], conv5
                # reads [32:] (its first 64 input channels are c4), conv_last reads [0:] (c5, then c4, then conv4's input)
                wB = torch.cat([w4, w5[:, 64:], wl[:, 96:]], 0).contiguous()
                bB = torch.cat([b4, torch.zeros_like(b5), bl], 0).contiguous()
                wC = torch.cat([w5[:, :64], wl[:, 32:96]], 0).contiguous()
                bC = torch.cat([b5, torch.zeros_like(bl)], 0).contiguous()
                wD = wl[:, :32].contiguous()
                bD = torch.zeros_like(bl)
                M = ops.MATH_TC_3XF16
                hit = (key, ((ops.pack_weights(wB, M), bB), (ops.pack_weights(wC, M), bC), (ops.pack_weights(wD, M), bD)))
            self._fused_cache = hit
        return hit[1]

    def forward(self, x):
        B, C, H, W = x.shape
        buf = ops.empty(B, self.total_ch, H, W, x.device)
        ops.scale_channels(x, out=buf[:, 448:])
        x_out = self.forward_into(buf)
        return buf, x_out


class FlowEstimatorDense(_DenseEstimator):
    def __init__(self, ch_in):
        super().__init__(ch_in, 2)


class OccEstimatorDense(_DenseEstimator):
    def __init__(self, ch_in):
        super().__init__(ch_in, 1)


class _Context(nn.Module):
    """ContextNetwork / OccContextNetwork (models/pwc_modules.py:210-243): dilations 1,2,4,8,16,1,1."""

    def __init__(self, ch_in, ch_out):
        super().__init__()
        self.convs = nn.Sequential(
            conv(ch_in, 128, 3, 1, 1), conv(128, 128, 3, 1, 2), conv(128, 128, 3, 1, 4), conv(128, 96, 3, 1, 8),
            conv(96, 64, 3, 1, 16), conv(64, 32, 3, 1, 1), conv(32, ch_out, isReLU=False))

    def forward(self, x, out=None, addend=None):
        for c in list(self.convs)[:-1]:
            x = c(x)
        return self.convs[-1](x, out=out, addend=addend)


class ContextNetwork(_Context):
    def __init__(self, ch_in):
        super().__init__(ch_in, 2)


class OccContextNetwork(_Context):
    def __init__(self, ch_in):
        super().__init__(ch_in, 1)
