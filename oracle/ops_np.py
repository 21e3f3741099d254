"""numpy fp32 restatements of the op-level arithmetic on the hot path (TEST INFRASTRUCTURE ONLY).

These are the portable known-answer oracles for the C-ABI kernels:
  * cost_volume_np  — models/pwc_modules.py:42-62 (== correlation_cuda_kernel.cu:41-114 at k=1, s1=s2=1)
  * corr_ref_c      — ctypes binding of oracle/corr_ref.c (generic kernel_size / strides)
  * warp_np         — models/pwc_modules.py:107-133 incl. the rounding-sensitive validity mask; every
                      fp32 op is separately rounded, left to right, exactly as PyTorch's grid_sampler_2d does
                      (torch/include/ATen/native/cuda/GridSampler.cuh:22-31 unnormalize; the bilinear weights
                      nw=(x1-ix)(y1-iy) ... of aten/src/ATen/native/cuda/GridSampler.cu)
  * resize_ac_np    — bilinear align_corners=True (models/pwc_modules.py:65-67; ATen UpSample.cuh:96-130
                      area_pixel_compute_scale / source index)
Pinned against the reference's Python by oracle/gen_golden.py (fixtures in tests/golden/).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

f32 = np.float32


def cost_volume_np(f1: np.ndarray, f2: np.ndarray, max_disp: int = 4) -> np.ndarray:
    B, C, H, W = f1.shape
    d = max_disp
    f2p = np.zeros((B, C, H + 2 * d, W + 2 * d), f32)
    f2p[:, :, d:d + H, d:d + W] = f2
    out = np.empty((B, (2 * d + 1) ** 2, H, W), f32)
    t = 0
    for i in range(2 * d + 1):
        for j in range(2 * d + 1):
            prod = f1 * f2p[:, :, i:i + H, j:j + W]
            out[:, t] = prod.sum(axis=1, dtype=f32) / f32(C)
            t += 1
    return out


_LIB = None


def _corr_lib():
    global _LIB
    if _LIB is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "_build", "libcorr_ref.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", here, "-s"])
        _LIB = ctypes.CDLL(path)
        _LIB.corr_ref_forward.restype = ctypes.c_int
        _LIB.corr_ref_out_shape.restype = ctypes.c_int
    return _LIB


def corr_ref_c(in1: np.ndarray, in2: np.ndarray, pad=4, ksize=1, max_disp=4, s1=1, s2=1) -> np.ndarray:
    lib = _corr_lib()
    in1 = np.ascontiguousarray(in1, f32)
    in2 = np.ascontiguousarray(in2, f32)
    B, C, H, W = in1.shape
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    if lib.corr_ref_out_shape(H, W, pad, ksize, max_disp, s1, s2, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)):
        raise ValueError("bad correlation parameters")
    out = np.empty((B, oc.value, oh.value, ow.value), f32)
    fp = ctypes.POINTER(ctypes.c_float)
    rc = lib.corr_ref_forward(in1.ctypes.data_as(fp), in2.ctypes.data_as(fp), out.ctypes.data_as(fp),
                              B, C, H, W, pad, ksize, max_disp, s1, s2)
    if rc:
        raise RuntimeError("corr_ref_forward failed")
    return out


def grid_coords_np(flow, height_im, width_im, div_flow, lin_x, lin_y):
    """Unnormalised sample coordinates (ix, iy), every op rounded to fp32 in the reference's order."""
    B, _, H, W = flow.shape
    u = flow[:, 0].astype(f32)
    v = flow[:, 1].astype(f32)
    fx = ((u * f32(2)) / f32(max(width_im - 1, 1))) / f32(div_flow)  # pwc_modules.py:121
    fy = ((v * f32(2)) / f32(max(height_im - 1, 1))) / f32(div_flow)  # pwc_modules.py:122
    gx = lin_x.astype(f32)[None, None, :] + fx  # pwc_modules.py:126
    gy = lin_y.astype(f32)[None, :, None] + fy
    ix = ((gx + f32(1)) / f32(2)) * f32(W - 1)  # GridSampler.cuh:27-28 (align_corners)
    iy = ((gy + f32(1)) / f32(2)) * f32(H - 1)
    return ix.astype(f32), iy.astype(f32)


def warp_np(x, flow, height_im, width_im, div_flow, lin_x, lin_y):
    """Returns (warped*mask, mask) — bilinear, zeros padding, align_corners=True, mask = (sum_w >= 1)."""
    x = x.astype(f32)
    B, C, H, W = x.shape
    ix, iy = grid_coords_np(flow, height_im, width_im, div_flow, lin_x, lin_y)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1 = x0 + f32(1)
    y1 = y0 + f32(1)
    nw = (x1 - ix) * (y1 - iy)
    ne = (ix - x0) * (y1 - iy)
    sw = (x1 - ix) * (iy - y0)
    se = (ix - x0) * (iy - y0)

    def inb(xx, yy):
        return (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)

    # non-finite coordinates (inf/nan flow) sample nothing; clip before the int cast
    big = f32(2 ** 30)
    xi0 = np.clip(np.nan_to_num(x0, nan=-big), -big, big).astype(np.int64)
    yi0 = np.clip(np.nan_to_num(y0, nan=-big), -big, big).astype(np.int64)
    xi1, yi1 = xi0 + 1, yi0 + 1
    taps = [(xi0, yi0, nw), (xi1, yi0, ne), (xi0, yi1, sw), (xi1, yi1, se)]
    msum = np.zeros((B, H, W), f32)
    out = np.zeros((B, C, H, W), f32)
    bidx = np.arange(B)[:, None, None]
    for xi, yi, w in taps:
        ok = inb(xi, yi)
        wz = np.where(ok, w, f32(0)).astype(f32)
        msum = (msum + wz).astype(f32)
        xc = np.clip(xi, 0, W - 1)
        yc = np.clip(yi, 0, H - 1)
        vals = x[bidx, :, yc, xc]  # B,H,W,C
        out = (out + np.moveaxis(vals, -1, 1) * wz[:, None]).astype(f32)
    mask = (msum >= f32(1)).astype(f32)
    return out * mask[:, None], mask


def resize_ac_np(x: np.ndarray, oh: int, ow: int) -> np.ndarray:
    """Bilinear align_corners=True resize; CUDA-kernel operation order (UpSampleBilinear2d.cu):
    h1r = rheight*h2; h1 = int(h1r); h1lambda = h1r - h1; val = h0l*(w0l*a + w1l*b) + h1l*(w0l*c + w1l*d)."""
    x = x.astype(f32)
    B, C, H, W = x.shape
    rh = f32(0) if oh <= 1 else f32(f32(H - 1) / f32(oh - 1))
    rw = f32(0) if ow <= 1 else f32(f32(W - 1) / f32(ow - 1))
    h1r = (rh * np.arange(oh, dtype=f32)).astype(f32)
    w1r = (rw * np.arange(ow, dtype=f32)).astype(f32)
    h1 = h1r.astype(np.int64)
    w1 = w1r.astype(np.int64)
    h1p = (h1 < H - 1).astype(np.int64)
    w1p = (w1 < W - 1).astype(np.int64)
    h1l = (h1r - h1.astype(f32)).astype(f32)
    w1l = (w1r - w1.astype(f32)).astype(f32)
    h0l = (f32(1) - h1l).astype(f32)
    w0l = (f32(1) - w1l).astype(f32)
    a = x[:, :, h1][:, :, :, w1]
    b = x[:, :, h1][:, :, :, w1 + w1p]
    c = x[:, :, h1 + h1p][:, :, :, w1]
    d = x[:, :, h1 + h1p][:, :, :, w1 + w1p]
    top = (w0l * a + w1l * b).astype(f32)
    bot = (w0l * c + w1l * d).astype(f32)
    return (h0l[:, None] * top + h1l[:, None] * bot).astype(f32)


def cost_volume_backward_np(f1: np.ndarray, f2: np.ndarray, grad_out: np.ndarray, max_disp: int = 4):
    """Backward of cost_volume_np == what correlation_backward_input1 / _input2 return for the PWC parameters
    (correlation_cuda_kernel.cu:116-300; sumelems = C there, :195/:289):
        grad_f1[b,c,y,x] = (1/C) sum_d g[b,d,y,x] * f2[b,c,y+dy,x+dx]
        grad_f2[b,c,y,x] = (1/C) sum_d g[b,d,y-dy,x-dx] * f1[b,c,y-dy,x-dx]
    float64 accumulation (checker only)."""
    B, C, H, W = f1.shape
    md, nd = max_disp, 2 * max_disp + 1
    f1p = np.pad(f1.astype(np.float64), ((0, 0), (0, 0), (md, md), (md, md)))
    f2p = np.pad(f2.astype(np.float64), ((0, 0), (0, 0), (md, md), (md, md)))
    g = grad_out.astype(np.float64)
    g1 = np.zeros((B, C, H, W), np.float64)
    g2p = np.zeros((B, C, H + 2 * md, W + 2 * md), np.float64)
    for dyi in range(nd):
        for dxi in range(nd):
            gd = g[:, dyi * nd + dxi][:, None]                      # (B, 1, H, W)
            g1 += gd * f2p[:, :, dyi:dyi + H, dxi:dxi + W]
            g2p[:, :, dyi:dyi + H, dxi:dxi + W] += gd * f1p[:, :, md:md + H, md:md + W]
    return (g1 / C).astype(np.float32), (g2p[:, :, md:md + H, md:md + W] / C).astype(np.float32)
