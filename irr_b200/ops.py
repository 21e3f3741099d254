"""Tensor-level wrappers over the C ABI (include/irr_b200.h).

Every function takes fp32 CUDA tensors that are NCHW channel-slices (``buf[:, a:b]`` of a contiguous buffer is
fine: stride(1)=H*W, stride(2)=W, stride(3)=1) and launches on torch's current stream.  ``out=`` lets callers write
straight into a channel slice of a pre-allocated concat buffer — that is how every torch.cat of the reference
disappears.  No function here has a non-CUDA path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

MATH_FP32_SIMT = 0
MATH_TC_3XTF32 = 1
MATH_TC_TF32 = 2
MATH_TC_3XF16 = 3

GRID_TRUE_DIV = 0   # reference-on-CPU arithmetic (IEEE divisions)
GRID_RECIP_MUL = 1  # torch-CUDA `tensor / python_scalar` arithmetic (a * (1/b))

_grid_mode = GRID_TRUE_DIV


def set_grid_mode(mode: int) -> None:
    """Select which reference arithmetic the warp grid reproduces bit-for-bit (see include/irr_b200.h)."""
    global _grid_mode
    assert mode in (GRID_TRUE_DIV, GRID_RECIP_MUL)
    _grid_mode = mode


def get_grid_mode() -> int:
    return _grid_mode


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# --- instrumentation used by bench.py: a launch counter, and optional per-launch CUDA-event timing on the launching
# stream (TIMING = [] to enable; entries are (name, meta, start_event, end_event)).
LAUNCHES = 0
TIMING = None


def _launch(what: str, meta, fn, *args) -> None:
    global LAUNCHES
    t = TIMING
    if t is not None:
        s = torch.cuda.Event(enable_timing=True)
        s.record()
    rc = fn(*args)
    if t is not None:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        t.append((what, meta, s, e))
    LAUNCHES += 1
    _lib.check(rc, what)


def _vp(t: torch.Tensor, name: str = "tensor", dtype=torch.float32) -> Tuple[int, int, int]:
    """(data_ptr, batch_stride, row_pitch) in elements of an NCHW channel-slice view whose rows may be stored with a
    pitch P >= W (``buf[:, a:b, :, :W]`` of a ``(B, C, H, P)`` buffer: stride(2) = P, stride(1) = H*P) — include/irr_b200.h,
    "ROW PITCH"."""
    if not (t.is_cuda and t.dtype == dtype and t.dim() == 4):
        raise RuntimeError(f"irr_b200: {name} must be a 4-D {dtype} CUDA tensor (got {t.dtype}, {t.device}, dim {t.dim()})")
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"irr_b200: {name} lives on {t.device} but cuda:{torch.cuda.current_device()} is the current device — "
                           f"wrap the call in `with torch.cuda.device_of(tensor):` (the kernels launch on the current device)")
    B, Cc, H, W = t.shape
    s = t.stride()
    # a single row of a single channel has no pitch: 0 = "any" (callers merge it with the other operands' pitch)
    P = s[2] if H > 1 else (s[1] if Cc > 1 else 0)
    ok = (W == 1 or s[3] == 1) and (P == 0 or P >= W) and (Cc == 1 or s[1] == H * P)
    if not ok:
        raise RuntimeError(f"irr_b200: {name} is not an NCHW channel-slice view (shape {tuple(t.shape)}, strides {s})")
    bs = s[0] if B > 1 else Cc * H * (P or W)
    return t.data_ptr(), bs, P


def _v(t: torch.Tensor, name: str = "tensor") -> Tuple[int, int]:
    """(data_ptr, batch_stride_in_elements) of a DENSE NCHW channel-slice view (row pitch == W)."""
    ptr, bs, P = _vp(t, name)
    if P and P != t.shape[3]:
        raise RuntimeError(f"irr_b200: {name} must have dense rows here (width {t.shape[3]}, row pitch {P})")
    return ptr, bs


def _pitch(name: str, *ps: int) -> int:
    """The common row pitch of the same-sized tensors of one call (0 = no operand has one: dense)."""
    nz = [p for p in ps if p]
    if any(p != nz[0] for p in nz):
        raise RuntimeError(f"irr_b200.{name}: tensors of one spatial size must share the row pitch (got {ps})")
    return nz[0] if nz else 0


def _is_pitched(t: torch.Tensor) -> Optional[bool]:
    """True / False: ``t`` has padded / dense rows; None: it has no pitch (one row of one channel) -> module default."""
    if t.dtype != torch.float32:
        return None
    P = _vp(t)[2]
    return (P != t.shape[3]) if P else None


# Row pitch of the buffers this module and the model classes allocate: widths that are not a multiple of 4 (KITTI's
# 621 / 311 / 78 / 39 ...) are stored with the next multiple of 4 as row pitch, which puts them on the TMA paths of the
# conv and correlation kernels.  Only the 3xF16 conv path understands pitched tensors, so pwc_modules.set_conv_math()
# switches this off for the other math modes.
PITCH_ALIGN = 4
_pitch_enabled = True


def set_pitch_enabled(flag: bool) -> None:
    global _pitch_enabled
    _pitch_enabled = bool(flag)


def empty(B: int, C: int, H: int, W: int, device, pitched: Optional[bool] = None, zero: bool = False) -> torch.Tensor:
    """A (B, C, H, W) fp32 tensor; when ``W`` is not a multiple of PITCH_ALIGN (and pitching is on) it is the ``[..., :W]``
    view of a buffer whose rows are padded to the next multiple (the pad columns are never read)."""
    use = _pitch_enabled if pitched is None else pitched
    Pw = (W + PITCH_ALIGN - 1) // PITCH_ALIGN * PITCH_ALIGN if use else W
    f = torch.zeros if zero else torch.empty
    t = f((B, C, H, Pw), dtype=torch.float32, device=device)
    return t if Pw == W else t[:, :, :, :W]


def new_like(ref: torch.Tensor, C: int, zero: bool = False) -> torch.Tensor:
    """A (B, C, H, W) tensor with the batch, spatial size, device and row pitch of ``ref``."""
    B, _, H, W = ref.shape
    return empty(B, C, H, W, ref.device, pitched=_is_pitched(ref), zero=zero)


def pitched(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """``t`` in the row pitch this module allocates with: unchanged when it already has it (always the case inside a
    model's forward), else a pitched copy (stage-level entry points fed with dense tensors)."""
    if t is None:
        return None
    B, C, H, W = t.shape
    want = (W + PITCH_ALIGN - 1) // PITCH_ALIGN * PITCH_ALIGN if _pitch_enabled else W
    if t.is_cuda and t.dtype == torch.float32 and t.dim() == 4:
        s = t.stride()
        if (W == 1 or s[3] == 1) and (H == 1 or s[2] == want) and (C == 1 or s[1] == H * want):
            return t
    return scale_channels(t.contiguous().float(), out=empty(B, C, H, W, t.device))


def stack_pair(x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """torch.cat([x1, x2], 0) into a (possibly row-pitched) fp32 buffer — the only data movement done by torch."""
    B, C, H, W = x1.shape
    t = empty(2 * B, C, H, W, x1.device)
    t[:B].copy_(x1)
    t[B:].copy_(x2)
    return t


def flat_hw(t: torch.Tensor) -> int:
    """H * row_pitch: the per-channel element count the flat kernels (scale / round / l2norm) iterate over."""
    return t.shape[2] * (_vp(t)[2] or t.shape[3])


def _p(t: torch.Tensor, name: str, like: Optional[torch.Tensor] = None, numel: Optional[int] = None) -> int:
    """data_ptr of a dense fp32 CUDA tensor (weights, bias, linspace vectors, masks, packed images) after checking what
    the kernels assume: CUDA, float32, contiguous, on the same device as ``like``, at least ``numel`` elements."""
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError(f"irr_b200: {name} must be a contiguous fp32 CUDA tensor "
                           f"(got {getattr(t, 'dtype', type(t))}, {getattr(t, 'device', '?')})")
    if like is not None and t.device != like.device:
        raise RuntimeError(f"irr_b200: {name} is on {t.device}, expected {like.device}")
    if numel is not None and t.numel() < numel:
        raise RuntimeError(f"irr_b200: {name} has {t.numel()} elements, needs {numel}")
    return t.data_ptr()


_lin_cache = {}


def host_linspace(n: int, device) -> torch.Tensor:
    """The base-grid vector exactly as the reference makes it: CPU torch.linspace(-1, 1, n), uploaded
    (models/pwc_modules.py:108-111) — cached per (n, device) instead of rebuilt on every call."""
    key = (n, str(device))
    t = _lin_cache.get(key)
    if t is None:
        t = torch.linspace(-1.0, 1.0, n).to(device)
        _lin_cache[key] = t
    return t


def _new(like: torch.Tensor, C: int, H: int, W: int, pitched: Optional[bool] = None) -> torch.Tensor:
    return empty(like.shape[0], C, H, W, like.device, pitched=pitched)


_corr_ws_bytes = {}


def _corr_workspace(B, C, H, W, device, fused=False):
    """Scratch for the channel-split plan of coarse-level cost volumes and, for fused launches, the per-pixel tap table
    of the pre-pass (0 bytes = neither applies to this shape)."""
    key = (B, C, H, W, bool(fused))
    n = _corr_ws_bytes.get(key)
    if n is None:
        n = _lib.load().irr_correlation_workspace_bytes(B, C, H, W, 1 if fused else 0)
        _corr_ws_bytes[key] = n
    return (torch.empty(n // 4, dtype=torch.float32, device=device), n) if n else (None, 0)


def correlation(f1, f2, out=None, shift: int = 0, slope: float = 1.0, max_disp: int = 4):
    B, C, H, W = f1.shape
    assert f2.shape == f1.shape
    D = (2 * max_disp + 1) ** 2
    if out is None:
        out = _new(f1, D, H, W, pitched=_is_pitched(f1))
    assert out.shape == (B, D, H, W)
    ws, nws = _corr_workspace(B, C, H, W, f1.device)
    if f1.dtype == torch.bfloat16:   # bf16 STORAGE for f1 / f2 (include/irr_b200.h irr_warp_correlation_fwd_dt)
        p1, s1, q1 = _vp(f1, "f1", torch.bfloat16); p2, s2, q2 = _vp(f2, "f2", torch.bfloat16); po, so, qo = _vp(out, "out")
        _launch("correlation", (B, C, H, W), _lib.load().irr_warp_correlation_fwd_dt, p1, s1, p2, s2, DTYPE_BF16,
                _pitch("correlation", q1, q2), None, 0, None, None, po, so, B, C, H, W, H, W, 1.0, max_disp, shift, slope, 0,
                ws.data_ptr() if ws is not None else None, nws, qo, _stream())
        return out
    p1, s1, q1 = _vp(f1, "f1"); p2, s2, q2 = _vp(f2, "f2"); po, so, qo = _vp(out, "out")
    P = _pitch("correlation", q1, q2, qo)
    _launch("correlation", (B, C, H, W), _lib.load().irr_warp_correlation_fwd_ws, p1, s1, p2, s2, None, 0, None, None,
            po, so, B, C, H, W, H, W, 1.0, max_disp, shift, slope, 0, ws.data_ptr() if ws is not None else None, nws, P,
            _stream())
    return out


def warp_correlation(f1, f2, flow, height_im: int, width_im: int, div_flow: float, out=None, shift: int = 0,
                     slope: float = 1.0, max_disp: int = 4, lin_x=None, lin_y=None):
    B, C, H, W = f1.shape
    assert f2.shape == f1.shape and flow.shape == (B, 2, H, W)
    D = (2 * max_disp + 1) ** 2
    if out is None:
        out = _new(f1, D, H, W, pitched=_is_pitched(f1))
    lx = host_linspace(W, f1.device) if lin_x is None else lin_x
    ly = host_linspace(H, f1.device) if lin_y is None else lin_y
    ws, nws = _corr_workspace(B, C, H, W, f1.device, fused=True)
    if f1.dtype == torch.bfloat16:
        p1, s1, q1 = _vp(f1, "f1", torch.bfloat16); p2, s2, q2 = _vp(f2, "f2", torch.bfloat16)
        pf, sf, qf = _vp(flow, "flow"); po, so, qo = _vp(out, "out")
        _launch("warp_correlation", (B, C, H, W), _lib.load().irr_warp_correlation_fwd_dt, p1, s1, p2, s2, DTYPE_BF16,
                _pitch("warp_correlation", q1, q2), pf, sf, _p(lx, "lin_x", flow, W), _p(ly, "lin_y", flow, H), po, so, B, C, H, W,
                height_im, width_im, div_flow, max_disp, shift, slope, _grid_mode, ws.data_ptr() if ws is not None else None,
                nws, _pitch("warp_correlation", qf, qo), _stream())
        return out
    p1, s1, q1 = _vp(f1, "f1"); p2, s2, q2 = _vp(f2, "f2"); pf, sf, qf = _vp(flow, "flow"); po, so, qo = _vp(out, "out")
    P = _pitch("warp_correlation", q1, q2, qf, qo)
    _launch("warp_correlation", (B, C, H, W), _lib.load().irr_warp_correlation_fwd_ws, p1, s1, p2, s2, pf, sf,
            _p(lx, "lin_x", f1, W), _p(ly, "lin_y", f1, H), po, so, B, C, H, W, height_im, width_im, div_flow, max_disp,
            shift, slope, _grid_mode, ws.data_ptr() if ws is not None else None, nws, P, _stream())
    return out


def warp(x, flow, height_im: int, width_im: int, div_flow: float, out=None, minuend=None, shift: int = 0,
         mask_out=None, lin_x=None, lin_y=None):
    B, C, H, W = x.shape
    assert flow.shape == (B, 2, H, W)
    if out is None:
        out = _new(x, C, H, W, pitched=_is_pitched(x))
    lx = host_linspace(W, x.device) if lin_x is None else lin_x
    ly = host_linspace(H, x.device) if lin_y is None else lin_y
    px, sx, qx = _vp(x, "x"); pf, sf, qf = _vp(flow, "flow"); po, so, qo = _vp(out, "out")
    pm, sm, qm = (_vp(minuend, "minuend") if minuend is not None else (None, 0, qx))
    P = _pitch("warp", qx, qf, qo, qm)
    _launch("warp", (B, C, H, W), _lib.load().irr_warp_fwd, px, sx, pf, sf, _p(lx, "lin_x", x, W), _p(ly, "lin_y", x, H),
            pm, sm, po, so, _p(mask_out, "mask_out", x, B * H * W) if mask_out is not None else None, B, C, H, W,
            height_im, width_im, div_flow, shift, _grid_mode, P, _stream())
    return out


def warp_backward(x, flow, grad_out, height_im: int, width_im: int, div_flow: float, need_x: bool = True,
                  need_flow: bool = True, lin_x=None, lin_y=None):
    """(grad_x, grad_flow) of ``warp(x, flow)`` (no batch rotation, no minuend) — include/irr_b200.h irr_warp_bwd.  The
    hard mask is a constant, as under autograd in the reference.  A gradient that is not needed is None."""
    B, C, H, W = x.shape
    assert flow.shape == (B, 2, H, W) and tuple(grad_out.shape) == (B, C, H, W)
    gx = torch.zeros_like(x) if need_x else None
    gf = torch.zeros((B, 2, H, W), dtype=torch.float32, device=x.device) if need_flow else None
    lx = host_linspace(W, x.device) if lin_x is None else lin_x
    ly = host_linspace(H, x.device) if lin_y is None else lin_y
    px, sx = _v(x, "x"); pf, sf = _v(flow, "flow"); pg, sg = _v(grad_out, "grad_out")
    qx, tx = _v(gx, "grad_x") if gx is not None else (None, 0)
    qf, tf = _v(gf, "grad_flow") if gf is not None else (None, 0)
    _launch("warp_bwd", (B, C, H, W), _lib.load().irr_warp_bwd, px, sx, pf, sf, _p(lx, "lin_x", x, W), _p(ly, "lin_y", x, H),
            pg, sg, qx, tx, qf, tf, B, C, H, W, height_im, width_im, div_flow, _grid_mode, _stream())
    return gx, gf


def correlation_generic(in1, in2, pad_size, kernel_size, max_displacement, stride1, stride2):
    import ctypes
    B, C, H, W = in1.shape
    if tuple(in2.shape) != tuple(in1.shape):
        raise RuntimeError(f"irr_b200: correlation inputs differ in shape ({tuple(in1.shape)} vs {tuple(in2.shape)})")
    in1 = in1.contiguous(); in2 = in2.contiguous()
    p1, p2 = _p(in1, "input1"), _p(in2, "input2", in1)
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib = _lib.load()
    _lib.check(lib.irr_correlation_generic_out_shape(H, W, pad_size, kernel_size, max_displacement, stride1, stride2,
                                                     ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)),
               "correlation_generic_out_shape")
    out = torch.empty((B, oc.value, oh.value, ow.value), dtype=torch.float32, device=in1.device)
    _launch("correlation_generic", (B, C, H, W), lib.irr_correlation_generic_fwd, p1, p2,
            out.data_ptr(), B, C, H, W, pad_size, kernel_size, max_displacement, stride1, stride2, _stream())
    return out


def conv_out_hw(H: int, W: int, ks: int, stride: int, dil: int) -> Tuple[int, int]:
    pad = ((ks - 1) * dil) // 2
    return (H + 2 * pad - dil * (ks - 1) - 1) // stride + 1, (W + 2 * pad - dil * (ks - 1) - 1) // stride + 1


def direct_supported(Cout, Cin, ks) -> bool:
    """The layer runs on the direct thin-layer kernel of the fp32 path (include/irr_b200.h irr_conv2d_direct_supported)."""
    return bool(_lib.load().irr_conv2d_direct_supported(Cout, Cin, ks))


def tc_supported(Cout, Cin, ks, stride, dil, math: int = MATH_TC_3XTF32) -> bool:
    return bool(_lib.load().irr_conv2d_math_supported(Cout, Cin, ks, stride, dil, math))


def pack_weights(w: torch.Tensor, math: int = MATH_FP32_SIMT) -> torch.Tensor:
    Cout, Cin, ks, ks2 = w.shape
    assert ks == ks2
    lib = _lib.load()
    n = lib.irr_conv2d_packed_bytes(Cout, Cin, ks, math)
    if n == 0:
        raise RuntimeError(f"irr_b200: no packed layout for conv {Cout}x{Cin}x{ks} math={math}")
    packed = torch.empty(n // 4, dtype=torch.float32, device=w.device)
    wc = w.detach().contiguous().float()
    _launch("conv2d_pack_weights", (Cout, Cin, ks), lib.irr_conv2d_pack_weights, _p(wc, "weight"), packed.data_ptr(), Cout,
            Cin, ks, math, _stream())
    return packed


_ws_bytes = {}


def _conv_ws(lib, x, B, Cin, H, W, Cout, ks, stride, dil, math):
    """split-K scratch for layers with far fewer tiles than SMs (coarse pyramid levels); 0 bytes = never split"""
    key = (B, Cin, H, W, Cout, ks, stride, dil, math)
    nws = _ws_bytes.get(key)
    if nws is None:
        nws = lib.irr_conv2d_workspace_bytes(B, Cin, H, W, Cout, ks, stride, dil, math)
        _ws_bytes[key] = nws
    return (torch.empty(nws // 4, dtype=torch.float32, device=x.device) if nws else None), nws


def conv2d(x, packed, bias, Cout: int, ks: int, stride: int = 1, dil: int = 1, slope: float = 0.1, out=None,
           addend=None, alpha: float = 1.0, math: int = MATH_FP32_SIMT):
    B, Cin, H, W = x.shape
    Ho, Wo = conv_out_hw(H, W, ks, stride, dil)
    if out is None:   # only the 3xF16 path and the direct thin-layer kernel understand pitched rows
        out = _new(x, Cout, Ho, Wo, pitched=(_pitch_enabled and (math == MATH_TC_3XF16 or
                                                                 (math == MATH_FP32_SIMT and direct_supported(Cout, Cin, ks)))))
    assert out.shape == (B, Cout, Ho, Wo), (out.shape, (B, Cout, Ho, Wo))
    px, sx, qx = _vp(x, "x"); po, so, qo = _vp(out, "out")
    pa, sa, qa = (_vp(addend, "addend") if addend is not None else (None, 0, qo))
    if addend is not None:
        assert addend.shape == out.shape
    qo = _pitch("conv2d", qo, qa)
    lib = _lib.load()
    ws, nws = _conv_ws(lib, x, B, Cin, H, W, Cout, ks, stride, dil, math)
    _launch("conv2d", (B, Cin, H, W, Cout, ks, stride, dil, Ho, Wo, math), lib.irr_conv2d_fwd_ws, px, sx,
            _p(packed, "packed weights", x), _p(bias, "bias", x, Cout), pa, sa, po, so, B, Cin, H, W, Cout, ks, stride, dil, slope, alpha, math,
            ws.data_ptr() if ws is not None else None, nws, qx, qo, _stream())
    return out


def conv2d_dual(x, packed, bias, Cout: int, n_split: int, ks: int, out, out2, stride: int = 1, dil: int = 1,
                slope: float = 0.1, addend=None, alpha: float = 1.0, slope2: float = 1.0, addend2=None,
                alpha2: float = 1.0, math: int = MATH_TC_3XF16):
    """Two layers over the same input in one pass (include/irr_b200.h irr_conv2d_fwd_dual): channels [0, n_split) ->
    ``out`` with (slope, addend, alpha); channels [n_split, Cout) -> ``out2`` with (slope2, addend2, alpha2)."""
    B, Cin, H, W = x.shape
    Ho, Wo = conv_out_hw(H, W, ks, stride, dil)
    assert out.shape == (B, n_split, Ho, Wo) and out2.shape == (B, Cout - n_split, Ho, Wo)
    px, sx, qx = _vp(x, "x"); po, so, qo = _vp(out, "out"); po2, so2, qo2 = _vp(out2, "out2")
    pa, sa, qa = (_vp(addend, "addend") if addend is not None else (None, 0, qo))
    pa2, sa2, qa2 = (_vp(addend2, "addend2") if addend2 is not None else (None, 0, qo))
    qo = _pitch("conv2d_dual", qo, qo2, qa, qa2)
    lib = _lib.load()
    ws, nws = _conv_ws(lib, x, B, Cin, H, W, Cout, ks, stride, dil, math)
    _launch("conv2d", (B, Cin, H, W, Cout, ks, stride, dil, Ho, Wo, math), lib.irr_conv2d_fwd_dual, px, sx,
            _p(packed, "packed weights", x), _p(bias, "bias", x, Cout), pa, sa, po, so, B, Cin, H, W, Cout, ks, stride, dil, slope, alpha, n_split,
            pa2, sa2, po2, so2, slope2, alpha2, math, ws.data_ptr() if ws is not None else None, nws, qx, qo, _stream())
    return out, out2


def conv2d_multi(x, packed, bias, Cout: int, ks: int, segs, stride: int = 1, dil: int = 1, math: int = MATH_TC_3XF16):
    """One conv pass whose output channels are cut into segments with their own destination and epilogue
    (include/irr_b200.h irr_conv2d_fwd_multi).  ``segs``: list of dicts ``{n_begin, out, slope=1.0, alpha=1.0,
    addend=None, pre=False}`` in increasing n_begin (multiples of 16, the first 0); segment i holds channels
    [n_begin_i, n_begin_{i+1}) and ``out`` must have exactly that many channels."""
    B, Cin, H, W = x.shape
    Ho, Wo = conv_out_hw(H, W, ks, stride, dil)
    px, sx, qx = _vp(x, "x")
    arr = (_lib.ConvSeg * len(segs))()
    qs = []
    for i, sg in enumerate(segs):
        n_end = segs[i + 1]["n_begin"] if i + 1 < len(segs) else Cout
        out = sg["out"]
        assert out.shape == (B, n_end - sg["n_begin"], Ho, Wo), (tuple(out.shape), (B, n_end - sg["n_begin"], Ho, Wo))
        po, so, qo = _vp(out, f"out[{i}]")
        qs.append(qo)
        add = sg.get("addend")
        if add is not None:
            assert add.shape == out.shape
        pa, sa, qa = (_vp(add, f"addend[{i}]") if add is not None else (None, 0, qo))
        qs.append(qa)
        arr[i].n_begin = sg["n_begin"]; arr[i].addend_pre = 1 if sg.get("pre") else 0
        arr[i].leaky_slope = sg.get("slope", 1.0); arr[i].alpha = sg.get("alpha", 1.0)
        arr[i].addend = pa; arr[i].addend_bs = sa; arr[i].y = po; arr[i].y_bs = so
    qo = _pitch("conv2d_multi", *qs)
    lib = _lib.load()
    ws, nws = _conv_ws(lib, x, B, Cin, H, W, Cout, ks, stride, dil, math)
    _launch("conv2d", (B, Cin, H, W, Cout, ks, stride, dil, Ho, Wo, math), lib.irr_conv2d_fwd_multi, px, sx,
            _p(packed, "packed weights", x), _p(bias, "bias", x, Cout), B, Cin, H, W, Cout, ks, stride, dil, arr, len(segs),
            math, ws.data_ptr() if ws is not None else None, nws, qx, qo, _stream())
    return [sg["out"] for sg in segs]


def resize_ac(x, OH: int, OW: int, out=None, s_even: float = 1.0, s_odd: float = 1.0, pitched: Optional[bool] = None):
    B, C, H, W = x.shape
    if out is None:
        out = _new(x, C, OH, OW, pitched=pitched)
    px, sx, qx = _vp(x, "x"); po, so, qo = _vp(out, "out")
    _launch("resize_bilinear_ac", None, _lib.load().irr_resize_bilinear_ac_fwd, px, sx, po, so, B, C, H, W, OH, OW, s_even,
            s_odd, qx, qo, _stream())
    return out


def scale_channels(x, out=None, s_even: float = 1.0, s_odd: float = 1.0):
    B, C, H, W = x.shape
    if out is None:
        out = _new(x, C, H, W, pitched=_is_pitched(x))
    px, sx, qx = _vp(x, "x"); po, so, qo = _vp(out, "out")
    if qx and qo and qx != qo:   # dense <-> pitched copy: the resize kernel at identical size is an exact copy with both pitches
        return resize_ac(x, H, W, out=out, s_even=s_even, s_odd=s_odd)
    _launch("scale_channels", None, _lib.load().irr_scale_channels_fwd, px, sx, po, so, B, C, H * (qx or qo or W), s_even, s_odd,
            _stream())
    return out


DTYPE_F32, DTYPE_BF16 = 0, 1


def empty_bf16(B: int, C: int, H: int, W: int, device) -> torch.Tensor:
    """A (B, C, H, W) bf16 tensor whose rows are padded to a multiple of 8 elements (16 bytes): what the bf16 correlation
    path needs (irr_warp_correlation_fwd_dt)."""
    P = (W + 7) // 8 * 8
    t = torch.empty((B, C, H, P), dtype=torch.bfloat16, device=device)
    return t if P == W else t[:, :, :, :W]


def round_bf16_store(x, round_in_place: bool = True, out16=None):
    """Packed bf16 copy of ``x`` (round to nearest even) for the correlation's bf16 input path; with ``round_in_place`` ``x``
    itself is also replaced by float(bf16(x)), as round_bf16(x, out=x) does — one pass over x for both."""
    B, C, H, W = x.shape
    if out16 is None:
        out16 = empty_bf16(B, C, H, W, x.device)
    px, sx, qx = _vp(x, "x"); p16, s16, q16 = _vp(out16, "out16", torch.bfloat16)
    _launch("round_bf16_store", None, _lib.load().irr_round_bf16_store_fwd, px, sx, qx, px if round_in_place else None, sx, p16,
            s16, q16, B, C, H, W, _stream())
    return out16


def round_bf16(x, out=None):
    """y = x.bfloat16().float() (BASELINE config 5: bf16-valued features in the fp32 layout); ``out=x`` rounds in place."""
    B, C, H, W = x.shape
    if out is None:
        out = _new(x, C, H, W, pitched=_is_pitched(x))
    px, sx, qx = _vp(x, "x"); po, so, qo = _vp(out, "out")
    P = _pitch("round_bf16", qx, qo) or W
    _launch("round_bf16", None, _lib.load().irr_round_bf16_fwd, px, sx, po, so, B, C, H * P, _stream())
    return out


def upsample_nearest2x(x, OH: int, OW: int, out=None):
    B, C, H, W = x.shape
    if out is None:
        out = _new(x, C, OH, OW)
    px, sx, qx = _vp(x, "x"); po, so, qo = _vp(out, "out")
    _launch("upsample_nearest2x", None, _lib.load().irr_upsample_nearest2x_fwd, px, sx, po, so, B, C, H, W, OH, OW, qx, qo,
            _stream())
    return out


def sub_spatial_mean(x, out=None):
    B, C, H, W = x.shape
    if out is None:
        out = _new(x, C, H, W, pitched=_is_pitched(x))
    px, sx, qx = _vp(x, "x"); po, so, qo = _vp(out, "out")
    P = _pitch("sub_spatial_mean", qx, qo)
    _launch("sub_spatial_mean", None, _lib.load().irr_sub_spatial_mean_fwd, px, sx, po, so, B, C, H, W, P, _stream())
    return out


def channel_l2norm(x, out=None):
    B, C, H, W = x.shape
    if out is None:
        out = _new(x, 1, H, W, pitched=_is_pitched(x))
    px, sx, qx = _vp(x, "x"); po, so, qo = _vp(out, "out")
    P = _pitch("channel_l2norm", qx, qo) or W
    _launch("channel_l2norm", None, _lib.load().irr_channel_l2norm_fwd, px, sx, po, so, B, C, H * P, _stream())
    return out


def refine_gather(logits, src, out=None):
    B, C, H, W = src.shape
    assert logits.shape == (B, 9, H, W)
    if out is None:
        out = _new(src, C, H, W, pitched=_is_pitched(src))
    pl, sl, ql = _vp(logits, "logits"); ps, ss, qs = _vp(src, "src"); po, so, qo = _vp(out, "out")
    P = _pitch("refine_gather", ql, qs, qo)
    _launch("refine_gather", None, _lib.load().irr_refine_gather_fwd, pl, sl, ps, ss, po, so, B, C, H, W, P, _stream())
    return out


def eval_metrics(flow, target, valid=None, occ_logits=None, target_occ=None):
    """Per-image sums for the reference's eval-mode losses (include/irr_b200.h irr_eval_metrics_fwd): float64 (B, 8)
    tensor { S epe*valid, S valid, S outlier, S pred*true, S pred, S true, 0, 0 }."""
    B, C, H, W = flow.shape
    assert C == 2 and tuple(target.shape) == (B, 2, H, W)
    pf, sf = _v(flow, "flow"); pt, st = _v(target, "target")
    pv, sv = _v(valid, "valid") if valid is not None else (None, 0)
    po, so = _v(occ_logits, "occ_logits") if occ_logits is not None else (None, 0)
    pq, sq = _v(target_occ, "target_occ") if target_occ is not None else (None, 0)
    sums = torch.empty((B, 8), dtype=torch.float64, device=flow.device)
    _launch("eval_metrics", None, _lib.load().irr_eval_metrics_fwd, pf, sf, pt, st, pv, sv, po, so, pq, sq,
            sums.data_ptr(), B, H, W, _stream())
    return sums


def correlation_backward(f1, f2, grad_out, need_f1: bool = True, need_f2: bool = True, max_disp: int = 4):
    """(grad_f1, grad_f2) of ``correlation(f1, f2)`` (no LeakyReLU, no batch shift) — include/irr_b200.h
    irr_correlation_bwd.  A gradient that is not needed is returned as None and not computed."""
    B, C, H, W = f1.shape
    D = (2 * max_disp + 1) ** 2
    assert f2.shape == f1.shape and tuple(grad_out.shape) == (B, D, H, W)
    g1 = torch.empty_like(f1) if need_f1 else None
    g2 = torch.empty_like(f2) if need_f2 else None
    p1, s1 = _v(f1, "f1"); p2, s2 = _v(f2, "f2"); pg, sg = _v(grad_out, "grad_out")
    q1, t1 = _v(g1, "grad_f1") if g1 is not None else (None, 0)
    q2, t2 = _v(g2, "grad_f2") if g2 is not None else (None, 0)
    _launch("correlation_bwd", (B, C, H, W), _lib.load().irr_correlation_bwd, p1, s1, p2, s2, pg, sg, q1, t1, q2, t2, B, C,
            H, W, max_disp, _stream())
    return g1, g2
