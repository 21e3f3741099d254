"""Op-level parity of the C-ABI kernels against the oracle (numpy / torch-CPU restatements pinned to the reference)
and against the committed golden vectors.  Tolerances: 1e-4 abs (north_star) for values — tighter where the op is a
pure gather — and BITWISE for the warp validity mask."""
import numpy as np
import pytest
import torch

from oracle import irr_oracle as O
from oracle import ops_np as N

pytestmark = pytest.mark.gpu


def rs(seed, shape, kind="normal"):
    r = np.random.RandomState(seed)
    a = r.standard_normal(shape) if kind == "normal" else np.abs(r.standard_normal(shape)) * 0.3
    return a.astype("float32")


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


# ------------------------------------------------------------------ cost volume
CV_SHAPES = [(1, 64, 64, 128), (2, 196, 7, 16), (2, 32, 109, 256), (1, 96, 24, 78), (1, 3, 5, 5), (3, 17, 9, 13)]


@pytest.mark.parametrize("si", range(len(CV_SHAPES)))
@pytest.mark.parametrize("kind", ["normal", "lrelu"])
def test_cost_volume_golden(cuda, golden_dir, si, kind):
    from irr_b200 import ops
    g = np.load(f"{golden_dir}/cost_volume.npz")
    shape = CV_SHAPES[si]
    key = "x".join(map(str, shape)) + "_" + kind
    seed = int(g[key + "__seed"])
    f1, f2 = rs(seed, shape, kind), rs(seed + 1000, shape, kind)
    out = ops.correlation(dev(f1, cuda), dev(f2, cuda)).cpu().numpy()
    sub = out if out.size <= 60000 else out[:, :, ::5, ::7]
    assert np.abs(sub - g[key + "__sub"]).max() <= 1e-4
    np.testing.assert_allclose(out.sum(axis=(2, 3), dtype=np.float64), g[key + "__sum"], atol=2e-2, rtol=1e-4)
    # full tensor vs the numpy oracle
    assert np.abs(out - N.cost_volume_np(f1, f2)).max() <= 1e-4


def test_cost_volume_analytic(cuda):
    """Delta images prove channel order (dy+4)*9+(dx+4); constant images prove zero padding and the /C."""
    from irr_b200 import ops
    C, H, W = 5, 12, 40
    f1 = np.zeros((1, C, H, W), "float32"); f2 = np.zeros((1, C, H, W), "float32")
    f1[0, 2, 6, 20] = 3.0
    f2[0, 2, 6 + 2, 20 - 3] = 5.0  # dy=+2, dx=-3
    out = ops.correlation(dev(f1, cuda), dev(f2, cuda)).cpu().numpy()
    ch = (2 + 4) * 9 + (-3 + 4)
    assert out[0, ch, 6, 20] == pytest.approx(15.0 / C)
    out[0, ch, 6, 20] = 0
    assert np.abs(out).max() == 0
    ones = np.ones((1, C, H, W), "float32")
    out = ops.correlation(dev(ones, cuda), dev(ones, cuda)).cpu().numpy()
    assert out[0, 40, 5, 5] == pytest.approx(1.0)
    assert out[0, 0, 0, 0] == 0.0 and out[0, 80, H - 1, W - 1] == 0.0  # (dy,dx)=(-4,-4) at the top-left corner
    assert out[0, 80, 0, 0] == pytest.approx(1.0)


def test_cost_volume_slice_shift_lrelu(cuda):
    """Write into a channel slice of a bigger buffer, rotate the f2 batch, fuse LeakyReLU."""
    from irr_b200 import ops
    B, C, H, W = 4, 32, 28, 64
    f = rs(1, (B, C, H, W))
    ft = dev(f, cuda)
    buf = torch.full((B, 120, H, W), 7.0, device=cuda)
    ops.correlation(ft, ft, out=buf[:, 10:91], shift=2, slope=0.1)
    ref = N.cost_volume_np(f, np.roll(f, -2, axis=0))
    ref = np.where(ref > 0, ref, ref * np.float32(0.1))
    got = buf.cpu().numpy()
    assert np.abs(got[:, 10:91] - ref).max() <= 1e-4
    assert (got[:, :10] == 7.0).all() and (got[:, 91:] == 7.0).all()


def test_correlation_generic_vs_c_oracle(cuda):
    from irr_b200 import ops
    f1, f2 = rs(5, (2, 6, 20, 24)), rs(6, (2, 6, 20, 24))
    for (pad, k, md, s1, s2) in [(4, 1, 4, 1, 1), (20, 1, 20, 1, 2), (3, 3, 2, 2, 1), (4, 1, 4, 1, 2)]:
        ref = N.corr_ref_c(f1, f2, pad, k, md, s1, s2)
        got = ops.correlation_generic(dev(f1, cuda), dev(f2, cuda), pad, k, md, s1, s2).cpu().numpy()
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-5


# ------------------------------------------------------------------ warp
@pytest.mark.parametrize("ci", range(5))
def test_warp_golden_mask_bitexact(cuda, golden_dir, ci):
    from irr_b200 import ops
    ops.set_grid_mode(ops.GRID_TRUE_DIV)
    g = np.load(f"{golden_dir}/warp.npz")
    seed, B, C, H, W, him, wim = [int(v) for v in g[f"case{ci}__meta"]]
    x = rs(seed, (B, C, H, W))
    flow = g[f"case{ci}__flow"]
    lx, ly = dev(g[f"case{ci}__lin_x"], cuda), dev(g[f"case{ci}__lin_y"], cuda)
    mask = torch.empty((B, H, W), device=cuda)
    out = ops.warp(dev(x, cuda), dev(flow, cuda), him, wim, 0.05, mask_out=mask, lin_x=lx, lin_y=ly)
    assert (mask.cpu().numpy() != g[f"case{ci}__mask"]).sum() == 0  # bitwise
    assert np.abs(out.cpu().numpy() - g[f"case{ci}__out"]).max() <= 2e-6


def test_warp_identity_and_minuend(cuda):
    from irr_b200 import ops
    x = rs(11, (2, 7, 13, 39))
    xt = dev(x, cuda)
    zero = torch.zeros((2, 2, 13, 39), device=cuda)
    mask = torch.empty((2, 13, 39), device=cuda)
    out = ops.warp(xt, zero, 375, 1242, 0.05, mask_out=mask)
    assert (mask == 1).all()
    # zero flow is the identity only up to the rounding of ((lin+1)/2)*(W-1) — same as the reference
    assert (out - xt).abs().max().item() <= 2e-5
    assert (out.cpu() - O.warp(torch.from_numpy(x), torch.zeros(2, 2, 13, 39), 375, 1242, 0.05)).abs().max().item() <= 2e-6
    d = ops.warp(xt, zero, 375, 1242, 0.05, minuend=xt, shift=1)
    ref = x - np.roll(out.cpu().numpy(), -1, axis=0)
    assert np.abs(d.cpu().numpy() - ref).max() <= 1e-6


@pytest.mark.parametrize("mode", ["true_div", "recip"])
def test_warp_matches_torch_device_ops(cuda, mode):
    """TRUE_DIV reproduces the reference run on the CPU; RECIP_MUL reproduces the reference's torch ops run on THIS
    GPU (torch's CUDA `tensor / scalar` multiplies by the reciprocal).  Mask must be bit-identical in both."""
    from irr_b200 import ops
    B, C, H, W, him, wim = 2, 8, 55, 128, 436, 1024
    x = torch.from_numpy(rs(21, (B, C, H, W)))
    flow = torch.from_numpy(rs(22, (B, 2, H, W))) * torch.tensor([wim / W, him / H]).view(1, 2, 1, 1) * 0.05 * 6.0
    if mode == "true_div":
        ops.set_grid_mode(ops.GRID_TRUE_DIV)
        grid = O.sampling_grid(flow, him, wim, 0.05)
        ref = O.warp(x, flow, him, wim, 0.05)
        m = (torch.nn.functional.grid_sample(torch.ones_like(x), grid, align_corners=True) >= 1.0)[:, 0]
    else:
        # NB: with cuDNN enabled torch routes this grid_sample (4-D, bilinear, zeros, align_corners, C <= 1024) to
        # cudnnSpatialTfSamplerForward, whose internal rounding is not ATen's (1 of 14080 mask pixels differed on the
        # B200); the recipe in common.cuh reproduces ATen's own grid_sampler_2d_kernel, so pin that implementation.
        ops.set_grid_mode(ops.GRID_RECIP_MUL)
        xg, fg = x.to(cuda), flow.to(cuda)
        with torch.backends.cudnn.flags(enabled=False):
            grid = O.sampling_grid(fg, him, wim, 0.05)
            ref = O.warp(xg, fg, him, wim, 0.05).cpu()
            m = (torch.nn.functional.grid_sample(torch.ones_like(xg), grid, align_corners=True) >= 1.0)[:, 0].cpu()
    try:
        mask = torch.empty((B, H, W), device=cuda)
        out = ops.warp(x.to(cuda), flow.to(cuda), him, wim, 0.05, mask_out=mask)
    finally:
        ops.set_grid_mode(ops.GRID_TRUE_DIV)
    nbad = int((mask.cpu() != m.float()).sum())
    assert nbad == 0, f"{nbad} mask pixels differ ({mode})"
    assert (out.cpu() - ref).abs().max().item() <= 2e-6


def test_warp_correlation_fused_equals_unfused(cuda):
    from irr_b200 import ops
    B, C, H, W, him, wim = 4, 32, 28, 64, 436, 1024
    f = dev(rs(31, (B, C, H, W)), cuda)
    flow = dev(rs(32, (B, 2, H, W)), cuda) * torch.tensor([wim / W, him / H], device=cuda).view(1, 2, 1, 1) * 0.05 * 3.0
    fused = ops.warp_correlation(f, f, flow, him, wim, 0.05, shift=2, slope=0.1)
    w = ops.warp(f, flow, him, wim, 0.05, shift=2)
    unfused = ops.correlation(f, w, slope=0.1)
    assert (fused - unfused).abs().max().item() <= 1e-6
    # and against the oracle on the CPU
    fc, flc = f.cpu(), flow.cpu()
    ref = torch.nn.functional.leaky_relu(O.cost_volume(fc, O.warp(torch.roll(fc, -2, 0), flc, him, wim, 0.05)), 0.1)
    assert (fused.cpu() - ref).abs().max().item() <= 1e-4


def test_cost_volume_full_size_properties(cuda):
    """BASELINE config 3's largest correlation launch (2B=16, C=32, 109 x 256 — too big for the CPU oracle in a unit
    test) through size-independent properties: displacement symmetry (exact: the same products in the same order),
    linearity in f1, the fused warp == warp kernel + plain kernel, a strided sample of pixels against a
    direct dot product, and the cp.async fallback kernel == the TMA kernel (exact)."""
    import os
    from irr_b200 import ops
    B, C, H, W = 16, 32, 109, 256
    g = torch.Generator(device="cpu").manual_seed(123)
    f1 = torch.randn(B, C, H, W, generator=g).to(cuda)
    f2 = torch.randn(B, C, H, W, generator=g).to(cuda)
    g1 = torch.randn(B, C, H, W, generator=g).to(cuda)
    a = ops.correlation(f1, f2)
    # symmetry: corr(f1, f2)[dy, dx](y, x) == corr(f2, f1)[-dy, -dx](y + dy, x + dx)
    b = ops.correlation(f2, f1)
    for dy, dx in [(-4, -4), (0, 3), (2, -1), (4, 4), (0, 0)]:
        ch, chm = (dy + 4) * 9 + (dx + 4), (-dy + 4) * 9 + (-dx + 4)
        ys, ye = max(0, -dy), min(H, H - dy)
        xs, xe = max(0, -dx), min(W, W - dx)
        assert torch.equal(a[:, ch, ys:ye, xs:xe], b[:, chm, ys + dy:ye + dy, xs + dx:xe + dx]), (dy, dx)
    # linearity in f1
    lin = ops.correlation(2.5 * f1 + g1, f2)
    assert (lin - (2.5 * a + ops.correlation(g1, f2))).abs().max().item() <= 2e-5
    # fused warp == warp kernel followed by the plain kernel, for a smooth sub-pixel flow and for zero flow (which is
    # NOT the identity at every size: the rounded linspace grid leaves some pixels with a weight sum < 1)
    lo = torch.randn(B, 2, 3, 4, generator=g).to(cuda)
    smooth = torch.nn.functional.interpolate(lo, size=[H, W], mode="bicubic", align_corners=True) * 0.4
    for flow in (smooth, torch.zeros(B, 2, H, W, device=cuda)):
        fused = ops.warp_correlation(f1, f2, flow, 436, 1024, 0.05, shift=B // 2)
        unfused = ops.correlation(f1, ops.warp(f2, flow, 436, 1024, 0.05, shift=B // 2))
        assert (fused - unfused).abs().max().item() <= 1e-6
    # direct dot products at a strided sample of pixels / displacements
    f2p = torch.nn.functional.pad(f2, (4, 4, 4, 4))
    for (y, x, dy, dx) in [(0, 0, -4, -4), (0, 255, 4, 4), (108, 0, 4, -4), (54, 128, 1, -2), (108, 255, -3, 0), (7, 31, 0, 4)]:
        want = (f1[:, :, y, x] * f2p[:, :, y + dy + 4, x + dx + 4]).sum(1) / C
        assert (a[:, (dy + 4) * 9 + (dx + 4), y, x] - want).abs().max().item() <= 1e-5
    # the tap table of the pre-pass gives bit-identical results to the in-kernel tap set-up (same recipe, same order)
    # ... and so do the two experimental variants kept for A/B runs (both measured slower): the tap table
    # from a pre-pass (IRR_CORR_PRETAB=1) and 7-row tiles with 8 compute warps (IRR_CORR_TH7=1)
    from irr_b200 import ops as _ops
    fused_ref = ops.warp_correlation(f1, f2, smooth, 436, 1024, 0.05, shift=B // 2, slope=0.1)
    for var in ({"IRR_CORR_PRETAB": "1"}, {"IRR_CORR_TH7": "1"}, {"IRR_CORR_PRETAB": "1", "IRR_CORR_TH7": "1"}):
        os.environ.update(var)
        _ops._corr_ws_bytes.clear()   # the workspace size depends on the variant
        try:
            assert torch.equal(ops.warp_correlation(f1, f2, smooth, 436, 1024, 0.05, shift=B // 2, slope=0.1), fused_ref), var
        finally:
            for k in var:
                del os.environ[k]
            _ops._corr_ws_bytes.clear()
    wild = torch.randn(B, 2, H, W, generator=g).to(cuda) * 3.0      # divergent flow: tiles fall back to global gathers
    fw = ops.warp_correlation(f1, f2, wild, 436, 1024, 0.05, shift=B // 2)
    assert (fw - ops.correlation(f1, ops.warp(f2, wild, 436, 1024, 0.05, shift=B // 2))).abs().max().item() <= 1e-6
    # the cp.async fallback kernel computes the same sums in the same order
    os.environ["IRR_CORR_NO_TMA"] = "1"
    try:
        assert torch.equal(ops.correlation(f1, f2), a)
    finally:
        del os.environ["IRR_CORR_NO_TMA"]


@pytest.mark.parametrize("shape", [(16, 196, 7, 16), (16, 128, 14, 32), (2, 196, 7, 16), (2, 96, 28, 64), (1, 20, 9, 40)])
def test_cost_volume_channel_split(cuda, shape):
    """Coarse pyramid levels: the channel chunks of a tile are dealt to several CTAs and a second launch adds the partial
    sums in a fixed order (irr_warp_correlation_fwd_ws).  Plain and fused, against the un-split kernel (same products,
    different association: <= 1e-6), against the oracle, and bit-reproducible from run to run."""
    import os
    from irr_b200 import ops, _lib
    B, C, H, W = shape
    f1 = dev(rs(41, (B, C, H, W)), cuda)
    f2 = dev(rs(42, (B, C, H, W)), cuda)
    flow = dev(rs(43, (B, 2, H, W)), cuda) * 0.3
    nws = _lib.load().irr_correlation_workspace_bytes(B, C, H, W, 0)
    if shape[0] * ((H + 7) // 8) * ((W + 31) // 32) * 2 <= 148 and C > 8:
        assert nws > 0      # these launches are split on a 148-SM part
    a = ops.correlation(f1, f2, shift=B // 2, slope=0.1)
    fz = ops.warp_correlation(f1, f2, flow, 16 * H, 16 * W, 0.05, shift=B // 2, slope=0.1)
    assert torch.equal(a, ops.correlation(f1, f2, shift=B // 2, slope=0.1))
    assert torch.equal(fz, ops.warp_correlation(f1, f2, flow, 16 * H, 16 * W, 0.05, shift=B // 2, slope=0.1))
    os.environ["IRR_CORR_NO_SPLIT"] = "1"
    try:
        a0 = ops.correlation(f1, f2, shift=B // 2, slope=0.1)
        fz0 = ops.warp_correlation(f1, f2, flow, 16 * H, 16 * W, 0.05, shift=B // 2, slope=0.1)
    finally:
        del os.environ["IRR_CORR_NO_SPLIT"]
    assert (a - a0).abs().max().item() <= 1e-6 and (fz - fz0).abs().max().item() <= 1e-6
    ref = torch.nn.functional.leaky_relu(O.cost_volume(f1.cpu(), torch.roll(f2.cpu(), -(B // 2), 0)), 0.1)
    assert (a.cpu() - ref).abs().max().item() <= 1e-5
    refw = torch.nn.functional.leaky_relu(
        O.cost_volume(f1.cpu(), O.warp(torch.roll(f2.cpu(), -(B // 2), 0), flow.cpu(), 16 * H, 16 * W, 0.05)), 0.1)
    assert (fz.cpu() - refw).abs().max().item() <= 1e-4


# ------------------------------------------------------------------ conv
CONV_CASES = [
    # (B, Cin, H, W, Cout, k, stride, dil)
    (2, 3, 64, 96, 16, 3, 2, 1), (2, 16, 32, 48, 16, 3, 1, 1), (1, 128, 14, 32, 196, 3, 2, 1),
    (2, 115, 28, 64, 128, 3, 1, 1), (1, 563, 14, 32, 2, 3, 1, 1), (1, 565, 28, 64, 128, 3, 1, 1),
    (1, 128, 28, 64, 128, 3, 1, 2), (1, 128, 28, 64, 96, 3, 1, 8), (1, 96, 28, 64, 64, 3, 1, 16),
    (2, 196, 7, 16, 32, 1, 1, 1), (1, 16, 47, 156, 3, 1, 1, 1), (1, 35, 13, 39, 128, 3, 1, 1),
    (1, 32, 21, 37, 9, 3, 1, 1), (1, 11, 47, 155, 32, 3, 1, 1), (1, 32, 24, 78, 1, 3, 1, 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_vs_torch_cpu(cuda, case):
    from irr_b200 import ops
    B, Cin, H, W, Cout, k, s, d = case
    x = torch.from_numpy(rs(41, (B, Cin, H, W)))
    w = torch.from_numpy(rs(42, (Cout, Cin, k, k))) * float(np.sqrt(2.0 / (Cin * k * k)))
    b = torch.from_numpy(rs(43, (Cout,))) * 0.1
    ref = torch.nn.functional.leaky_relu(
        torch.nn.functional.conv2d(x, w, b, stride=s, padding=((k - 1) * d) // 2, dilation=d), 0.1)
    packed = ops.pack_weights(w.to(cuda))
    got = ops.conv2d(x.to(cuda), packed, b.to(cuda), Cout, k, s, d, slope=0.1)
    assert got.shape == ref.shape
    assert (got.cpu() - ref).abs().max().item() <= 1e-4


def test_conv2d_slices_addend_alpha(cuda):
    """Dense-block style: read a channel suffix of the concat buffer, write a slice, y = addend + alpha*(conv+b)."""
    from irr_b200 import ops
    B, Ct, H, W = 2, 60, 20, 36
    buf = torch.from_numpy(rs(51, (B, Ct, H, W))).to(cuda)
    w = torch.from_numpy(rs(52, (24, 40, 3, 3))) * 0.05
    b = torch.from_numpy(rs(53, (24,))) * 0.1
    add = torch.from_numpy(rs(54, (B, 24, H, W))).to(cuda)
    xin = buf[:, 20:60].cpu()
    ref = add.cpu() + 0.1 * torch.nn.functional.conv2d(xin, w, b, padding=1)
    before = buf.clone()
    out = torch.zeros((B, 30, H, W), device=cuda)
    ops.conv2d(buf[:, 20:60], ops.pack_weights(w.to(cuda)), b.to(cuda), 24, 3, slope=1.0, out=out[:, 3:27],
               addend=add, alpha=0.1)
    assert (out[:, 3:27].cpu() - ref).abs().max().item() <= 1e-4
    assert (out[:, :3] == 0).all() and (out[:, 27:] == 0).all() and torch.equal(buf, before)


# ------------------------------------------------------------------ small ops
def test_resize_scale_nearest(cuda, golden_dir):
    from irr_b200 import ops
    g = np.load(f"{golden_dir}/modules.npz")
    t = rs(310, (2, 2, 7, 16))
    a = ops.resize_ac(dev(t, cuda), 14, 32).cpu().numpy()
    b = ops.resize_ac(dev(t, cuda), 13, 39).cpu().numpy()
    assert np.abs(a - g["resize_ac__out"]).max() <= 1e-6 and np.abs(b - g["resize_ac__odd"]).max() <= 1e-6
    c = ops.resize_ac(dev(t, cuda), 13, 39, s_even=2.0, s_odd=3.0).cpu().numpy()
    assert np.abs(c[:, 0] - 2 * b[:, 0]).max() <= 1e-6 and np.abs(c[:, 1] - 3 * b[:, 1]).max() <= 1e-6
    # down-sizing (image pyramid for the refinement input, IRR_PWC.py:126-127)
    img = rs(311, (1, 3, 94, 156))
    ref = torch.nn.functional.interpolate(torch.from_numpy(img), size=[24, 39], mode="bilinear", align_corners=True)
    assert (ops.resize_ac(dev(img, cuda), 24, 39).cpu() - ref).abs().max().item() <= 1e-6
    s = ops.scale_channels(dev(t, cuda), s_even=0.5, s_odd=-2.0).cpu().numpy()
    assert np.array_equal(s[:, 0], t[:, 0] * np.float32(0.5)) and np.array_equal(s[:, 1], t[:, 1] * np.float32(-2))
    o = rs(312, (2, 1, 12, 20))
    for (oh, ow) in [(24, 40), (23, 39), (24, 39)]:
        like = torch.zeros(1, 1, oh, ow)
        ref = O.upsample_x2(torch.from_numpy(o), like)
        got = ops.upsample_nearest2x(dev(o, cuda), oh, ow).cpu()
        assert (got - ref).abs().max().item() <= 5e-6  # CPU vs CUDA-order bilinear (align_corners=False) differ by ulps


@pytest.mark.parametrize("si", [0, 1, 2])
def test_correlation_backward_golden(cuda, golden_dir, si):
    """irr_correlation_bwd vs the gradients of the reference's compute_cost_volume (golden) and the numpy oracle; through
    the autograd Function, including a one-sided gradient."""
    import irr_b200
    from irr_b200 import ops
    g = np.load(f"{golden_dir}/corr_grad.npz")
    shape = tuple(int(v) for v in g[f"shape__{si}"])
    rs = lambda seed, shp: torch.from_numpy(np.random.RandomState(seed).standard_normal(shp).astype("float32"))
    f1, f2, go = rs(500 + si, shape), rs(520 + si, shape), rs(540 + si, (shape[0], 81, shape[2], shape[3]))
    a = f1.to(cuda).requires_grad_(True)
    b = f2.to(cuda).requires_grad_(True)
    corr = irr_b200.Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1)
    out = corr(a, b)
    out.backward(go.to(cuda))
    assert np.abs(a.grad.cpu().numpy() - g[f"g1__{si}"]).max() <= 2e-6
    assert np.abs(b.grad.cpu().numpy() - g[f"g2__{si}"]).max() <= 2e-6
    o1, o2 = N.cost_volume_backward_np(f1.numpy(), f2.numpy(), go.numpy())
    assert np.abs(a.grad.cpu().numpy() - o1).max() <= 2e-6 and np.abs(b.grad.cpu().numpy() - o2).max() <= 2e-6
    g1, none = ops.correlation_backward(f1.to(cuda), f2.to(cuda), go.to(cuda), need_f1=True, need_f2=False)
    assert none is None and torch.equal(g1, a.grad)
    with pytest.raises(NotImplementedError):
        irr_b200.Correlation(20, 1, 20, 1, 2)(a, b)


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_warp_backward_golden(cuda, golden_dir, ci):
    """irr_warp_bwd (through WarpFunction / WarpingLayer under autograd) vs the gradients of the REFERENCE's WarpingLayer
    under autograd (tests/golden/warp_grad.npz, oracle/gen_golden.py warp_grad): d/dx and d/dflow, the hard mask a constant."""
    from irr_b200 import ops, pwc_modules
    g = np.load(f"{golden_dir}/warp_grad.npz")
    seed, B, C, H, W, him, wim = (int(v) for v in g[f"case{ci}__meta"])
    rsn = lambda sd, shp: torch.from_numpy(np.random.RandomState(sd).standard_normal(shp).astype("float32"))
    x = rsn(seed, (B, C, H, W)).to(cuda).requires_grad_(True)
    flow = torch.from_numpy(g[f"case{ci}__flow"]).to(cuda).requires_grad_(True)
    go = rsn(seed + 2000, (B, C, H, W)).to(cuda)
    out = pwc_modules.WarpingLayer()(x, flow, him, wim, 0.05)
    assert out.grad_fn is not None
    out.backward(go)
    gx_ref, gf_ref = g[f"case{ci}__gx"], g[f"case{ci}__gflow"]
    assert np.abs(x.grad.cpu().numpy() - gx_ref).max() <= 1e-5 * max(1.0, np.abs(gx_ref).max())
    assert np.abs(flow.grad.cpu().numpy() - gf_ref).max() <= 1e-4 * max(1.0, np.abs(gf_ref).max())
    # one-sided requests
    gx, none = ops.warp_backward(x.detach(), flow.detach(), go, him, wim, 0.05, need_flow=False)
    assert none is None and (gx - x.grad).abs().max().item() <= 1e-6 * max(1.0, np.abs(gx_ref).max())


def test_autograd_safe_rescale_flow_and_conv(cuda):
    """rescale_flow under autograd leaves its argument untouched and is differentiable (the reference's
    in-place ``u *= scale`` on a chunk view raises on torch >= 2); conv() under autograd = our forward kernel + ATen's
    convolution_backward on the LeakyReLU-masked gradient — both against plain torch autograd."""
    from irr_b200 import ops, pwc_modules
    import torch.nn.functional as F
    fl = torch.from_numpy(rs(7, (2, 2, 9, 13))).to(cuda).requires_grad_(True)
    before = fl.detach().clone()
    r = pwc_modules.rescale_flow(fl, 0.05, 64, 48, to_local=False)
    assert torch.equal(fl.detach(), before) and r.grad_fn is not None
    su, sv = pwc_modules.flow_scales(9, 13, 0.05, 64, 48, False)
    r.sum().backward()
    want = torch.tensor([su, sv], device=cuda).view(1, 2, 1, 1).expand_as(fl)
    assert (fl.grad - want).abs().max().item() <= 1e-6 * max(su, sv)
    with torch.no_grad():   # inference keeps the reference's in-place side effect (F6)
        t = before.clone()
        r2 = pwc_modules.rescale_flow(t, 0.05, 64, 48, to_local=False)
        assert torch.equal(r2, t) and not torch.equal(t, before)
    pwc_modules.set_conv_math(ops.MATH_TC_3XF16)
    blk = pwc_modules.conv(19, 24, kernel_size=3, stride=1, dilation=2).to(cuda)
    x = torch.from_numpy(rs(8, (2, 19, 17, 20))).to(cuda).requires_grad_(True)
    y = blk(x)
    assert y.grad_fn is not None
    go = torch.from_numpy(rs(9, tuple(y.shape))).to(cuda)
    y.backward(go)
    x2 = x.detach().clone().requires_grad_(True)
    w2, b2 = blk[0].weight.detach().clone().requires_grad_(True), blk[0].bias.detach().clone().requires_grad_(True)
    y2 = F.leaky_relu(F.conv2d(x2, w2, b2, padding=2, dilation=2), 0.1)
    y2.backward(go)
    assert (y - y2).abs().max().item() <= 1e-4
    for a, b in ((x.grad, x2.grad), (blk[0].weight.grad, w2.grad), (blk[0].bias.grad, b2.grad)):
        # a handful of outputs within 1e-6 of zero may take the other LeakyReLU branch: compare in aggregate
        assert (a - b).abs().max().item() <= 2e-3 * max(1.0, b.abs().max().item())


def test_round_bf16(cuda):
    """irr_round_bf16_fwd == x.bfloat16().float() bit for bit (RNE, ties, subnormals, inf, NaN), slices, in place."""
    from irr_b200 import ops
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 5, 7, 9, generator=g) * torch.logspace(-42, 30, 2 * 5 * 7 * 9).view(2, 5, 7, 9)
    x.view(-1)[:8] = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), 1.00390625, 1.01171875, 3.3895314e38, 1e-45])
    want = x.bfloat16().float()
    got = ops.round_bf16(x.to(cuda))
    assert torch.equal(got.cpu(), want)
    nan = ops.round_bf16(torch.full((1, 1, 1, 4), float("nan"), device=cuda))
    assert torch.isnan(nan).all()
    buf = torch.zeros(2, 8, 7, 9, device=cuda)
    buf[:, 2:7] = x.to(cuda)
    ops.round_bf16(buf[:, 2:7], out=buf[:, 2:7])  # in place on a channel slice
    assert torch.equal(buf[:, 2:7].cpu(), want) and float(buf[:, :2].abs().max()) == 0.0 and float(buf[:, 7:].abs().max()) == 0.0


def test_refine_pieces(cuda):
    from irr_b200 import ops
    fl = torch.from_numpy(rs(61, (2, 2, 14, 22)))
    ref = fl - fl.mean(2).mean(2)[:, :, None, None]
    assert (ops.sub_spatial_mean(fl.to(cuda)).cpu() - ref).abs().max().item() <= 1e-6
    d = torch.from_numpy(rs(62, (2, 3, 14, 22)))
    assert (ops.channel_l2norm(d.to(cuda)).cpu() - torch.norm(d, p=2, dim=1, keepdim=True)).abs().max().item() <= 1e-6
    logits = torch.from_numpy(rs(63, (2, 9, 14, 22))) * 2
    ref = O._kernel_gather(fl, logits)
    assert (ops.refine_gather(logits.to(cuda), fl.to(cuda)).cpu() - ref).abs().max().item() <= 1e-5


# ------------------------------------------------------------------ tcgen05 conv path
TC_CASES = [
    # (B, Cin, H, W, Cout, k, stride, dil)
    (1, 32, 16, 32, 32, 1, 1, 1),      # smallest: one K block, N = 32
    (1, 32, 16, 32, 32, 3, 1, 1),
    (2, 115, 28, 64, 128, 3, 1, 1),    # dense block conv1 (ragged channel tail 115 = 3*32 + 19)
    (1, 565, 14, 32, 128, 3, 1, 1),    # context conv0
    (1, 128, 28, 64, 96, 3, 1, 8), (1, 96, 28, 64, 64, 3, 1, 16),
    (2, 16, 47, 78, 32, 3, 2, 1),      # feature-extractor style stride 2, odd sizes
    (1, 128, 14, 32, 196, 3, 2, 1),    # two N tiles (196 -> 2 x 112)
    (2, 196, 7, 16, 32, 1, 1, 1), (1, 35, 13, 39, 128, 3, 1, 1), (1, 243, 24, 39, 128, 3, 1, 1),
    # thin layers (N padded to 16 / K tail padded to 8) and >1 work item per CTA (persistent loop, resident weights)
    (1, 563, 14, 32, 2, 3, 1, 1), (2, 32, 61, 97, 1, 3, 1, 1), (2, 11, 47, 155, 32, 3, 1, 1), (2, 3, 64, 96, 16, 3, 2, 1),
    (1, 32, 21, 37, 9, 3, 1, 1), (1, 16, 47, 156, 3, 1, 1, 1), (40, 32, 64, 96, 32, 3, 1, 1), (24, 64, 64, 96, 64, 3, 1, 1),
    (12, 160, 64, 96, 128, 3, 1, 1),
]


# extra shapes for the TMA-staged 3xF16 kernel: every half geometry (RW = 16..128), row-split dilations, ragged
# tiles in x and y, partial last row pair, wide rows (several x tiles), >1 item per CTA with streamed weights
H16_CASES = TC_CASES + [
    (1, 64, 109, 256, 64, 3, 1, 1), (2, 128, 30, 256, 128, 3, 1, 2), (1, 40, 20, 512, 32, 3, 1, 1),
    (1, 32, 12, 16, 16, 3, 1, 1), (1, 48, 9, 8, 32, 3, 1, 1), (2, 128, 28, 64, 128, 3, 1, 4), (1, 96, 33, 132, 96, 3, 1, 2),
    (1, 33, 7, 260, 48, 1, 1, 1), (6, 243, 55, 128, 128, 3, 1, 1), (3, 64, 50, 1024, 32, 3, 1, 1),
    # row-rolling kernel (Cin <= 32, Cout <= 64, W >= 96): partial strips, short/long segments, one- and two-step K
    (2, 32, 37, 256, 32, 3, 1, 1), (1, 11, 50, 512, 32, 3, 1, 1), (2, 32, 20, 132, 1, 3, 1, 1), (1, 16, 33, 128, 16, 3, 1, 1),
    (1, 32, 70, 1024, 64, 3, 1, 1), (16, 32, 109, 256, 9, 3, 1, 1), (5, 3, 45, 100, 40, 3, 1, 1),
]


@pytest.mark.parametrize("case", H16_CASES)
def test_conv2d_3xf16_vs_torch_cpu(cuda, case):
    """tcgen05 kind::f16 hi/lo-split conv (TMA-staged and gather variants) against an fp64 CPU conv."""
    from irr_b200 import ops
    B, Cin, H, W, Cout, k, s, d = case
    assert ops.tc_supported(Cout, Cin, k, s, d, ops.MATH_TC_3XF16)
    torch.manual_seed(41)
    x = torch.randn(B, Cin, H, W)
    x[0, 0, 0, :4] = torch.tensor([3.0e-6, -2.0e-5, 900.0, -4.0e3])[: min(4, W)]   # fp16-subnormal / large magnitudes
    w = torch.from_numpy(rs(42, (Cout, Cin, k, k))) * float(np.sqrt(2.0 / (Cin * k * k)))
    b = torch.from_numpy(rs(43, (Cout,))) * 0.1
    ref = torch.nn.functional.leaky_relu(
        torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=s, padding=((k - 1) * d) // 2, dilation=d), 0.1)
    packed = ops.pack_weights(w.to(cuda), ops.MATH_TC_3XF16)
    got = ops.conv2d(x.to(cuda), packed, b.to(cuda), Cout, k, s, d, slope=0.1, math=ops.MATH_TC_3XF16)
    err = (got.cpu().double() - ref).abs().max().item()
    refmax = ref.abs().max().item()
    print(f"[h16] {case}: max-abs {err:.3e} (|ref|max {refmax:.2f})")
    assert err <= 5e-5 * max(2.0, refmax)


def test_conv2d_3xf16_rolling_slices_addend(cuda):
    """Row-rolling kernel writing into / reading from channel slices with the residual epilogue (occlusion up-sampler)."""
    from irr_b200 import ops
    B, Ct, H, W = 3, 50, 45, 200
    buf = torch.from_numpy(rs(61, (B, Ct, H, W))).to(cuda)
    w = torch.from_numpy(rs(62, (32, 32, 3, 3))) * 0.06
    b = torch.from_numpy(rs(63, (32,))) * 0.1
    add = torch.from_numpy(rs(64, (B, 40, H, W))).to(cuda)
    ref = add[:, 4:36].cpu() + 0.1 * torch.nn.functional.leaky_relu(
        torch.nn.functional.conv2d(buf[:, 10:42].cpu(), w, b, padding=1), 0.1)
    out = torch.zeros((B, 44, H, W), device=cuda)
    ops.conv2d(buf[:, 10:42], ops.pack_weights(w.to(cuda), ops.MATH_TC_3XF16), b.to(cuda), 32, 3, slope=0.1,
               out=out[:, 7:39], addend=add[:, 4:36], alpha=0.1, math=ops.MATH_TC_3XF16)
    assert (out[:, 7:39].cpu() - ref).abs().max().item() <= 1e-4
    assert (out[:, :7] == 0).all() and (out[:, 39:] == 0).all()


@pytest.mark.parametrize("shape", [(2, 531, 20, 64, 2), (1, 530, 7, 16, 1), (2, 115, 33, 39, 2)])
def test_conv2d_3xf16_dual_output(cuda, shape):
    """Fused conv5 + conv_last tail (irr_conv2d_fwd_dual) against two separate fp64 convs; includes a split-K shape and a
    gather-variant shape."""
    from irr_b200 import ops
    B, Cin, H, W, co = shape
    x = torch.from_numpy(rs(71, (B, Cin, H, W)))
    w = torch.from_numpy(rs(72, (32 + co, Cin, 3, 3))) * float(np.sqrt(2.0 / (Cin * 9)))
    b = torch.from_numpy(rs(73, (32 + co,))) * 0.1
    add2 = torch.from_numpy(rs(74, (B, co, H, W)))
    full = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1)
    ref1 = torch.nn.functional.leaky_relu(full[:, :32], 0.1)
    ref2 = full[:, 32:] + add2.double()
    out = torch.zeros((B, 40, H, W), device=cuda)
    out2 = torch.empty((B, co, H, W), device=cuda)
    ops.conv2d_dual(x.to(cuda), ops.pack_weights(w.to(cuda), ops.MATH_TC_3XF16), b.to(cuda), 32 + co, 32, 3,
                    out=out[:, 3:35], out2=out2, slope=0.1, slope2=1.0, addend2=add2.to(cuda))
    assert (out[:, 3:35].cpu().double() - ref1).abs().max().item() <= 1e-4
    assert (out2.cpu().double() - ref2).abs().max().item() <= 1e-4
    assert (out[:, :3] == 0).all() and (out[:, 35:] == 0).all()


@pytest.mark.parametrize("shape", [(2, 467, 20, 64, 2), (1, 466, 7, 16, 1), (2, 147, 33, 39, 2), (16, 467, 28, 64, 2)])
def test_conv2d_3xf16_multi_segment(cuda, shape):
    """irr_conv2d_fwd_multi: three output segments with their own destination / activation / residual, one of them with
    the addend BEFORE the activation (a partial sum of the same layer) — against fp64 convs.  Shapes: TMA-staged, split-K
    (coarse level), gather variant (odd width), and a many-item launch."""
    from irr_b200 import ops
    B, Cin, H, W, co = shape
    x = torch.from_numpy(rs(81, (B, Cin, H, W)))
    w = torch.from_numpy(rs(82, (96 + co, Cin, 3, 3))) * float(np.sqrt(2.0 / (Cin * 9)))
    b = torch.from_numpy(rs(83, (96 + co,))) * 0.1
    pre = torch.from_numpy(rs(84, (B, 64, H, W)))
    add3 = torch.from_numpy(rs(85, (B, co, H, W)))
    full = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1)
    ref1 = 0.5 * torch.nn.functional.leaky_relu(full[:, :64] + pre.double(), 0.1)     # pre-activation addend, alpha
    ref2 = full[:, 64:96]                                                             # raw partial sums
    ref3 = full[:, 96:] + add3.double()                                               # post-activation addend
    o1 = torch.zeros((B, 70, H, W), device=cuda)
    o2 = torch.empty((B, 32, H, W), device=cuda)
    o3 = add3.to(cuda).clone()                                                        # in place: y == addend
    ops.conv2d_multi(x.to(cuda), ops.pack_weights(w.to(cuda), ops.MATH_TC_3XF16), b.to(cuda), 96 + co, 3, [
        dict(n_begin=0, out=o1[:, 3:67], slope=0.1, alpha=0.5, addend=pre.to(cuda), pre=True),
        dict(n_begin=64, out=o2, slope=1.0),
        dict(n_begin=96, out=o3, slope=1.0, addend=o3)])
    # same gate as test_conv2d_3xf16_vs_torch_cpu: the residual is the tensor core's fp32 accumulator, relative ~2e-5
    tol = lambda ref: 5e-5 * max(2.0, ref.abs().max().item())
    assert (o1[:, 3:67].cpu().double() - ref1).abs().max().item() <= tol(full[:, :64])
    assert (o2.cpu().double() - ref2).abs().max().item() <= tol(ref2)
    assert (o3.cpu().double() - ref3).abs().max().item() <= tol(ref3)
    assert (o1[:, :3] == 0).all() and (o1[:, 67:] == 0).all()


@pytest.mark.parametrize("hw", [(20, 64), (7, 16), (33, 39)])
@pytest.mark.parametrize("kind", ["flow", "occ"])
def test_dense_estimator_fused_tail_vs_fp64(cuda, kind, hw):
    """FlowEstimatorDense / OccEstimatorDense with the fused conv4|conv5|conv_last tail (pwc_modules._DenseEstimator):
    every dense output and conv_last (+ skip) against an fp64 restatement of models/pwc_modules.py:153-170,190-207."""
    from irr_b200 import ops, pwc_modules
    import torch.nn.functional as F
    H, W = hw
    est = (pwc_modules.FlowEstimatorDense(115) if kind == "flow" else pwc_modules.OccEstimatorDense(114)).to(cuda).eval()
    torch.manual_seed(5)
    for prm in est.parameters():
        if prm.dim() == 1:
            prm.data.normal_(0, 0.05)
    co = est.ch_out
    x = torch.from_numpy(rs(91, (2, est.ch_in, H, W))) * 0.5
    skip = torch.from_numpy(rs(92, (2, co, H, W)))
    cur = x.double()
    for c in (est.conv1, est.conv2, est.conv3, est.conv4, est.conv5):
        y = F.leaky_relu(F.conv2d(cur, c[0].weight.detach().cpu().double(), c[0].bias.detach().cpu().double(), padding=1), 0.1)
        cur = torch.cat([y, cur], 1)
    ref_last = F.conv2d(cur, est.conv_last[0].weight.detach().cpu().double(), est.conv_last[0].bias.detach().cpu().double(),
                        padding=1) + skip.double()
    pwc_modules.set_conv_math(ops.MATH_TC_3XF16)
    buf = torch.empty((2, est.total_ch, H, W), device=cuda)
    buf[:, 448:] = x.to(cuda)
    with torch.no_grad():
        out = est.forward_into(buf, addend=skip.to(cuda))
    scale = max(1.0, cur.abs().max().item())
    assert (buf.cpu().double() - cur).abs().max().item() <= 1e-4 * scale
    assert (out.cpu().double() - ref_last).abs().max().item() <= 1e-4 * max(1.0, ref_last.abs().max().item())


def test_conv2d_3xf16_slices_addend(cuda):
    from irr_b200 import ops
    B, Ct, H, W = 2, 100, 20, 36
    buf = torch.from_numpy(rs(51, (B, Ct, H, W))).to(cuda)
    w = torch.from_numpy(rs(52, (48, 72, 3, 3))) * 0.05
    b = torch.from_numpy(rs(53, (48,))) * 0.1
    add = torch.from_numpy(rs(54, (B, 48, H, W))).to(cuda)
    ref = add.cpu() + 0.1 * torch.nn.functional.conv2d(buf[:, 28:100].cpu(), w, b, padding=1)
    out = torch.zeros((B, 60, H, W), device=cuda)
    ops.conv2d(buf[:, 28:100], ops.pack_weights(w.to(cuda), ops.MATH_TC_3XF16), b.to(cuda), 48, 3, slope=1.0,
               out=out[:, 5:53], addend=add, alpha=0.1, math=ops.MATH_TC_3XF16)
    assert (out[:, 5:53].cpu() - ref).abs().max().item() <= 1e-4
    assert (out[:, :5] == 0).all() and (out[:, 53:] == 0).all()


@pytest.mark.parametrize("case", TC_CASES)
@pytest.mark.parametrize("mode", ["3xtf32", "tf32"])
def test_conv2d_tcgen05_vs_torch_cpu(cuda, case, mode):
    from irr_b200 import ops
    B, Cin, H, W, Cout, k, s, d = case
    math = ops.MATH_TC_3XTF32 if mode == "3xtf32" else ops.MATH_TC_TF32
    assert ops.tc_supported(Cout, Cin, k, s, d)
    torch.manual_seed(41)
    x = torch.randn(B, Cin, H, W)
    w = torch.from_numpy(rs(42, (Cout, Cin, k, k))) * float(np.sqrt(2.0 / (Cin * k * k)))
    b = torch.from_numpy(rs(43, (Cout,))) * 0.1
    ref = torch.nn.functional.leaky_relu(
        torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=s, padding=((k - 1) * d) // 2, dilation=d), 0.1)
    packed = ops.pack_weights(w.to(cuda), math)
    got = ops.conv2d(x.to(cuda), packed, b.to(cuda), Cout, k, s, d, slope=0.1, math=math)
    err = (got.cpu().double() - ref).abs().max().item()
    refmax = ref.abs().max().item()
    print(f"[tc] {case} {mode}: max-abs {err:.3e} (|ref|max {refmax:.2f})")
    # 3xTF32: products are fp32-exact; what remains is the tensor core's fp32 accumulator over K/8 steps.  Gate: 1e-4
    # abs for O(1) outputs (north_star), scaled with the output magnitude beyond that (N(0,1) inputs reach |y| ~ 6).
    assert err <= (5e-5 * max(2.0, refmax) if mode == "3xtf32" else 2e-2)


def test_conv2d_tcgen05_slices_addend(cuda):
    from irr_b200 import ops
    B, Ct, H, W = 2, 100, 20, 36
    buf = torch.from_numpy(rs(51, (B, Ct, H, W))).to(cuda)
    w = torch.from_numpy(rs(52, (48, 72, 3, 3))) * 0.05
    b = torch.from_numpy(rs(53, (48,))) * 0.1
    add = torch.from_numpy(rs(54, (B, 48, H, W))).to(cuda)
    ref = add.cpu() + 0.1 * torch.nn.functional.conv2d(buf[:, 28:100].cpu(), w, b, padding=1)
    out = torch.zeros((B, 60, H, W), device=cuda)
    ops.conv2d(buf[:, 28:100], ops.pack_weights(w.to(cuda), ops.MATH_TC_3XTF32), b.to(cuda), 48, 3, slope=1.0,
               out=out[:, 5:53], addend=add, alpha=0.1, math=ops.MATH_TC_3XTF32)
    assert (out[:, 5:53].cpu() - ref).abs().max().item() <= 1e-4
    assert (out[:, :5] == 0).all() and (out[:, 53:] == 0).all()
