"""Row-pitched tensors (include/irr_b200.h, "ROW PITCH", ABI 2): widths that are not a multiple of 4 — KITTI's pyramid
levels 621 / 311 / 78 / 39, which the reference handles natively (models/irr_modules.py:21-27, pwc_modules.py:65-67) — are
stored with the next multiple of 4 as row pitch so that the TMA-fed conv and correlation kernels take them.  Every op of
the C-ABI that has a `*_pitch` argument is run on pitched operands whose pad columns hold NaN (a kernel that reads a pad
column poisons its result) and compared with the oracle / an fp64 CPU reference at the same tolerances as the dense tests."""
import numpy as np
import pytest
import torch

from oracle import irr_oracle as O

pytestmark = pytest.mark.gpu


def rs(seed, shape):
    return np.random.RandomState(seed).standard_normal(shape).astype("float32")


def pit(t, cuda, extra=0):
    """``t`` (CPU, NCHW) as the [..., :W] view of a NaN-filled CUDA buffer whose rows are padded to a multiple of 4
    (+ ``extra`` more groups of 4)."""
    t = torch.as_tensor(t)
    B, C, H, W = t.shape
    P = (W + 3) // 4 * 4 + 4 * extra
    base = torch.full((B, C, H, P), float("nan"), device=cuda)
    v = base[:, :, :, :W]
    v.copy_(t.to(cuda))
    return v


def nanbuf(B, C, H, W, cuda):
    P = (W + 3) // 4 * 4
    return torch.full((B, C, H, P), float("nan"), device=cuda)[:, :, :, :W]


PITCH_CONV = [
    # (B, Cin, H, W, Cout, k, stride, dil)    -> variant
    (1, 35, 13, 39, 128, 3, 1, 1),     # staged TMA, RW = 64, one x tile, split-K (few tiles)
    (2, 115, 24, 78, 128, 3, 1, 1),    # staged TMA, RW = 128
    (1, 243, 24, 39, 96, 3, 1, 1),
    (2, 128, 47, 155, 96, 3, 1, 8),    # dilated, two x tiles
    (1, 96, 20, 78, 64, 3, 1, 16),
    (2, 32, 37, 311, 32, 3, 1, 1),     # rolling kernel, 3 column strips, ragged last strip
    (1, 11, 50, 621, 32, 3, 1, 1),     # rolling kernel, KITTI level-5 width
    (2, 32, 20, 155, 1, 3, 1, 1),      # rolling kernel, N = 16 tile with one valid channel
    (2, 16, 47, 78, 32, 3, 2, 1),      # stride 2: gather variant reading a pitched input, pitched output (39 -> 40)
    (2, 3, 45, 621, 16, 3, 2, 1),      # first pyramid layer at KITTI width (output 311 -> 312)
    (2, 196, 6, 20, 32, 1, 1, 1),      # 1x1, width a multiple of 4 (pitch == W)
    (1, 33, 7, 261, 48, 1, 1, 1),      # 1x1, odd width
    (1, 565, 6, 39, 128, 3, 1, 1),     # split-K at a coarse odd level
]


@pytest.mark.parametrize("case", PITCH_CONV)
def test_conv2d_3xf16_pitched(cuda, case):
    from irr_b200 import ops
    B, Cin, H, W, Cout, k, s, d = case
    torch.manual_seed(141)
    x = torch.randn(B, Cin, H, W)
    w = torch.from_numpy(rs(142, (Cout, Cin, k, k))) * float(np.sqrt(2.0 / (Cin * k * k)))
    b = torch.from_numpy(rs(143, (Cout,))) * 0.1
    Ho, Wo = ops.conv_out_hw(H, W, k, s, d)
    add = torch.from_numpy(rs(144, (B, Cout, Ho, Wo)))
    ref = add.double() + 0.5 * torch.nn.functional.leaky_relu(
        torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=s, padding=((k - 1) * d) // 2, dilation=d), 0.1)
    packed = ops.pack_weights(w.to(cuda), ops.MATH_TC_3XF16)
    out = nanbuf(B, Cout + 5, Ho, Wo, cuda)
    ops.conv2d(pit(x, cuda), packed, b.to(cuda), Cout, k, s, d, slope=0.1, out=out[:, 2:2 + Cout], addend=pit(add, cuda),
               alpha=0.5, math=ops.MATH_TC_3XF16)
    got = out[:, 2:2 + Cout].cpu().double()
    assert torch.isfinite(got).all(), "a pad column was read (NaN in the result)"
    err = (got - ref).abs().max().item()
    assert err <= 5e-5 * max(2.0, ref.abs().max().item()), err
    assert torch.isnan(out[:, :2]).all() and torch.isnan(out[:, 2 + Cout:]).all()   # neighbours of the slice untouched
    # dense input, dense output: the same numbers up to the association of the K loop
    dense = ops.conv2d(x.to(cuda), packed, b.to(cuda), Cout, k, s, d, slope=0.1, out=torch.empty(B, Cout, Ho, Wo, device=cuda),
                       addend=add.to(cuda), alpha=0.5, math=ops.MATH_TC_3XF16)
    assert (dense.cpu().double() - got).abs().max().item() <= 2e-5 * max(2.0, ref.abs().max().item())


def test_conv2d_pitched_input_larger_pitch_and_other_math_refuses(cuda):
    """Any pitch >= W that is a multiple of 4 works (a view into a wider buffer); the fp32 SIMT / TF32 paths do not
    implement pitches and must say so instead of reading the wrong pixels."""
    from irr_b200 import ops
    x = torch.from_numpy(rs(150, (2, 40, 21, 37)))
    w = torch.from_numpy(rs(151, (24, 40, 3, 3))) * 0.05
    b = torch.from_numpy(rs(152, (24,))) * 0.1
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1), 0.1)
    out = nanbuf(2, 24, 21, 37, cuda)
    ops.conv2d(pit(x, cuda, extra=3), ops.pack_weights(w.to(cuda), ops.MATH_TC_3XF16), b.to(cuda), 24, 3, slope=0.1, out=out,
               math=ops.MATH_TC_3XF16)
    assert (out.cpu().double() - ref).abs().max().item() <= 1e-4
    with pytest.raises(RuntimeError):
        ops.conv2d(pit(x, cuda), ops.pack_weights(w.to(cuda)), b.to(cuda), 24, 3, slope=0.1,
                   out=torch.empty(2, 24, 21, 37, device=cuda))


@pytest.mark.parametrize("shape", [(2, 467, 20, 39, 2), (1, 466, 7, 21, 1), (16, 467, 12, 78, 2)])
def test_conv2d_multi_segment_pitched(cuda, shape):
    """The fused dense-estimator tail (output segments with their own destination / residual / activation, pre-activation
    partial sums) on pitched buffers."""
    from irr_b200 import ops
    B, Cin, H, W, nl = shape
    torch.manual_seed(161)
    x = torch.randn(B, Cin, H, W)
    C4, C5 = 64, 32
    Cout = C4 + C5 + 16
    w = torch.from_numpy(rs(162, (Cout, Cin, 3, 3))) * float(np.sqrt(2.0 / (Cin * 9)))
    bias = torch.from_numpy(rs(163, (Cout,))) * 0.1
    skip = torch.from_numpy(rs(164, (B, nl, H, W)))
    part = torch.from_numpy(rs(165, (B, C5, H, W)))
    full = torch.nn.functional.conv2d(x.double(), w.double(), bias.double(), padding=1)
    ref4 = torch.nn.functional.leaky_relu(full[:, :C4], 0.1)
    ref5 = torch.nn.functional.leaky_relu(full[:, C4:C4 + C5] + part.double(), 0.1)     # residual BEFORE the activation
    refl = full[:, C4 + C5:C4 + C5 + nl] + skip.double()
    o4, o5, ol = nanbuf(B, C4, H, W, cuda), nanbuf(B, C5, H, W, cuda), nanbuf(B, 16, H, W, cuda)
    skip16 = torch.zeros(B, 16, H, W)
    skip16[:, :nl] = skip
    ops.conv2d_multi(pit(x, cuda), ops.pack_weights(w.to(cuda), ops.MATH_TC_3XF16), bias.to(cuda), Cout, 3,
                     [dict(n_begin=0, out=o4, slope=0.1),
                      dict(n_begin=C4, out=o5, slope=0.1, addend=pit(part, cuda), pre=True),
                      dict(n_begin=C4 + C5, out=ol, slope=1.0, addend=pit(skip16, cuda))], math=ops.MATH_TC_3XF16)
    for got, ref in ((o4, ref4), (o5, ref5), (ol[:, :nl], refl)):
        g = got.cpu().double()
        assert torch.isfinite(g).all()
        assert (g - ref).abs().max().item() <= 5e-5 * max(2.0, ref.abs().max().item())


CORR_PITCH = [(1, 96, 24, 78), (2, 32, 47, 155), (2, 64, 24, 39), (3, 17, 9, 13), (16, 196, 6, 21), (2, 128, 12, 39)]


@pytest.mark.parametrize("shape", CORR_PITCH)
def test_correlation_pitched(cuda, shape):
    """Plain and fused (warp) cost volume, TMA kernels and the channel-split launches, on pitched f1 / f2 / flow / out."""
    from irr_b200 import ops
    B, C, H, W = shape
    him, wim = 16 * H, 16 * W
    f1, f2 = torch.from_numpy(rs(171, shape)), torch.from_numpy(rs(172, shape))
    flow = torch.from_numpy(rs(173, (B, 2, H, W))) * torch.tensor([wim / W, him / H]).view(1, 2, 1, 1) * 0.05 * 2.5
    sh = B // 2
    ref = torch.nn.functional.leaky_relu(O.cost_volume(f1, torch.roll(f2, -sh, 0)), 0.1)
    refw = torch.nn.functional.leaky_relu(O.cost_volume(f1, O.warp(torch.roll(f2, -sh, 0), flow, him, wim, 0.05)), 0.1)
    p1, p2, pf = pit(f1, cuda), pit(f2, cuda), pit(flow, cuda)
    out = nanbuf(B, 90, H, W, cuda)
    ops.correlation(p1, p2, out=out[:, 4:85], shift=sh, slope=0.1)
    got = out[:, 4:85].cpu()
    assert torch.isfinite(got).all() and (got - ref).abs().max().item() <= 1e-4
    assert torch.isnan(out[:, :4]).all() and torch.isnan(out[:, 85:]).all()
    outw = nanbuf(B, 81, H, W, cuda)
    ops.warp_correlation(p1, p2, pf, him, wim, 0.05, out=outw, shift=sh, slope=0.1)
    gotw = outw.cpu()
    assert torch.isfinite(gotw).all() and (gotw - refw).abs().max().item() <= 1e-4
    # the dense tensors through the same entry point: equal up to the kernel variant's summation order
    d = ops.correlation(f1.to(cuda), f2.to(cuda), shift=sh, slope=0.1)
    dw = ops.warp_correlation(f1.to(cuda), f2.to(cuda), flow.to(cuda), him, wim, 0.05, shift=sh, slope=0.1)
    assert (d.cpu() - got).abs().max().item() <= 2e-6 and (dw.cpu() - gotw).abs().max().item() <= 2e-6
    # determinism
    outw2 = nanbuf(B, 81, H, W, cuda)
    ops.warp_correlation(p1, p2, pf, him, wim, 0.05, out=outw2, shift=sh, slope=0.1)
    assert torch.equal(outw2, outw)


def test_warp_pitched_mask_bitexact(cuda, golden_dir):
    from irr_b200 import ops
    ops.set_grid_mode(ops.GRID_TRUE_DIV)
    g = np.load(f"{golden_dir}/warp.npz")
    for ci in range(5):
        seed, B, C, H, W, him, wim = [int(v) for v in g[f"case{ci}__meta"]]
        x = rs(seed, (B, C, H, W))
        flow = g[f"case{ci}__flow"]
        lx, ly = torch.from_numpy(g[f"case{ci}__lin_x"]).to(cuda), torch.from_numpy(g[f"case{ci}__lin_y"]).to(cuda)
        mask = torch.empty((B, H, W), device=cuda)
        out = nanbuf(B, C, H, W, cuda)
        ops.warp(pit(x, cuda), pit(flow, cuda), him, wim, 0.05, out=out, mask_out=mask, lin_x=lx, lin_y=ly)
        assert (mask.cpu().numpy() != g[f"case{ci}__mask"]).sum() == 0
        assert np.abs(out.cpu().numpy() - g[f"case{ci}__out"]).max() <= 2e-6
    x = torch.from_numpy(rs(181, (2, 7, 13, 39)))
    fl = torch.from_numpy(rs(182, (2, 2, 13, 39))) * 3.0
    ref = x - torch.roll(O.warp(x, fl, 375, 1242, 0.05), 0, 0)
    d = nanbuf(2, 7, 13, 39, cuda)
    ops.warp(pit(x, cuda), pit(fl, cuda), 375, 1242, 0.05, out=d, minuend=pit(x, cuda))
    assert (d.cpu() - ref).abs().max().item() <= 2e-6


def test_small_ops_pitched(cuda):
    """resize (align_corners=True), nearest x2 (+ the odd-size bilinear), spatial-mean subtraction, channel norm, the
    refinement gather, scale / round on pitched operands == the dense results, bit for bit (same kernels, same order)."""
    from irr_b200 import ops
    t = torch.from_numpy(rs(191, (2, 2, 13, 39)))
    for (oh, ow) in [(26, 78), (24, 77), (13, 39), (7, 21)]:
        a = ops.resize_ac(t.to(cuda), oh, ow, s_even=2.0, s_odd=3.0, pitched=False)
        o = nanbuf(2, 2, oh, ow, cuda)
        ops.resize_ac(pit(t, cuda), oh, ow, out=o, s_even=2.0, s_odd=3.0)
        assert torch.equal(o, a)
    o1 = torch.from_numpy(rs(192, (2, 1, 12, 39)))
    for (oh, ow) in [(24, 78), (23, 77), (24, 77)]:
        a = ops.upsample_nearest2x(o1.to(cuda), oh, ow, out=torch.empty(2, 1, oh, ow, device=cuda))
        o = nanbuf(2, 1, oh, ow, cuda)
        ops.upsample_nearest2x(pit(o1, cuda), oh, ow, out=o)
        assert torch.equal(o, a)
    fl = torch.from_numpy(rs(193, (2, 2, 14, 21)))
    ref = fl - fl.mean(2).mean(2)[:, :, None, None]
    o = nanbuf(2, 2, 14, 21, cuda)
    ops.sub_spatial_mean(pit(fl, cuda), out=o)
    assert (o.cpu() - ref).abs().max().item() <= 1e-6
    d = torch.from_numpy(rs(194, (2, 3, 14, 21)))
    o = nanbuf(2, 1, 14, 21, cuda)
    ops.channel_l2norm(pit(d, cuda), out=o)
    assert (o.cpu() - torch.norm(d, p=2, dim=1, keepdim=True)).abs().max().item() <= 1e-6
    logits = torch.from_numpy(rs(195, (2, 9, 14, 21))) * 2
    o = nanbuf(2, 2, 14, 21, cuda)
    ops.refine_gather(pit(logits, cuda), pit(fl, cuda), out=o)
    assert (o.cpu() - O._kernel_gather(fl, logits)).abs().max().item() <= 1e-5
    s = nanbuf(2, 2, 14, 21, cuda)
    ops.scale_channels(pit(fl, cuda), out=s, s_even=0.5, s_odd=-2.0)
    assert torch.equal(s[:, 0].cpu(), fl[:, 0] * 0.5) and torch.equal(s[:, 1].cpu(), fl[:, 1] * -2.0)
    # dense -> pitched and pitched -> dense copies
    p = ops.pitched(fl.to(cuda))
    assert p.stride(2) == 24 and torch.equal(p.cpu(), fl)
    back = ops.scale_channels(p, out=torch.empty(2, 2, 14, 21, device=cuda))
    assert torch.equal(back.cpu(), fl)
    r = nanbuf(2, 2, 14, 21, cuda)
    ops.round_bf16(pit(fl, cuda), out=r)
    assert torch.equal(r.cpu(), fl.bfloat16().float())


def test_mixed_pitch_is_refused(cuda):
    from irr_b200 import ops
    f = torch.from_numpy(rs(196, (1, 8, 9, 13)))
    with pytest.raises(RuntimeError):
        ops.correlation(pit(f, cuda), f.to(cuda))


# ------------------------------------------------------------------ bf16 STORAGE of the correlation inputs
@pytest.mark.parametrize("shape", [(2, 32, 47, 156), (1, 96, 24, 78), (2, 64, 24, 39), (16, 196, 6, 21), (2, 128, 12, 40),
                                   (3, 17, 9, 13), (2, 32, 109, 256)])
def test_correlation_bf16_storage(cuda, shape):
    """irr_warp_correlation_fwd_dt with dtype_in = bf16 (BASELINE configs[4] "mixed bf16 features"; SURVEY.md §8(b) sketch):
    f1 / f2 are read as packed bf16 (half the bytes), the arithmetic is the fp32 kernel's on the converted values — so the
    result must equal, BIT FOR BIT, the fp32-storage kernel run on the same bf16-rounded values; and both match the oracle.
    Plain, fused (warp) and the channel-split launches; odd widths (row pitch 8 for bf16, 4 for fp32)."""
    from irr_b200 import ops
    B, C, H, W = shape
    him, wim = 16 * H, 16 * W
    f = torch.from_numpy(rs(201, shape))
    flow = torch.from_numpy(rs(203, (B, 2, H, W))) * torch.tensor([wim / W, him / H]).view(1, 2, 1, 1) * 0.05 * 2.5
    if B == 3:   # one tile with wildly divergent flow: the footprint window does not fit, taps come from global memory
        flow[0] = flow[0] * 20.0
    sh = B // 2
    fr = f.bfloat16().float()                       # the values both paths see
    x = pit(f, cuda)                                # fp32 layout, rounded in place by the store op
    x16 = ops.round_bf16_store(x, round_in_place=True)
    assert x16.dtype == torch.bfloat16 and x16.stride(2) % 8 == 0
    assert torch.equal(x.cpu(), fr) and torch.equal(x16.cpu().float(), fr)
    pf = pit(flow, cuda)
    a32 = nanbuf(B, 81, H, W, cuda); a16 = nanbuf(B, 81, H, W, cuda)
    ops.correlation(x, x, out=a32, shift=sh, slope=0.1)
    ops.correlation(x16, x16, out=a16, shift=sh, slope=0.1)
    assert torch.isfinite(a16).all() and torch.equal(a16, a32)
    w32 = nanbuf(B, 81, H, W, cuda); w16 = nanbuf(B, 81, H, W, cuda)
    ops.warp_correlation(x, x, pf, him, wim, 0.05, out=w32, shift=sh, slope=0.1)
    ops.warp_correlation(x16, x16, pf, him, wim, 0.05, out=w16, shift=sh, slope=0.1)
    assert torch.isfinite(w16).all() and torch.equal(w16, w32)
    if B * C * H * W <= 2_000_000:
        ref = torch.nn.functional.leaky_relu(O.cost_volume(fr, torch.roll(fr, -sh, 0)), 0.1)
        refw = torch.nn.functional.leaky_relu(O.cost_volume(fr, O.warp(torch.roll(fr, -sh, 0), flow, him, wim, 0.05)), 0.1)
        assert (a16.cpu() - ref).abs().max().item() <= 1e-4 and (w16.cpu() - refw).abs().max().item() <= 1e-4


def test_correlation_bf16_needs_aligned_rows(cuda):
    """No unaligned bf16 path: a dense bf16 tensor with an odd width is refused (IRR_E_ALIGN), not mis-read."""
    from irr_b200 import ops
    x = torch.randn(1, 8, 9, 13, device=cuda).bfloat16()
    with pytest.raises(RuntimeError):
        ops.correlation(x, x)


def test_irr_pwc_bf16_storage_is_output_identical(cuda):
    """IRR_PWC with bf16-valued features: cost volumes fed from the packed bf16 copy == fed from the fp32 layout."""
    import irr_b200
    from irr_b200 import pwc_modules, ops
    pwc_modules.set_conv_math(ops.MATH_TC_3XF16)
    p = O.synthetic_params("IRR_PWC", seed=1234, gain=0.7)
    m = irr_b200.IRR_PWC(None)
    irr_b200.load_state_dict_strict(m, p)
    m = m.to(cuda).eval().set_feature_dtype("bf16")
    i1, i2, _ = O.synthetic_pair(1, 94, 156, seed=9, max_flow=5.0)
    inp = {"input1": i1.to(cuda), "input2": i2.to(cuda)}
    m.bf16_storage = True
    a = m(inp)
    m.bf16_storage = False
    b = m(inp)
    assert torch.equal(a["flow"], b["flow"]) and torch.equal(a["occ"], b["occ"])


def test_round_bf16_store_special_values(cuda):
    """irr_round_bf16_store_fwd == x.bfloat16() bit for bit (RNE, ties, subnormals, inf, NaN quieted), packed copy and the
    optional in-place fp32 rounding; without round_in_place the source stays untouched."""
    from irr_b200 import ops
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 5, 7, 9, generator=g) * torch.logspace(-42, 30, 2 * 5 * 7 * 9).view(2, 5, 7, 9)
    x.view(-1)[:8] = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), 1.00390625, 1.01171875, 3.3895314e38, 1e-45])
    want = x.bfloat16()
    src = pit(x, cuda)
    y16 = ops.round_bf16_store(src, round_in_place=False)
    assert torch.equal(y16.cpu().view(torch.int16), want.view(torch.int16)) and torch.equal(src.cpu(), x)
    y16b = ops.round_bf16_store(src, round_in_place=True)
    assert torch.equal(y16b.cpu().view(torch.int16), want.view(torch.int16)) and torch.equal(src.cpu(), want.float())
    n16 = ops.round_bf16_store(torch.full((1, 1, 2, 8), float("nan"), device=cuda))
    assert torch.isnan(n16.float()).all()


# ------------------------------------------------------------------ direct thin-layer conv kernel (fp32 path)
@pytest.mark.parametrize("case", [(2, 3, 45, 621, 16, 3, 2), (1, 3, 64, 96, 16, 3, 2), (2, 3, 33, 50, 9, 3, 1),
                                  (2, 16, 47, 311, 3, 1, 1), (1, 16, 20, 64, 3, 1, 1)])
def test_conv2d_direct_thin_layers(cuda, case):
    """3 -> 16 (k 3, stride 2: first pyramid layer, pwc_modules.py:95-104) and 16 -> 3 1x1 (IRR_PWC.py:44-45) on the direct
    kernel of the fp32 path: fp32 FMAs, dense and row-pitched operands (NaN pad columns), residual epilogue; and the module
    wrapper routes these layers there in the default 3xF16 mode."""
    from irr_b200 import ops, pwc_modules
    B, Cin, H, W, Cout, k, s = case
    assert ops.direct_supported(Cout, Cin, k)
    torch.manual_seed(211)
    x = torch.randn(B, Cin, H, W)
    w = torch.from_numpy(rs(212, (Cout, Cin, k, k))) * float(np.sqrt(2.0 / (Cin * k * k)))
    b = torch.from_numpy(rs(213, (Cout,))) * 0.1
    Ho, Wo = ops.conv_out_hw(H, W, k, s, 1)
    add = torch.from_numpy(rs(214, (B, Cout, Ho, Wo)))
    ref = add.double() + 0.5 * torch.nn.functional.leaky_relu(
        torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=s, padding=(k - 1) // 2), 0.1)
    packed = ops.pack_weights(w.to(cuda), ops.MATH_FP32_SIMT)
    out = nanbuf(B, Cout + 3, Ho, Wo, cuda)
    ops.conv2d(pit(x, cuda), packed, b.to(cuda), Cout, k, s, 1, slope=0.1, out=out[:, 1:1 + Cout], addend=pit(add, cuda),
               alpha=0.5, math=ops.MATH_FP32_SIMT)
    got = out[:, 1:1 + Cout].cpu().double()
    assert torch.isfinite(got).all() and (got - ref).abs().max().item() <= 1e-5
    assert torch.isnan(out[:, :1]).all() and torch.isnan(out[:, 1 + Cout:]).all()
    dense = ops.conv2d(x.to(cuda), packed, b.to(cuda), Cout, k, s, 1, slope=0.1, out=torch.empty(B, Cout, Ho, Wo, device=cuda),
                       addend=add.to(cuda), alpha=0.5, math=ops.MATH_FP32_SIMT)
    assert torch.equal(dense.cpu().double(), got)
    pwc_modules.set_conv_math(ops.MATH_TC_3XF16)
    blk = pwc_modules.conv(Cin, Cout, kernel_size=k, stride=s).to(cuda)
    assert blk._math() == ops.MATH_FP32_SIMT
    with torch.no_grad():
        blk[0].weight.copy_(w.to(cuda)); blk[0].bias.copy_(b.to(cuda))
        y = blk(x.to(cuda))
    refy = torch.nn.functional.leaky_relu(
        torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=s, padding=(k - 1) // 2), 0.1)
    assert (y.cpu().double() - refy).abs().max().item() <= 1e-5
