"""Model-level parity on the GPU: teacher-forced stage tests (<= 1e-4, SURVEY.md §8(c) protocol), then end-to-end vs
the committed golden outputs of the reference and vs the oracle, reporting EPE / max-abs (the end-to-end forward is
chaotic at the 1e-4 level between ANY two conv implementations because of the hard warp mask — SURVEY.md F4/F5 — so the
end-to-end gate is EPE-level, the 1e-4 gate is per stage)."""
import numpy as np
import pytest
import torch

from oracle import irr_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4


def build(name, cuda, seed=1234, gain=0.7):
    import irr_b200
    p = O.synthetic_params(name, seed=seed, gain=gain)
    m = irr_b200.MODELS[name](None)
    irr_b200.load_state_dict_strict(m, p)
    return m.to(cuda).eval(), p


def cat2(rec, key, l, cuda):
    return torch.cat([rec[f"l{l}.{key}_f"], rec[f"l{l}.{key}_b"]], 0).to(cuda).contiguous()


@pytest.fixture(scope="module", params=["fp32", "3xtf32", "3xf16"], autouse=True)
def conv_math(request):
    """Every model-level test runs three times: CUDA-core fp32 convs, the tcgen05 3xTF32 convs and the tcgen05 3xF16
    convs with TMA-staged activations (both fp32-grade; 3xF16 is the bench default)."""
    from irr_b200 import ops, pwc_modules
    pwc_modules.set_conv_math({"fp32": ops.MATH_FP32_SIMT, "3xtf32": ops.MATH_TC_3XTF32,
                               "3xf16": ops.MATH_TC_3XF16}[request.param])
    yield request.param
    pwc_modules.set_conv_math(ops.MATH_TC_3XF16)


@pytest.fixture(scope="module", params=[(128, 192), (94, 156)])
def irr_case(request, cuda, conv_math):
    H, W = request.param
    m, p = build("IRR_PWC", cuda)
    i1, i2, gt = O.synthetic_pair(1, H, W, seed=7, max_flow=6.0)
    rec = {}
    with torch.no_grad():
        out = O.irr_pwc_forward(p, i1, i2, record=rec)
    return dict(m=m, p=p, i1=i1, i2=i2, gt=gt, rec=rec, out=out, H=H, W=W)


def maxdiff(a, b):
    return (a.detach().cpu() - b.detach().cpu()).abs().max().item()


@pytest.mark.parametrize("l", range(5))
def test_irr_estimate_stage_teacher_forced(irr_case, cuda, l):
    c = irr_case
    rec, m = c["rec"], c["m"]
    feat = torch.cat([rec[f"l{l}.x1"], rec[f"l{l}.x2"]], 0).to(cuda).contiguous()
    flow_up, occ_up = cat2(rec, "flow_up", l, cuda), cat2(rec, "occ_up", l, cuda)
    imgs = torch.cat([c["i1"], c["i2"]], 0).to(cuda)
    mine = {}
    m.estimator_level(l, feat, flow_up, occ_up, imgs, c["H"], c["W"], record=mine)
    assert maxdiff(mine["corr"], torch.cat([rec[f"l{l}.corr_f"], rec[f"l{l}.corr_b"]], 0)) <= TOL
    assert maxdiff(mine["x_1by1"], torch.cat([rec[f"l{l}.x1_1by1"], rec[f"l{l}.x2_1by1"]], 0)) <= TOL
    assert maxdiff(mine["flow_est"], torch.cat([rec[f"l{l}.flow_est_f"], rec[f"l{l}.flow_est_b"]], 0)) <= TOL
    assert maxdiff(mine["flow_cont"], torch.cat([rec[f"l{l}.flow_cont_f"], rec[f"l{l}.flow_cont_b"]], 0)) <= TOL
    assert maxdiff(mine["occ_cont"], torch.cat([rec[f"l{l}.occ_cont_f"], rec[f"l{l}.occ_cont_b"]], 0)) <= TOL


@pytest.mark.parametrize("l", range(5))
def test_irr_refine_stages_teacher_forced(irr_case, cuda, l):
    c = irr_case
    rec, m = c["rec"], c["m"]
    imgs = torch.cat([c["i1"], c["i2"]], 0).to(cuda)
    x1by1 = torch.cat([rec[f"l{l}.x1_1by1"], rec[f"l{l}.x2_1by1"]], 0).to(cuda).contiguous()
    flow_cont = cat2(rec, "flow_cont", l, cuda)  # local units, as produced at IRR_PWC.py:113-114
    flow = m.refine_flow_stage(flow_cont, x1by1, imgs, c["H"], c["W"])
    assert maxdiff(flow_cont, cat2(rec, "flow_cont_glob", l, cuda)) <= 1e-5  # F6: now in global units, in place
    assert maxdiff(flow, cat2(rec, "flow", l, cuda)) <= TOL
    occ = m.refine_occ_stage(cat2(rec, "occ_cont", l, cuda), x1by1, cat2(rec, "flow", l, cuda), c["H"], c["W"])
    assert maxdiff(occ, cat2(rec, "occ", l, cuda)) <= TOL


@pytest.mark.parametrize("l", [5, 6])
def test_irr_upsample_stage_teacher_forced(irr_case, cuda, l):
    c = irr_case
    rec, m = c["rec"], c["m"]
    feat = torch.cat([rec[f"l{l}.x1"], rec[f"l{l}.x2"]], 0).to(cuda).contiguous()
    occ = m.upsample_level(l, feat, cat2(rec, "flow_up", l, cuda), cat2(rec, "occ_in", l, cuda), c["H"], c["W"])
    assert maxdiff(occ, cat2(rec, "occ", l, cuda)) <= TOL


def test_irr_f6_aliasing_matters(irr_case, cuda):
    """SURVEY.md F6 regression: a 'clean' functional rescale (RefineFlow fed LOCAL-unit flow) must give a different
    answer than the as-executed dataflow we implement."""
    c = irr_case
    rec, m, l = c["rec"], c["m"], 3
    imgs = torch.cat([c["i1"], c["i2"]], 0).to(cuda)
    x1by1 = torch.cat([rec[f"l{l}.x1_1by1"], rec[f"l{l}.x2_1by1"]], 0).to(cuda).contiguous()
    ours = m.refine_flow_stage(cat2(rec, "flow_cont", l, cuda), x1by1, imgs, c["H"], c["W"])
    from irr_b200 import ops
    fc = cat2(rec, "flow_cont", l, cuda)
    clean_in = fc.clone()
    B2, _, h, w = fc.shape
    # the non-mutating variant: RefineFlow sees the LOCAL-unit tensor
    img_r = ops.resize_ac(imgs, h, w)
    from irr_b200.pwc_modules import flow_scales
    su, sv = flow_scales(h, w, 0.05, c["W"], c["H"], False)
    glob = ops.scale_channels(fc, s_even=su, s_odd=sv)
    diff = ops.warp(img_r, glob, c["H"], c["W"], 0.05, minuend=img_r, shift=B2 // 2)
    clean = m.refine_flow(clean_in, diff, x1by1)
    clean = ops.scale_channels(clean, s_even=su, s_odd=sv)
    assert maxdiff(ours, cat2(rec, "flow", l, cuda)) <= TOL
    assert maxdiff(clean, ours) > 1e-2


def _report(name, got, ref, gt=None):
    d = {k: maxdiff(got[k], ref[k]) for k in ref}
    e = O.epe(got["flow"].cpu(), ref["flow"].cpu()).item()
    from irr_b200 import pwc_modules
    msg = f"[parity] math={pwc_modules.get_conv_math()} {name}: max-abs {d}  EPE(new,ref)={e:.3e}"
    if gt is not None:
        msg += f"  EPE(new,GT)={O.epe(got['flow'].cpu(), gt).item():.4f} EPE(ref,GT)={O.epe(ref['flow'].cpu(), gt).item():.4f}"
    print(msg)
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_report.txt", "a") as f:
        f.write(msg + "\n")
    return d, e


K_FLOOR = 5.0       # end-to-end gates: <= K_FLOOR x what the reference differs from ITSELF by (GPU vs host)
ABS_FLOOR = 1e-3    # px / logit: lower bound of the floor so a lucky tiny floor does not make the gate impossible


def _self_floor(name, p, i1, i2, ref_cpu, cuda):
    """The reference's own noise floor for this (weights, input): the oracle's torch-op sequence run on THIS GPU (fp32,
    TF32 off) against the same sequence on the host.  Returns (EPE floor in px, mean |d occ logit| floor or None).
    The hard warp mask turns 1e-7 op-level differences into mask flips; every teacher-forced level is
    <= 1e-6 (test_pwc_classes_every_level_teacher_forced), so an end-to-end difference of the size of this floor is
    chaos, not arithmetic — and the gates below are multiples of it, not absolute numbers."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            g = O.FORWARDS[name]({k: v.to(cuda) for k, v in p.items()}, i1.to(cuda), i2.to(cuda))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    f = O.epe(g["flow"].cpu(), ref_cpu["flow"]).item()
    fo = (g["occ"].cpu() - ref_cpu["occ"]).abs().mean().item() if "occ" in ref_cpu else None
    return f, fo


def _gate(name, got, ref, floor, floor_occ):
    e = O.epe(got["flow"].cpu(), ref["flow"]).item()
    msg = f"[gate] {name}: EPE(new,oracle)={e:.2e} <= {K_FLOOR:g} x max(floor {floor:.2e}, {ABS_FLOOR:g})"
    ok = e <= K_FLOOR * max(floor, ABS_FLOOR)
    if floor_occ is not None:
        do = (got["occ"].cpu() - ref["occ"]).abs().mean().item()
        msg += f";  mean|d occ|={do:.2e} <= {K_FLOOR:g} x max(floor {floor_occ:.2e}, {ABS_FLOOR:g})"
        ok = ok and do <= K_FLOOR * max(floor_occ, ABS_FLOOR)
    _note(msg)
    assert ok, msg


def test_irr_end_to_end_vs_oracle_and_golden(irr_case, cuda, golden_dir):
    c = irr_case
    with torch.no_grad():
        got = c["m"]({"input1": c["i1"].to(cuda), "input2": c["i2"].to(cuda)})
    d, e = _report(f"IRR_PWC {c['H']}x{c['W']} vs oracle(CPU)", got, c["out"], c["gt"])
    g = np.load(f"{golden_dir}/models.npz")
    key = f"IRR_PWC_{c['H']}x{c['W']}"
    ref = {"flow": torch.from_numpy(g[key + "__flow"]), "occ": torch.from_numpy(g[key + "__occ"])}
    # the oracle itself must reproduce the reference's golden output bit for bit on this host or to rounding
    # the oracle on THIS host vs the golden made in the build container: equal up to the reference's own host-to-host
    # noise (different CPU vector width / thread count flips masks)
    d0, e0 = _report(f"IRR_PWC {c['H']}x{c['W']} oracle(this host) vs golden(reference)", c["out"], ref)
    d2, e2 = _report(f"IRR_PWC {c['H']}x{c['W']} vs golden(reference)", got, ref)
    floor, floor_occ = _self_floor("IRR_PWC", c["p"], c["i1"], c["i2"], c["out"], cuda)
    _gate(f"IRR_PWC {c['H']}x{c['W']}", got, c["out"], floor, floor_occ)        # flow AND occ, in multiples of the floor
    assert e0 <= K_FLOOR * max(floor, ABS_FLOOR) and e2 <= K_FLOOR * max(floor, e0, ABS_FLOOR)
    assert d["flow"] <= 1.0                   # chaotic tail bound (reference fp32-vs-fp64 is 0.16-1.6 px)


def test_irr_eval_prune_dead_is_output_identical(irr_case, cuda):
    """eval_prune_dead drops the backward occlusion chain (dead in eval mode, IRR_PWC.py:176-184): flow and occ must
    not change.  The pruned run launches fewer kernels."""
    from irr_b200 import ops
    c = irr_case
    m = c["m"]
    inp = {"input1": c["i1"].to(cuda), "input2": c["i2"].to(cuda)}
    with torch.no_grad():
        ops.LAUNCHES = 0
        full = {k: v.clone() for k, v in m(inp).items()}
        n_full = ops.LAUNCHES
        m.eval_prune_dead = True
        try:
            ops.LAUNCHES = 0
            pruned = m(inp)
            n_pruned = ops.LAUNCHES
        finally:
            m.eval_prune_dead = False
    assert n_pruned <= n_full
    # Bit-identical in the fp32 math mode.  In the tensor-core modes two runs of the SAME forward already differ in the
    # last ulp (rolling kernel: MMA issue order; split-K plans change with the batch) and the hard warp
    # mask amplifies that, so compare the way the parity tests do: EPE and mean |d occ|.
    df, do = maxdiff(pruned["flow"], full["flow"]), maxdiff(pruned["occ"], full["occ"])
    e = O.epe(pruned["flow"].cpu(), full["flow"].cpu()).item()
    mo = (pruned["occ"] - full["occ"]).abs().mean().item()
    print(f"[prune] {c['H']}x{c['W']}: launches {n_full} -> {n_pruned}, max-abs flow {df:.2e} occ {do:.2e}, EPE {e:.2e}")
    from irr_b200 import pwc_modules
    if pwc_modules.get_conv_math() == ops.MATH_FP32_SIMT:
        assert df == 0.0 and do == 0.0
    assert e <= 5e-3 and mo <= 5e-2   # (occ logits: ours-vs-oracle max-abs is 0.2-0.5 in the same tests)


@pytest.mark.parametrize("name", ["IRR_PWC", "PWCNet_irr_occ_bi"])
def test_full_size_sintel_shape_end_to_end(cuda, conv_math, name):
    """BASELINE configs 3 and 4 at their full image size (436 x 1024; batch 2 instead of 8 / 32 so the CPU oracle finishes
    in seconds — the batch only replicates the per-pair work): every pyramid level at the sizes the bench runs, against
    the oracle, plus the size-independent properties: batch rows are independent (row 0 of a batch-2 run == a batch-1
    run of the same pair, to the run-to-run noise) and finite everywhere."""
    if conv_math != "3xf16":
        pytest.skip("full-size run only in the default conv math")
    m, p = build(name, cuda)
    i1, i2, gt = O.synthetic_pair(2, 436, 1024, seed=3, max_flow=20.0)
    with torch.no_grad():
        ref = O.FORWARDS[name](p, i1, i2)
        got = m({"input1": i1.to(cuda), "input2": i2.to(cuda)})
        one = m({"input1": i1[:1].to(cuda), "input2": i2[:1].to(cuda)})
    assert got["flow"].shape == (2, 2, 436, 1024) and torch.isfinite(got["flow"]).all() and torch.isfinite(got["occ"]).all()
    d, e = _report(f"{name} 436x1024 b2 vs oracle(CPU)", got, ref, gt)
    floor, floor_occ = _self_floor(name, p, i1, i2, ref, cuda)
    _gate(f"{name} 436x1024 b2", got, ref, floor, floor_occ)
    assert O.epe(got["flow"][:1].cpu(), one["flow"].cpu()).item() <= 5e-3   # batch independence


@pytest.mark.parametrize("feat", ["fp32", "bf16"])
def test_irr_kitti_shape_end_to_end(cuda, conv_math, feat):
    """BASELINE config 5's shape (375 x 1242: level widths 621/311/156/78/39/20 — only two of them 16-byte aligned, so the
    cp.async correlation kernel, the gather conv variant and every ragged tail run), fp32 and with the feature pyramid
    rounded to bf16 ("mixed bf16 features").  The oracle gets the same rounding; tolerance as for config 3 — the bf16
    cast happens once, identically on both sides, it does not loosen the comparison."""
    if conv_math != "3xf16":
        pytest.skip("KITTI-shape run only in the default conv math")
    m, p = build("IRR_PWC", cuda)
    m.set_feature_dtype(feat)
    i1, i2, gt = O.synthetic_pair(1, 375, 1242, seed=5, max_flow=10.0)
    with torch.no_grad():
        ref = O.irr_pwc_forward(p, i1, i2, feature_bf16=(feat == "bf16"))
        got = m({"input1": i1.to(cuda), "input2": i2.to(cuda)})
    assert got["flow"].shape == (1, 2, 375, 1242) and got["occ"].shape == (1, 1, 375, 1242)
    d, e = _report(f"IRR_PWC 375x1242 features={feat} vs oracle(CPU)", got, ref, gt)
    floor, floor_occ = _self_floor("IRR_PWC", p, i1, i2, O.irr_pwc_forward(p, i1, i2), cuda)   # fp32 reference floor
    _gate(f"IRR_PWC 375x1242 features={feat}", got, ref, floor, floor_occ)
    assert d["flow"] <= 1.0
    if feat == "bf16":  # and the switch really changes the result
        with torch.no_grad():
            ref32 = O.irr_pwc_forward(p, i1, i2)
        assert O.epe(ref32["flow"], ref["flow"]).item() > 1e-4


@pytest.mark.parametrize("name,hw", [("PWCNet", (128, 128)), ("PWCNet_irr_occ_bi", (128, 192))])
def test_other_models_end_to_end(cuda, golden_dir, name, hw):
    H, W = hw
    m, p = build(name, cuda)
    i1, i2, gt = O.synthetic_pair(1, H, W, seed=7, max_flow=6.0)
    with torch.no_grad():
        ref = O.FORWARDS[name](p, i1, i2)
        got = m({"input1": i1.to(cuda), "input2": i2.to(cuda)})
    g = np.load(f"{golden_dir}/models.npz")
    gold = {"flow": torch.from_numpy(g[f"{name}_{H}x{W}__flow"])}
    _report(f"{name} {H}x{W} oracle(this host) vs golden(reference)", {"flow": ref["flow"]}, gold)
    d, e = _report(f"{name} {H}x{W} vs oracle(CPU)", got, ref, gt)
    floor, floor_occ = _self_floor(name, p, i1, i2, ref, cuda)
    _gate(f"{name} {H}x{W}", got, ref, floor, floor_occ)


@pytest.mark.parametrize("name", sorted(O.FAMILY))
def test_family_models_end_to_end(cuda, golden_dir, conv_math, name):
    """The six remaining PWC-family classes on the CUDA path vs the oracle and the reference-generated
    golden output, even and odd sizes (the odd one only against the oracle)."""
    if conv_math == "3xtf32":
        pytest.skip("family models run in the default (3xF16) and the fp32 conv math")
    m, p = build(name, cuda)
    g = np.load(f"{golden_dir}/family.npz")
    for (H, W, seed, mf) in [(64, 128, 11, 5.0), (94, 156, 12, 5.0)]:
        i1, i2, gt = O.synthetic_pair(1, H, W, seed=seed, max_flow=mf)
        with torch.no_grad():
            ref = O.FORWARDS[name](p, i1, i2)
            got = m({"input1": i1.to(cuda), "input2": i2.to(cuda)})
        assert set(got) == set(ref)
        d, e = _report(f"{name} {H}x{W} vs oracle(CPU)", got, ref, gt)
        floor, floor_occ = _self_floor(name, p, i1, i2, ref, cuda)
        _gate(f"{name} {H}x{W}", got, ref, floor, floor_occ)
        if (H, W) == (64, 128):
            gold = torch.from_numpy(g[f"{name}__flow"])
            e0 = O.epe(ref["flow"], gold).item()   # oracle on this host vs the golden made in the build container
            assert O.epe(got["flow"].cpu(), gold).item() <= K_FLOOR * max(floor, e0, ABS_FLOOR)


_ORACLE_RUNS = {}


def _oracle_levels(name, H, W):
    """The oracle's per-level inputs and outputs for one (class, size); cached across conv-math parametrisations."""
    key = (name, H, W)
    if key not in _ORACLE_RUNS:
        p = O.synthetic_params(name, seed=1234, gain=0.7)
        i1, i2, _ = O.synthetic_pair(1, H, W, seed=9, max_flow=4.0)
        rec = {}
        with torch.no_grad():
            out = O.FORWARDS[name](p, i1, i2, record=rec)
        _ORACLE_RUNS[key] = (p, i1, i2, rec, out)
    return _ORACLE_RUNS[key]


ALL_BUT_IRR_PWC = ["PWCNet", "PWCNet_irr_occ_bi"] + sorted(O.FAMILY)


@pytest.mark.parametrize("hw", [(128, 192), (94, 156)])
@pytest.mark.parametrize("name", ALL_BUT_IRR_PWC)
def test_pwc_classes_every_level_teacher_forced(cuda, conv_math, name, hw):
    """VERDICT r1 'next' #1: every pyramid level l = 0..4 of the eight classes that are not IRR_PWC, fed the ORACLE's
    level inputs (features, up-sampled flow / occlusion), must reproduce the oracle's level outputs to 1e-4 — so a real
    discrepancy in a level body cannot hide behind the mask chaos of the end-to-end comparison."""
    import irr_b200
    H, W = hw
    p, i1, i2, rec, _ = _oracle_levels(name, H, W)
    m = irr_b200.MODELS[name](None)
    irr_b200.load_state_dict_strict(m, p)
    m = m.to(cuda).eval()
    irr, bi, occ = {"PWCNet": (False, False, False), "PWCNet_irr_occ_bi": (True, True, True)}.get(name) or O.FAMILY[name]
    worst = {}

    def both(key, l):  # (2B or B, ...) tensor in the layout of the CUDA path
        if name == "PWCNet":
            return rec[f"l{l}.{key}"].to(cuda).contiguous()
        if bi:
            return torch.cat([rec[f"l{l}.{key}_f"], rec[f"l{l}.{key}_b"]], 0).to(cuda).contiguous()
        return rec[f"l{l}.{key}_f"].to(cuda).contiguous()

    def check(what, got, ref):
        d = maxdiff(got, ref)
        tol = TOL * max(1.0, ref.abs().max().item())
        worst[what] = max(worst.get(what, 0.0), d / max(1.0, ref.abs().max().item()))
        assert d <= tol, f"{name} {H}x{W} level {what}: max-abs {d:.3e} > {tol:.3e}"

    with torch.no_grad():
        for l in range(5):
            feat = torch.cat([rec[f"l{l}.x1"], rec[f"l{l}.x2"]], 0).to(cuda).contiguous()
            flow_up = both("flow_up", l)
            if name == "PWCNet":
                mine = {}
                flow = m.estimator_level(l, feat, flow_up, H, W, record=mine)
                check(f"l{l}.corr", mine["corr"], rec[f"l{l}.corr"])
                check(f"l{l}.flow", flow, rec[f"l{l}.flow"])
                continue
            occ_up = both("occ_up", l) if occ else None
            flow, occ_out = m.estimator_level(l, feat, flow_up, occ_up, H, W)
            check(f"l{l}.flow", flow, both("flow", l))
            if occ:
                check(f"l{l}.occ", occ_out, both("occ", l))
    _note(f"[teacher-forced] {name} {H}x{W} math={conv_math}: worst scaled max-abs per stage "
          + ", ".join(f"{k}={v:.1e}" for k, v in worst.items()))


def _note(msg):
    import os
    print(msg)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_report.txt", "a") as f:
        f.write(msg + "\n")


def test_module_api_matches_reference_modules(cuda, golden_dir):
    """Standalone module forwards (reference signatures) against KATs produced by the reference nn.Modules."""
    import irr_b200
    g = np.load(f"{golden_dir}/modules.npz")
    m, p = build("IRR_PWC", cuda, gain=1.0)
    from irr_b200 import pwc_modules

    # KAT inputs are N(0,1) noise pushed through up to 7 chained MSRA-gain convs: outputs reach |y| ~ 10.  The fp32
    # CUDA-core path is gated at 1e-4 absolute; for the tensor-core path the gate scales with the output magnitude
    # (its residual is the tensor core's fp32 accumulator, relative ~2e-5 — see test_conv2d_tcgen05_vs_torch_cpu).
    def maxdiff(a, b):  # noqa: F811  (relative-to-scale variant for this test)
        d = (a.detach().cpu() - b.detach().cpu()).abs().max().item()
        if pwc_modules.get_conv_math() != 0:
            d /= max(1.0, b.abs().max().item())
        return d

    def rs(seed, shape):
        return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype("float32")).to(cuda)

    pyr = m.feature_pyramid_extractor(rs(300, (1, 3, 64, 96)))
    for i, t in enumerate(pyr):
        assert maxdiff(t, torch.from_numpy(g[f"fpe__{i}"])) <= TOL
    x5, out = m.flow_estimators(rs(301, (1, 115, 12, 20)))
    assert maxdiff(x5[:, :64], torch.from_numpy(g["dense__x5"])) <= TOL and maxdiff(out, torch.from_numpy(g["dense__out"])) <= TOL
    assert maxdiff(m.context_networks(rs(302, (1, 565, 20, 36))), torch.from_numpy(g["ctx__out"])) <= TOL
    fl, d, f = rs(303, (2, 2, 14, 22)), rs(304, (2, 3, 14, 22)), rs(305, (2, 32, 14, 22))
    assert maxdiff(m.refine_flow(fl, d, f), torch.from_numpy(g["refine_flow__out"])) <= TOL
    oc, f2 = rs(306, (2, 1, 14, 22)), rs(307, (2, 32, 14, 22))
    assert maxdiff(m.refine_occ(oc, f, f2), torch.from_numpy(g["refine_occ__out"])) <= TOL
    assert maxdiff(m.occ_shuffle_upsample(oc, rs(308, (2, 10, 28, 44))), torch.from_numpy(g["occ_up__even"])) <= TOL
    assert maxdiff(m.occ_shuffle_upsample(oc, rs(309, (2, 10, 27, 43))), torch.from_numpy(g["occ_up__odd"])) <= TOL
    # Correlation module + compute_cost_volume share one kernel
    a, b = rs(1, (1, 16, 20, 24)), rs(2, (1, 16, 20, 24))
    c1 = irr_b200.Correlation(4, 1, 4, 1, 1, 1)(a, b)
    c2 = irr_b200.pwc_modules.compute_cost_volume(a, b, {"max_disp": 4})
    assert torch.equal(c1, c2) and maxdiff(c1, O.cost_volume(a.cpu(), b.cpu())) <= TOL
    # rescale_flow keeps the reference's in-place side effect
    t = rs(3, (1, 2, 8, 8)); t0 = t.clone()
    r = irr_b200.pwc_modules.rescale_flow(t, 0.05, 64, 64, to_local=True)
    assert torch.equal(r, t) and not torch.equal(t, t0)
