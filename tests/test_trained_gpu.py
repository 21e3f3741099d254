"""Trained-checkpoint parity on the GPU against the UNMODIFIED reference classes, with the reference's own
same-GPU noise floor measured beside it (VERDICT r1 'next' #2, #9).

Needs ``baseline/_ref`` (``python scripts/stage_ref.py`` in the build container: the reference's ``models/*.py`` and
four ``checkpoint_best.ckpt`` files; git-ignored, shipped to the GPU box by gpurun).  Skipped when it is not staged.

For every checkpoint, on a smooth synthetic pair with known flow:
  R0   reference on this GPU, fp32, TF32 off, cudnn.benchmark off
  R1   the same with cudnn.benchmark = True (the reference's own setting, main.py:73)
  Rtf  the same at torch's stock setting (cudnn.allow_tf32 = True)
  Rcpu reference on the host CPU (``.cuda()`` shimmed to a no-op)
  N    irr_b200 (3xF16 tensor-core convs, grid mode = torch-CUDA arithmetic so the hard mask matches R0 bit for bit)
floor = max(EPE(R0,R1), EPE(R0,Rcpu)): what the reference differs from ITSELF by.  The gates are "<= K x floor" (with an
absolute 1e-3 px lower bound so a lucky zero floor does not make the gate impossible); all numbers go to
gpurun_out/parity_trained.txt (committed as profiles/r02_parity_trained.txt).
"""
import contextlib
import importlib
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
CK = os.path.join(REF, "saved_check_point", "pwcnet")
K_FLOOR = 3.0        # gate: EPE(new, ref) <= K_FLOOR * max(floor, ABS_FLOOR)
ABS_FLOOR = 1e-3     # px

CASES = [  # (checkpoint dir, class name, H, W, seed, max_flow)
    ("IRR-PWC_sintel", "IRR_PWC", 436, 1024, 3, 20.0),
    ("IRR-PWC_kitti", "IRR_PWC", 375, 1242, 5, 20.0),
    ("PWCNet", "PWCNet", 256, 256, 2, 12.0),
    ("PWCNet-irr", "PWCNet_irr", 256, 256, 2, 12.0),
    ("PWCNet", "PWCNet", 436, 1024, 3, 20.0),
]


def _note(msg):
    print(msg)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_trained.txt"), "a") as f:
        f.write(msg + "\n")


@pytest.fixture(scope="module")
def refmodels():
    if not os.path.isdir(os.path.join(REF, "models")):
        pytest.skip("baseline/_ref not staged (python scripts/stage_ref.py)")
    sys.path.insert(0, REF)
    try:
        models = importlib.import_module("models")
    finally:
        sys.path.remove(REF)
    assert os.path.abspath(models.__file__).startswith(REF)
    return models


def _state(ckpt):
    import irr_b200
    path = os.path.join(CK, ckpt, "checkpoint_best.ckpt")
    if not os.path.exists(path):
        pytest.skip(f"{path} not staged (python scripts/stage_ref.py)")
    sd = torch.load(path, map_location="cpu", weights_only=True)["state_dict"]
    return irr_b200.checkpoint.strip_prefix(sd)


@contextlib.contextmanager
def _backend(benchmark=False, tf32=False):
    old = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.benchmark = benchmark
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    try:
        yield
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@contextlib.contextmanager
def _cuda_shim():
    """The reference hard-codes .cuda() (pwc_modules.py:111,129): a no-op shim runs it on the host."""
    old = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = old


def _ref_forward(models, cls, sd, i1, i2, dev):
    m = getattr(models, cls)(None)
    m.load_state_dict(sd)
    m = m.to(dev).eval()
    with torch.no_grad():
        out = m({"input1": i1.to(dev), "input2": i2.to(dev)})
    return {k: v.float().cpu() for k, v in out.items() if torch.is_tensor(v)}


def _epe(a, b):
    return torch.norm(a - b, p=2, dim=1).mean().item()


def _occ_stats(a, b):
    """mean |d logit|, agreement of round(sigmoid(.)) in percent (losses.py:693 thresholds the sigmoid at 0.5)."""
    return (a - b).abs().mean().item(), 100.0 * ((a > 0) == (b > 0)).float().mean().item()


def _mask_agreement(fa, fb, cuda, div_flow=0.05):
    """Percent of full-resolution pixels on which the hard warp mask of the two final flows agrees (a proxy for the
    per-level masks, computed by the irr_b200 warp kernel on both flows)."""
    from irr_b200 import ops
    B, _, H, W = fa.shape
    ms = []
    for f in (fa, fb):
        fl = (f * div_flow).to(cuda).contiguous()
        mask = torch.empty((B, 1, H, W), dtype=torch.float32, device=cuda)
        ops.warp(torch.ones((B, 1, H, W), device=cuda), fl, H, W, div_flow, mask_out=mask)
        ms.append(mask.cpu())
    return 100.0 * (ms[0] == ms[1]).float().mean().item()


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}-{c[2]}x{c[3]}" for c in CASES])
def test_trained_checkpoint_parity_and_noise_floor(cuda, refmodels, case):
    import irr_b200
    from irr_b200 import ops, pwc_modules
    from irr_b200.synthetic import synthetic_pair
    ckpt, cls, H, W, seed, mf = case
    sd = _state(ckpt)
    i1, i2, gt = synthetic_pair(1, H, W, seed=seed, max_flow=mf)
    with _backend(False, False):
        R0 = _ref_forward(refmodels, cls, sd, i1, i2, cuda)
        R0b = _ref_forward(refmodels, cls, sd, i1, i2, cuda)
    with _backend(True, False):
        R1 = _ref_forward(refmodels, cls, sd, i1, i2, cuda)
    with _backend(False, True):
        Rtf = _ref_forward(refmodels, cls, sd, i1, i2, cuda)
    with _cuda_shim():
        Rcpu = _ref_forward(refmodels, cls, sd, i1, i2, torch.device("cpu"))

    pwc_modules.set_conv_math(ops.MATH_TC_3XF16)
    ops.set_grid_mode(ops.GRID_RECIP_MUL)   # the reference ON THE GPU divides by python scalars as a*(1/b)
    try:
        m = irr_b200.MODELS[cls](None)
        irr_b200.load_state_dict_strict(m, sd)
        m = m.to(cuda).eval()
        with torch.no_grad():
            N = {k: v.cpu() for k, v in m({"input1": i1.to(cuda), "input2": i2.to(cuda)}).items()}
            N2 = {k: v.cpu() for k, v in m({"input1": i1.to(cuda), "input2": i2.to(cuda)}).items()}
    finally:
        ops.set_grid_mode(ops.GRID_TRUE_DIV)
    assert all(torch.isfinite(v).all() for v in N.values())

    f_self = _epe(R0["flow"], R0b["flow"])
    f_bench = _epe(R0["flow"], R1["flow"])
    f_cpu = _epe(R0["flow"], Rcpu["flow"])
    f_tf32 = _epe(R0["flow"], Rtf["flow"])
    floor = max(f_bench, f_cpu)
    e_new = _epe(N["flow"], R0["flow"])
    tag = f"[trained] {ckpt} {cls} {H}x{W}:"
    _note(f"{tag} ref-vs-ref floor  EPE(R0,R0 again)={f_self:.2e}  EPE(R0,cudnn.benchmark)={f_bench:.2e}  "
          f"EPE(R0,ref-on-CPU)={f_cpu:.2e}  EPE(R0,stock TF32)={f_tf32:.2e}  "
          f"max-abs: bench {(R0['flow'] - R1['flow']).abs().max().item():.2e} cpu "
          f"{(R0['flow'] - Rcpu['flow']).abs().max().item():.2e} tf32 {(R0['flow'] - Rtf['flow']).abs().max().item():.2e}")
    _note(f"{tag} ours  EPE(new,R0)={e_new:.2e}  max-abs {(N['flow'] - R0['flow']).abs().max().item():.2e}  "
          f"EPE(new,GT)={_epe(N['flow'], gt):.4f}  EPE(R0,GT)={_epe(R0['flow'], gt):.4f}  "
          f"EPE(new,new again)={_epe(N['flow'], N2['flow']):.2e}  mask agreement(new,R0)="
          f"{_mask_agreement(N['flow'], R0['flow'], cuda):.4f}%  (R0,cpu)={_mask_agreement(R0['flow'], Rcpu['flow'], cuda):.4f}%")
    assert e_new <= K_FLOOR * max(floor, ABS_FLOOR), (e_new, floor)
    # accuracy against the ground truth must be the reference's accuracy
    assert abs(_epe(N["flow"], gt) - _epe(R0["flow"], gt)) <= K_FLOOR * max(floor, ABS_FLOOR)
    if "occ" in R0:
        d_new, a_new = _occ_stats(N["occ"], R0["occ"])
        d_b, a_b = _occ_stats(R1["occ"], R0["occ"])
        d_c, a_c = _occ_stats(Rcpu["occ"], R0["occ"])
        d_t, a_t = _occ_stats(Rtf["occ"], R0["occ"])
        _note(f"{tag} occ   mean|d logit| new {d_new:.2e} (floor: bench {d_b:.2e}, cpu {d_c:.2e}, tf32 {d_t:.2e})  "
              f"round(sigmoid) agreement new {a_new:.4f}% (bench {a_b:.4f}%, cpu {a_c:.4f}%, tf32 {a_t:.4f}%)  "
              f"max-abs new {(N['occ'] - R0['occ']).abs().max().item():.2e} cpu {(Rcpu['occ'] - R0['occ']).abs().max().item():.2e}")
        occ_floor = max(d_b, d_c)
        assert d_new <= K_FLOOR * max(occ_floor, 1e-3), (d_new, occ_floor)
        # percent of pixels whose thresholded occlusion DISAGREES: within K_FLOOR x the reference's own disagreement
        assert 100.0 - a_new <= K_FLOOR * max(100.0 - min(a_b, a_c), 0.05), (a_new, a_b, a_c)


def test_install_runs_unmodified_reference_model_on_our_kernels(cuda, refmodels):
    """irr_b200.install(): the UNMODIFIED models.IRR_PWC, constructed after install(), runs its cost volumes / warps /
    resizes on the irr_b200 kernels (launch counter moves) and reproduces the un-patched reference to the noise floor;
    patch_instances() does the same for a model that already exists; uninstall() restores the reference."""
    import irr_b200
    from irr_b200 import ops
    from irr_b200.synthetic import synthetic_pair
    sd = _state("IRR-PWC_sintel")
    i1, i2, gt = synthetic_pair(1, 218, 512, seed=3, max_flow=10.0)
    pm = sys.modules[refmodels.__name__ + ".pwc_modules"]
    orig_ccv = pm.compute_cost_volume
    with _backend(False, False):
        R0 = _ref_forward(refmodels, "IRR_PWC", sd, i1, i2, cuda)
        with _backend(True, False):
            R1 = _ref_forward(refmodels, "IRR_PWC", sd, i1, i2, cuda)
        with _cuda_shim():
            Rcpu = _ref_forward(refmodels, "IRR_PWC", sd, i1, i2, torch.device("cpu"))
        # what the reference differs from itself by (cudnn.benchmark on/off, GPU vs host): with trained weights the hard
        # warp mask amplifies 1e-7 differences to ~5e-2 px (profiles/r02_parity_trained.txt)
        floor = max(_epe(R0["flow"], R1["flow"]), _epe(R0["flow"], Rcpu["flow"]), ABS_FLOOR)
        existing = refmodels.IRR_PWC(None)
        existing.load_state_dict(sd)
        existing = existing.to(cuda).eval()
        ops.set_grid_mode(ops.GRID_RECIP_MUL)
        try:
            patched = irr_b200.install(refmodels)
            assert any(p.endswith("pwc_modules.compute_cost_volume") for p in patched)
            assert any(p.endswith("IRR_PWC.WarpingLayer") for p in patched)
            assert pm.compute_cost_volume is not orig_ccv
            ops.LAUNCHES = 0
            P = _ref_forward(refmodels, "IRR_PWC", sd, i1, i2, cuda)   # constructed AFTER install()
            n_new = ops.LAUNCHES
            assert n_new >= 10 + 36 + 20   # 10 cost volumes, 36 warps, >= 20 resizes
            # a model that already existed: cost volume / resize are module-level names (patched), its WarpingLayer
            # instance is not — until patch_instances()
            ops.LAUNCHES = 0
            with torch.no_grad():
                existing({"input1": i1.to(cuda), "input2": i2.to(cuda)})
            n_partial = ops.LAUNCHES
            assert 0 < n_partial < n_new
            assert irr_b200.patch_instances(existing) == 1
            ops.LAUNCHES = 0
            with torch.no_grad():
                E = {k: v.cpu() for k, v in existing({"input1": i1.to(cuda), "input2": i2.to(cuda)}).items()}
            assert ops.LAUNCHES == n_new
        finally:
            irr_b200.uninstall()
            ops.set_grid_mode(ops.GRID_TRUE_DIV)
    assert pm.compute_cost_volume is orig_ccv
    ops.LAUNCHES = 0
    with _backend(False, False), torch.no_grad():
        existing({"input1": i1.to(cuda), "input2": i2.to(cuda)})
    assert ops.LAUNCHES == 0   # fully restored
    e_p, e_e = _epe(P["flow"], R0["flow"]), _epe(E["flow"], R0["flow"])
    _note(f"[install] unmodified models.IRR_PWC on irr_b200 kernels 218x512: EPE(patched,R0)={e_p:.2e} "
          f"EPE(patched-instance,R0)={e_e:.2e} floor={floor:.2e} launches/forward={n_new}")
    assert e_p <= K_FLOOR * floor and e_e <= K_FLOOR * floor


def test_install_keeps_autograd(cuda, refmodels):
    """after install() a reference model in train mode must keep back-propagating through the cost volume,
    warp and resize (our correlation backward kernel; the reference's own code for warp / resize)."""
    import irr_b200
    pm = sys.modules[refmodels.__name__ + ".pwc_modules"]
    a = torch.randn(1, 8, 12, 16, device=cuda, requires_grad=True)
    b = torch.randn(1, 8, 12, 16, device=cuda, requires_grad=True)
    fl = (0.05 * torch.randn(1, 2, 12, 16, device=cuda)).requires_grad_(True)
    ref_cv = pm.compute_cost_volume(a, b, {"max_disp": 4})
    g_ref = torch.autograd.grad(ref_cv.square().sum(), [a, b])
    irr_b200.install(refmodels)
    try:
        cv = pm.compute_cost_volume(a, b, {"max_disp": 4})
        assert cv.grad_fn is not None
        g = torch.autograd.grad(cv.square().sum(), [a, b])
        for x, y in zip(g, g_ref):
            assert (x - y).abs().max().item() <= 1e-4 * max(1.0, y.abs().max().item())
        w = pm.WarpingLayer()(a, fl, 48, 64, 0.05)
        assert w.grad_fn is not None
        up = pm.upsample2d_as(a, torch.empty(1, 1, 24, 32, device=cuda))
        assert up.grad_fn is not None and tuple(up.shape) == (1, 8, 24, 32)
        with torch.no_grad():   # inference calls take the kernels
            from irr_b200 import ops
            ops.LAUNCHES = 0
            pm.WarpingLayer()(a, fl, 48, 64, 0.05); pm.upsample2d_as(a, up); pm.compute_cost_volume(a, b, {"max_disp": 4})
            assert ops.LAUNCHES == 3
    finally:
        irr_b200.uninstall()
