"""Generate tests/golden/*.npz by running the UNMODIFIED reference Python (imported read-only from
/root/reference) on seeded inputs.  Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

Inputs are reproducible from numpy's frozen RandomState stream (seed recorded in each file), so fixtures hold
outputs (+ the tiny host linspace vectors the reference used), not inputs.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = os.environ.get("IRR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the reference hard-codes .cuda() (models/pwc_modules.py:111,129; models/IRR_PWC.py:68-71): CPU shim
torch.Tensor.cuda = lambda self, *a, **k: self

import models  # noqa: E402  (the reference package)
from oracle import irr_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
pwc = sys.modules["models.pwc_modules"]


def rs_tensor(seed, shape, kind="normal"):
    rs = np.random.RandomState(seed)
    a = rs.standard_normal(shape) if kind == "normal" else np.abs(rs.standard_normal(shape)) * 0.3
    return torch.from_numpy(a.astype("float32"))


def gen_cost_volume():
    cases = {}
    shapes = [(1, 64, 64, 128), (2, 196, 7, 16), (2, 32, 109, 256), (1, 96, 24, 78), (1, 3, 5, 5), (3, 17, 9, 13)]
    for si, shape in enumerate(shapes):
        for kind in ("normal", "lrelu"):
            seed = 100 + si
            f1 = rs_tensor(seed, shape, kind)
            f2 = rs_tensor(seed + 1000, shape, kind)
            out = pwc.compute_cost_volume(f1, f2, {"max_disp": 4}).numpy()
            key = "x".join(map(str, shape)) + "_" + kind
            sub = out if out.size <= 60000 else out[:, :, ::5, ::7]
            cases[key + "__sub"] = sub
            cases[key + "__sum"] = out.sum(axis=(2, 3), dtype=np.float64)
            cases[key + "__seed"] = np.array(seed)
    np.savez_compressed(os.path.join(OUT, "cost_volume.npz"), **cases)
    print("cost_volume.npz", len(cases))


def gen_warp():
    cases = {}
    wl = pwc.WarpingLayer()
    cfgs = [((2, 32, 28, 64), 436, 1024, 10.0), ((1, 3, 55, 128), 436, 1024, 20.0), ((2, 16, 24, 39), 375, 1242, 6.0),
            ((1, 2, 7, 16), 436, 1024, 3.0), ((1, 5, 1, 9), 64, 64, 1.0)]
    for ci, (shape, him, wim, mag) in enumerate(cfgs):
        seed = 200 + ci
        B, C, H, W = shape
        x = rs_tensor(seed, shape)
        # flow in the reference's "global" units: pixels-at-full-res * div_flow
        flow_px = rs_tensor(seed + 1000, (B, 2, H, W)) * mag
        flow = flow_px * 0.05 * torch.tensor([wim / W, him / H]).view(1, 2, 1, 1)
        out = wl(x, flow, him, wim, 0.05).numpy()
        # the mask the reference computed (recomputed here through the same torch ops)
        grid = torch.add(pwc.get_grid(x), torch.stack([flow[:, 0] * 2 / max(wim - 1, 1) / 0.05,
                                                      flow[:, 1] * 2 / max(him - 1, 1) / 0.05]).transpose(0, 1))
        grid = grid.transpose(1, 2).transpose(2, 3)
        mask = (torch.nn.functional.grid_sample(torch.ones_like(x), grid, align_corners=True) >= 1.0)[:, 0].numpy()
        key = f"case{ci}"
        cases[key + "__out"] = out
        cases[key + "__mask"] = mask
        cases[key + "__flow"] = flow.numpy()  # stored: built with torch ops above
        cases[key + "__lin_x"] = torch.linspace(-1.0, 1.0, W).numpy()
        cases[key + "__lin_y"] = torch.linspace(-1.0, 1.0, H).numpy()
        cases[key + "__meta"] = np.array([seed, B, C, H, W, him, wim])
    np.savez_compressed(os.path.join(OUT, "warp.npz"), **cases)
    print("warp.npz", len(cases))


def gen_models():
    cases = {}
    for name, cls, (B, H, W) in [("IRR_PWC", models.IRR_PWC, (1, 128, 192)), ("PWCNet", models.PWCNet, (1, 128, 128)),
                                 ("PWCNet_irr_occ_bi", models.PWCNet_irr_occ_bi, (1, 128, 192)),
                                 ("IRR_PWC", models.IRR_PWC, (1, 94, 156))]:
        m = cls(None).eval()
        p = O.synthetic_params(name, seed=1234, gain=0.7)
        m.load_state_dict(p)
        i1, i2, gt = O.synthetic_pair(B, H, W, seed=7, max_flow=6.0)
        with torch.no_grad():
            out = m({"input1": i1, "input2": i2})
        for k, v in out.items():
            cases[f"{name}_{H}x{W}__{k}"] = v.numpy()
        cases[f"{name}_{H}x{W}__meta"] = np.array([1234, 7, B, H, W])
    np.savez_compressed(os.path.join(OUT, "models.npz"), **cases)
    print("models.npz", len(cases))


def gen_modules():
    """Module-level KATs through the reference nn.Modules (synthetic weights)."""
    cases = {}
    m = models.IRR_PWC(None).eval()
    p = O.synthetic_params("IRR_PWC", seed=1234)
    m.load_state_dict(p)
    with torch.no_grad():
        x = rs_tensor(300, (1, 3, 64, 96))
        pyr = m.feature_pyramid_extractor(x)
        for i, t in enumerate(pyr):
            cases[f"fpe__{i}"] = t.numpy()
        x = rs_tensor(301, (1, 115, 12, 20))
        a, b = m.flow_estimators(x)
        cases["dense__x5"] = a.numpy()[:, :64]
        cases["dense__out"] = b.numpy()
        x = rs_tensor(302, (1, 565, 20, 36))
        cases["ctx__out"] = m.context_networks(x).numpy()
        fl, d, f = rs_tensor(303, (2, 2, 14, 22)), rs_tensor(304, (2, 3, 14, 22)), rs_tensor(305, (2, 32, 14, 22))
        cases["refine_flow__out"] = m.refine_flow(fl, d, f).numpy()
        oc, f2 = rs_tensor(306, (2, 1, 14, 22)), rs_tensor(307, (2, 32, 14, 22))
        cases["refine_occ__out"] = m.refine_occ(oc, f, f2).numpy()
        xx = rs_tensor(308, (2, 10, 28, 44))
        cases["occ_up__even"] = m.occ_shuffle_upsample(oc, xx).numpy()
        xx = rs_tensor(309, (2, 10, 27, 43))
        cases["occ_up__odd"] = m.occ_shuffle_upsample(oc, xx).numpy()
        t = rs_tensor(310, (2, 2, 7, 16))
        cases["resize_ac__out"] = pwc.upsample2d_as(t, torch.zeros(1, 1, 14, 32)).numpy()
        cases["resize_ac__odd"] = pwc.upsample2d_as(t, torch.zeros(1, 1, 13, 39)).numpy()
    np.savez_compressed(os.path.join(OUT, "modules.npz"), **cases)
    print("modules.npz", len(cases))


def gen_family():
    """Eval outputs of the six remaining PWC-family classes through the reference's own modules."""
    cases = {}
    H, W = 64, 128
    for name in O.FAMILY:
        m = getattr(models, name)(None).eval()
        p = O.synthetic_params(name, seed=1234, gain=0.7)
        m.load_state_dict(p)
        i1, i2, gt = O.synthetic_pair(1, H, W, seed=11, max_flow=5.0)
        with torch.no_grad():
            out = m({"input1": i1, "input2": i2})
        for k, v in out.items():
            cases[f"{name}__{k}"] = v.numpy()
    cases["meta"] = np.array([1234, 11, 1, H, W])
    np.savez_compressed(os.path.join(OUT, "family.npz"), **cases)
    print("family.npz", len(cases))


def gen_corr_grad():
    """Gradients of the reference's own (differentiable) compute_cost_volume, pwc_modules.py:42-62, for a seeded
    upstream gradient — the values correlation_cuda.backward (correlation_cuda_kernel.cu:116-300) is defined to return."""
    cases = {}
    par = {"pad_size": 4, "kernel_size": 1, "max_disp": 4, "stride1": 1, "stride2": 1, "corr_multiply": 1}
    for si, shape in enumerate([(1, 5, 9, 13), (2, 16, 17, 40), (1, 32, 24, 64)]):
        f1 = rs_tensor(500 + si, shape).requires_grad_(True)
        f2 = rs_tensor(520 + si, shape).requires_grad_(True)
        go = rs_tensor(540 + si, (shape[0], 81, shape[2], shape[3]))
        out = pwc.compute_cost_volume(f1, f2, par)
        out.backward(go)
        cases[f"g1__{si}"] = f1.grad.numpy()
        cases[f"g2__{si}"] = f2.grad.numpy()
        cases[f"shape__{si}"] = np.array(shape)
    np.savez_compressed(os.path.join(OUT, "corr_grad.npz"), **cases)
    print("corr_grad.npz", len(cases))


def gen_warp_grad():
    """Gradients of the reference's own WarpingLayer (pwc_modules.py:115-133) under autograd for a seeded upstream
    gradient: d/dx and d/dflow (the hard mask is a constant for autograd).  Inputs are reproducible from the seeds; the
    flow (built with torch ops) and the host linspace vectors are stored."""
    cases = {}
    wl = pwc.WarpingLayer()
    cfgs = [((2, 5, 12, 20), 96, 160, 4.0), ((1, 16, 24, 39), 375, 1242, 6.0), ((2, 3, 28, 64), 436, 1024, 10.0)]
    for ci, (shape, him, wim, mag) in enumerate(cfgs):
        seed = 600 + ci
        B, C, H, W = shape
        x = rs_tensor(seed, shape).requires_grad_(True)
        flow_px = rs_tensor(seed + 1000, (B, 2, H, W)) * mag
        flow = (flow_px * 0.05 * torch.tensor([wim / W, him / H]).view(1, 2, 1, 1)).detach().requires_grad_(True)
        go = rs_tensor(seed + 2000, shape)
        out = wl(x, flow, him, wim, 0.05)
        out.backward(go)
        key = f"case{ci}"
        cases[key + "__gx"] = x.grad.numpy()
        cases[key + "__gflow"] = flow.grad.numpy()
        cases[key + "__flow"] = flow.detach().numpy()
        cases[key + "__lin_x"] = torch.linspace(-1.0, 1.0, W).numpy()
        cases[key + "__lin_y"] = torch.linspace(-1.0, 1.0, H).numpy()
        cases[key + "__meta"] = np.array([seed, B, C, H, W, him, wim])
    np.savez_compressed(os.path.join(OUT, "warp_grad.npz"), **cases)
    print("warp_grad.npz", len(cases))


def gen_losses():
    """Eval-branch outputs of the reference's losses.py on oracle/losses_oracle.synthetic_eval_case inputs."""
    import losses as ref_losses
    from oracle import losses_oracle as LO

    class Args:
        batch_size = 3
        model_div_flow = 0.05
    cases = {}
    for seed in (0, 1, 2):
        out, tgt = LO.synthetic_eval_case(seed)
        with torch.no_grad():
            a = ref_losses.MultiScaleEPE_PWC_Bi_Occ_upsample(Args()).eval()(dict(out), dict(tgt))
            b = ref_losses.MultiScaleEPE_PWC_Bi_Occ_upsample_KITTI(Args()).eval()(dict(out), dict(tgt))
        cases[f"sintel_{seed}"] = np.array([a["epe"].item(), a["F1"].item()], np.float64)
        cases[f"kitti_{seed}"] = np.array([b["epe"].item(), b["outlier"].item()], np.float64)
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **cases)
    print("losses.npz", {k: v.tolist() for k, v in cases.items()})


if __name__ == "__main__":
    torch.manual_seed(0)
    if len(sys.argv) > 1 and sys.argv[1] in ("losses", "family", "corr_grad", "warp_grad"):
        {"losses": gen_losses, "family": gen_family, "corr_grad": gen_corr_grad, "warp_grad": gen_warp_grad}[sys.argv[1]]()
        sys.exit(0)
    gen_cost_volume()
    gen_warp()
    gen_modules()
    gen_models()
    gen_family()
    gen_corr_grad()
    gen_warp_grad()
    gen_losses()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
