"""CPU tests: the oracle restatements against the committed golden vectors (generated from the UNMODIFIED reference by
oracle/gen_golden.py), against each other, and — when /root/reference is mounted (build container only) — live against
the reference's own Python."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import irr_oracle as O
from oracle import ops_np as N

REF = "/root/reference"


def rs(seed, shape, kind="normal"):
    r = np.random.RandomState(seed)
    a = r.standard_normal(shape) if kind == "normal" else np.abs(r.standard_normal(shape)) * 0.3
    return a.astype("float32")


CV_SHAPES = [(1, 64, 64, 128), (2, 196, 7, 16), (2, 32, 109, 256), (1, 96, 24, 78), (1, 3, 5, 5), (3, 17, 9, 13)]


@pytest.mark.parametrize("si", [1, 3, 4, 5])
@pytest.mark.parametrize("kind", ["normal", "lrelu"])
def test_cost_volume_oracles_vs_golden(golden_dir, si, kind):
    g = np.load(f"{golden_dir}/cost_volume.npz")
    shape = CV_SHAPES[si]
    key = "x".join(map(str, shape)) + "_" + kind
    seed = int(g[key + "__seed"])
    f1, f2 = rs(seed, shape, kind), rs(seed + 1000, shape, kind)
    out = N.cost_volume_np(f1, f2)
    sub = out if out.size <= 60000 else out[:, :, ::5, ::7]
    assert np.abs(sub - g[key + "__sub"]).max() <= 1e-6
    t = O.cost_volume(torch.from_numpy(f1), torch.from_numpy(f2)).numpy()
    assert np.abs(t - out).max() <= 1e-6
    if np.prod(shape) <= 40000:  # the scalar C restatement of the .cu kernel is slow by design
        assert np.abs(N.corr_ref_c(f1, f2) - out).max() <= 1e-6


def test_c_oracle_generic_parameters():
    """corr_ref.c (restating correlation_cuda_kernel.cu) at FlowNetC-style parameters against a direct numpy sum."""
    f1, f2 = rs(1, (1, 4, 12, 14)), rs(2, (1, 4, 12, 14))
    pad, k, md, s1, s2 = 6, 3, 4, 2, 2
    out = N.corr_ref_c(f1, f2, pad, k, md, s1, s2)
    p1 = np.pad(f1, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    p2 = np.pad(f2, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    d = md // s2
    assert out.shape[1] == (2 * d + 1) ** 2
    kr = 1
    for (by, bx, tj, ti) in [(0, 0, -2, -2), (2, 3, 1, -1), (out.shape[2] - 1, out.shape[3] - 1, 2, 2)]:
        y1, x1 = by * s1 + md, bx * s1 + md
        y2, x2 = y1 + tj * s2, x1 + ti * s2
        a = p1[0, :, y1 - kr:y1 + kr + 1, x1 - kr:x1 + kr + 1]
        if y2 - kr < 0 or x2 - kr < 0 or y2 + kr + 1 > p2.shape[2] or x2 + kr + 1 > p2.shape[3]:
            continue
        b = p2[0, :, y2 - kr:y2 + kr + 1, x2 - kr:x2 + kr + 1]
        ref = (a * b).sum() / (k * k * 4)
        assert out[0, (tj + d) * (2 * d + 1) + (ti + d), by, bx] == pytest.approx(ref, abs=1e-5)


@pytest.mark.parametrize("ci", range(5))
def test_warp_oracle_mask_bitexact_vs_golden(golden_dir, ci):
    g = np.load(f"{golden_dir}/warp.npz")
    seed, B, C, H, W, him, wim = [int(v) for v in g[f"case{ci}__meta"]]
    x = rs(seed, (B, C, H, W))
    out, mask = N.warp_np(x, g[f"case{ci}__flow"], him, wim, 0.05, g[f"case{ci}__lin_x"], g[f"case{ci}__lin_y"])
    assert (mask != g[f"case{ci}__mask"]).sum() == 0
    assert np.abs(out - g[f"case{ci}__out"]).max() <= 1e-6
    # torch restatement, with the stored linspace vectors (host-independent)
    t = O.warp(torch.from_numpy(x), torch.from_numpy(g[f"case{ci}__flow"]), him, wim, 0.05,
               torch.from_numpy(g[f"case{ci}__lin_x"]), torch.from_numpy(g[f"case{ci}__lin_y"])).numpy()
    assert np.abs(t - g[f"case{ci}__out"]).max() <= 1e-6


def test_resize_oracle_vs_golden(golden_dir):
    g = np.load(f"{golden_dir}/modules.npz")
    t = rs(310, (2, 2, 7, 16))
    assert np.abs(N.resize_ac_np(t, 14, 32) - g["resize_ac__out"]).max() <= 1e-6
    assert np.abs(N.resize_ac_np(t, 13, 39) - g["resize_ac__odd"]).max() <= 1e-6


def test_module_oracles_vs_golden(golden_dir):
    g = np.load(f"{golden_dir}/modules.npz")
    p = O.synthetic_params("IRR_PWC", seed=1234)
    T = lambda s, shape: torch.from_numpy(rs(s, shape))
    with torch.no_grad():
        pyr = O.feature_pyramid(p, T(300, (1, 3, 64, 96)))
        for i, t in enumerate(pyr):
            assert (t - torch.from_numpy(g[f"fpe__{i}"])).abs().max() <= 1e-5
        x5, out = O.dense_estimator(p, "flow_estimators", T(301, (1, 115, 12, 20)))
        assert (out - torch.from_numpy(g["dense__out"])).abs().max() <= 1e-5
        assert (O.context_net(p, "context_networks", T(302, (1, 565, 20, 36))) - torch.from_numpy(g["ctx__out"])).abs().max() <= 1e-5
        fl, d, f = T(303, (2, 2, 14, 22)), T(304, (2, 3, 14, 22)), T(305, (2, 32, 14, 22))
        assert (O.refine_flow(p, fl, d, f) - torch.from_numpy(g["refine_flow__out"])).abs().max() <= 1e-5
        oc, f2 = T(306, (2, 1, 14, 22)), T(307, (2, 32, 14, 22))
        assert (O.refine_occ(p, oc, f, f2) - torch.from_numpy(g["refine_occ__out"])).abs().max() <= 1e-5
        assert (O.occ_upsample(p, oc, T(308, (2, 10, 28, 44))) - torch.from_numpy(g["occ_up__even"])).abs().max() <= 1e-5
        assert (O.occ_upsample(p, oc, T(309, (2, 10, 27, 43))) - torch.from_numpy(g["occ_up__odd"])).abs().max() <= 1e-5


@pytest.mark.parametrize("name,hw", [("IRR_PWC", (128, 192)), ("PWCNet", (128, 128)), ("PWCNet_irr_occ_bi", (128, 192)),
                                     ("IRR_PWC", (94, 156))])
def test_model_oracles_vs_golden(golden_dir, name, hw):
    """Full forward of the oracle vs the reference's golden output.  Tolerance covers host-to-host differences in
    torch's CPU kernels (vector width / thread count) amplified by the hard mask; in the build container it is 0.0."""
    g = np.load(f"{golden_dir}/models.npz")
    H, W = hw
    p = O.synthetic_params(name, seed=1234, gain=0.7)
    i1, i2, _ = O.synthetic_pair(1, H, W, seed=7, max_flow=6.0)
    with torch.no_grad():
        out = O.FORWARDS[name](p, i1, i2)
    for k, v in out.items():
        ref = torch.from_numpy(g[f"{name}_{H}x{W}__{k}"])
        assert O.epe(v, ref).item() <= 1e-3 if k == "flow" else (v - ref).abs().mean().item() <= 1e-3


def test_param_shapes_match_reference_counts():
    n = lambda m: sum(int(np.prod(s)) for s in O.param_shapes(m).values())
    assert n("IRR_PWC") == 6362092
    assert n("PWCNet") == 8639230
    assert len(O.param_shapes("IRR_PWC")) == 124


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted (GPU box)")
def test_oracle_bit_exact_vs_live_reference():
    """Pin: the restatement equals the UNMODIFIED reference Python bit for bit on this host, for all three model
    classes and with a trained checkpoint."""
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    saved = getattr(torch.Tensor, "cuda")
    torch.Tensor.cuda = lambda self, *a, **k: self  # the reference hard-codes .cuda()
    try:
        import models
        for name, cls in [("IRR_PWC", models.IRR_PWC), ("PWCNet", models.PWCNet),
                          ("PWCNet_irr_occ_bi", models.PWCNet_irr_occ_bi)]:
            m = cls(None).eval()
            p = O.synthetic_params(name, seed=99)
            m.load_state_dict(p)
            i1, i2, _ = O.synthetic_pair(1, 64, 96, seed=5, max_flow=3.0)
            with torch.no_grad():
                ref = m({"input1": i1, "input2": i2})
                out = O.FORWARDS[name](p, i1, i2)
            for k in ref:
                assert torch.equal(ref[k], out[k]), (name, k)
        ck = os.path.join(REF, "saved_check_point/pwcnet/IRR-PWC_sintel/checkpoint_best.ckpt")
        sd = torch.load(ck, map_location="cpu", weights_only=False)["state_dict"]
        p = {k[len("_model."):]: v for k, v in sd.items()}
        m = models.IRR_PWC(None).eval()
        m.load_state_dict(p)
        i1, i2, gt = O.synthetic_pair(1, 128, 192, seed=3, max_flow=8.0)
        with torch.no_grad():
            ref = m({"input1": i1, "input2": i2})
            out = O.irr_pwc_forward(p, i1, i2)
        assert torch.equal(ref["flow"], out["flow"]) and torch.equal(ref["occ"], out["occ"])
        assert O.epe(out["flow"], gt).item() < 2.0  # the trained net actually tracks the synthetic flow
    finally:
        torch.Tensor.cuda = saved
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]


def test_oracle_bf16_feature_switch():
    """BASELINE config 5: feature_bf16 rounds the pyramid features (not the images) to bf16 values; default off = the
    reference's fp32 path, bit for bit unchanged."""
    p = O.synthetic_params("IRR_PWC", seed=1234, gain=0.7)
    i1, i2, _ = O.synthetic_pair(1, 64, 96, seed=9, max_flow=4.0)
    with torch.no_grad():
        a = O.irr_pwc_forward(p, i1, i2)
        b = O.irr_pwc_forward(p, i1, i2, feature_bf16=False)
        rec = {}
        c = O.irr_pwc_forward(p, i1, i2, feature_bf16=True, record=rec)
    assert torch.equal(a["flow"], b["flow"]) and torch.equal(a["occ"], b["occ"])
    assert not torch.equal(a["flow"], c["flow"])
    assert O.epe(a["flow"], c["flow"]).item() < 0.5          # a perturbation, not a different answer
    x = rec["l2.x1"]
    assert torch.equal(x, x.bfloat16().float())               # features carry bf16 values
    assert torch.equal(rec["l6.x1"], i1)                      # the image level is not cast


@pytest.mark.parametrize("name", sorted(O.FAMILY))
def test_family_oracle_vs_golden(golden_dir, name):
    """The six remaining PWC-family forwards: oracle restatement vs outputs generated from the
    reference's own classes (oracle/gen_golden.py family).  Host-to-host noise bound as for the other models (F5)."""
    g = np.load(f"{golden_dir}/family.npz")
    seed_p, seed_i, B, H, W = [int(v) for v in g["meta"]]
    p = O.synthetic_params(name, seed=seed_p, gain=0.7)
    i1, i2, _ = O.synthetic_pair(B, H, W, seed=seed_i, max_flow=5.0)
    with torch.no_grad():
        out = O.FORWARDS[name](p, i1, i2)
    assert set(out) == {k.split("__")[1] for k in g.files if k.startswith(name + "__")}
    for k, v in out.items():
        ref = torch.from_numpy(g[f"{name}__{k}"])
        assert v.shape == ref.shape
        if k == "flow":
            assert O.epe(v, ref).item() <= 2e-2
        else:
            assert (v - ref).abs().mean().item() <= 2e-2


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted (GPU box)")
@pytest.mark.parametrize("name", sorted(O.FAMILY))
def test_family_oracle_bit_exact_vs_live_reference(name):
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    saved = getattr(torch.Tensor, "cuda")
    torch.Tensor.cuda = lambda self, *a, **k: self  # the reference hard-codes .cuda()
    try:
        import models
        m = getattr(models, name)(None).eval()
        p = O.synthetic_params(name, seed=4321, gain=0.7)
        assert set(m.state_dict()) == set(p)
        m.load_state_dict(p)
        i1, i2, _ = O.synthetic_pair(1, 64, 64, seed=13, max_flow=4.0)
        with torch.no_grad():
            ref = m({"input1": i1, "input2": i2})
            got = O.FORWARDS[name](p, i1, i2)
        assert set(ref) == set(got)
        for k in ref:
            assert torch.equal(ref[k], got[k]), (name, k)
    finally:
        torch.Tensor.cuda = saved


@pytest.mark.parametrize("si", [0, 1, 2])
def test_cost_volume_backward_oracle_vs_reference_autograd(golden_dir, si):
    """§8(f).4: the numpy backward vs gradients of the reference's own differentiable compute_cost_volume
    (tests/golden/corr_grad.npz, oracle/gen_golden.py corr_grad)."""
    g = np.load(f"{golden_dir}/corr_grad.npz")
    shape = tuple(int(v) for v in g[f"shape__{si}"])
    rs = lambda seed, shp: np.random.RandomState(seed).standard_normal(shp).astype("float32")
    f1, f2, go = rs(500 + si, shape), rs(520 + si, shape), rs(540 + si, (shape[0], 81, shape[2], shape[3]))
    g1, g2 = N.cost_volume_backward_np(f1, f2, go)
    assert np.abs(g1 - g[f"g1__{si}"]).max() <= 2e-6 and np.abs(g2 - g[f"g2__{si}"]).max() <= 2e-6


def test_constant_division_recipe():
    """csrc/common.cuh div_const_rn: q = RN(a*rb); twice q += RN(a - q*b)*rb with rb = RN(1/b) equals IEEE a/b bit for bit.
    fp32 FMA emulated in float64 (the product of two fp32 is exact there and the residual cancels to few bits)."""
    rng = np.random.default_rng(0)

    def fma32(x, y, z):
        return (x.astype(np.float64) * y.astype(np.float64) + z.astype(np.float64)).astype(np.float32)

    def recipe(a, b):
        rb = (np.float32(1.0) / b).astype(np.float32)
        q = (a * rb).astype(np.float32)
        for _ in range(2):
            q = fma32(fma32(-q, b, a), rb, q)
        return q
    for b in (1023.0, 435.0, 1241.0, 374.0, 0.05, 3.0, 255.0, 7.0):
        a = (rng.standard_normal(500_000) * rng.choice([1e-3, 0.1, 1, 10, 100, 1e4], 500_000)).astype(np.float32)
        bb = np.full_like(a, np.float32(b))
        assert np.array_equal((a / bb).astype(np.float32), recipe(a, bb)), b
    a = (rng.standard_normal(1_000_000) * 10).astype(np.float32)
    b = rng.uniform(0.01, 4096, 1_000_000).astype(np.float32)
    assert np.array_equal((a / b).astype(np.float32), recipe(a, b))
