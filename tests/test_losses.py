"""Eval-mode losses + harness + on-disk formats.
CPU: the oracle restatement against the reference-generated golden values and (where /root/reference is mounted) the live
reference; the .flo / KITTI-PNG codecs.  GPU: the metrics kernel and the harness against the oracle."""
import os
import struct
import sys
import zlib

import numpy as np
import pytest
import torch

from oracle import irr_oracle as O
from oracle import losses_oracle as LO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_losses_oracle_vs_golden(golden_dir, seed):
    g = np.load(f"{golden_dir}/losses.npz")
    out, tgt = LO.synthetic_eval_case(seed)
    a = LO.eval_pwc_bi_occ_upsample(out, tgt)
    b = LO.eval_pwc_bi_occ_upsample_kitti(out, tgt)
    assert abs(a["epe"].item() - g[f"sintel_{seed}"][0]) <= 1e-6 and abs(a["F1"].item() - g[f"sintel_{seed}"][1]) <= 1e-6
    assert abs(b["epe"].item() - g[f"kitti_{seed}"][0]) <= 1e-6 and abs(b["outlier"].item() - g[f"kitti_{seed}"][1]) <= 1e-6


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted (GPU box)")
def test_losses_oracle_bit_exact_vs_live_reference():
    sys.path.insert(0, REF)
    try:
        import losses as ref_losses
    finally:
        sys.path.remove(REF)

    class Args:
        batch_size = 2
        model_div_flow = 0.05
    out, tgt = LO.synthetic_eval_case(5, B=2, H=21, W=40)
    with torch.no_grad():
        a = ref_losses.MultiScaleEPE_PWC_Bi_Occ_upsample(Args()).eval()(dict(out), dict(tgt))
        b = ref_losses.MultiScaleEPE_PWC_Bi_Occ_upsample_KITTI(Args()).eval()(dict(out), dict(tgt))
    x, y = LO.eval_pwc_bi_occ_upsample(out, tgt), LO.eval_pwc_bi_occ_upsample_kitti(out, tgt)
    for k in a:
        assert torch.equal(a[k], x[k]), k
    for k in b:
        assert torch.equal(b[k], y[k]), k


def test_loss_and_harness_api_mirrors_reference():
    from irr_b200 import harness, losses
    for name in ("MultiScaleEPE_PWC_Bi_Occ_upsample", "MultiScaleEPE_PWC_Bi_Occ_upsample_Sintel",
                 "MultiScaleEPE_PWC_Bi_Occ_upsample_KITTI", "MultiScaleEPE_PWC"):
        m = getattr(losses, name)(None)
        assert hasattr(m, "_weights") and hasattr(m, "_args")
        with pytest.raises(RuntimeError):      # training branch is out of scope and says so
            m.train()({}, {})
    assert losses.MultiScaleEPE_PWC_Bi_Occ_upsample(None)._weights == [0.32, 0.08, 0.02, 0.01, 0.005, 0.00125, 0.0003125]
    mal = harness.ModelAndLoss(None, torch.nn.Identity(), None, losses.MultiScaleEPE_PWC(None))
    assert mal.model is not None and mal.evaluation_loss is not None and mal.training_loss is None
    assert mal.num_parameters() == 0


# ------------------------------------------------------------------------------------------------ on-disk formats
def test_flo_roundtrip_and_layout(tmp_path):
    from irr_b200 import flow_io as F
    rng = np.random.default_rng(0)
    uv = (rng.normal(size=(5, 7, 2)) * 30).astype(np.float32)
    fn = str(tmp_path / "a.flo")
    F.write_flow(fn, uv)
    raw = open(fn, "rb").read()
    assert raw[:4] == b"PIEH" and struct.unpack("<ii", raw[4:12]) == (7, 5)       # tag, WIDTH, then HEIGHT
    assert np.frombuffer(raw[12:20], np.float32).tolist() == [uv[0, 0, 0], uv[0, 0, 1]]  # (u, v) interleaved per pixel
    assert len(raw) == 12 + 5 * 7 * 2 * 4
    assert np.array_equal(F.read_flo_as_float32(fn), uv)
    F.write_flow(fn, uv[:, :, 0], uv[:, :, 1])                                     # the (u, v) two-argument form
    assert np.array_equal(F.read_flo_as_float32(fn), uv)
    open(fn, "wb").write(b"XXXX" + raw[4:])
    with pytest.raises(ValueError):
        F.read_flo_as_float32(fn)


def test_kitti_png_flow_roundtrip(tmp_path):
    from irr_b200 import flow_io as F
    rng = np.random.default_rng(1)
    uv = rng.normal(size=(6, 9, 2)) * 40
    uv[0, 0] = [600.0, -600.0]                     # beyond the format's +-512 px range: clipped to the uint16 limits
    mask = (rng.random((6, 9)) > 0.3).astype(np.float64)
    mask[0, 0] = 1
    fn = str(tmp_path / "f.png")
    F.write_flow_png(fn, uv, mask=mask)
    img = F.read_png16_rgb(fn)
    assert img.dtype == np.uint16 and img.shape == (6, 9, 3)
    assert img[0, 0, 0] == 65535 and img[0, 0, 1] == 0
    assert img[2, 3, 0] == np.uint16(np.clip(uv[2, 3, 0] * 64 + 2 ** 15, 0, 65535))   # truncation, as the reference
    flow, valid = F.read_png_flow(fn)
    assert flow.dtype == np.float64 and valid.shape == (6, 9, 1)
    assert np.array_equal(valid[:, :, 0], mask.astype(np.int64))
    inside = (np.abs(uv) < 511).all(-1) & (mask > 0)
    assert np.abs(flow[inside] - uv[inside]).max() <= 1.0 / 64.0
    assert np.all(flow[mask == 0] == 0)


@pytest.mark.parametrize("ftype", [0, 1, 2, 3, 4])
def test_png_reader_handles_every_filter_type(tmp_path, ftype):
    """KITTI's own files use adaptive filtering: encode rows with each PNG filter type by hand, decode with ours."""
    from irr_b200 import flow_io as F
    rng = np.random.default_rng(ftype)
    h, w, bpp = 4, 5, 6
    img = rng.integers(0, 65536, size=(h, w, 3), dtype=np.uint16)
    rows = img.astype(">u2").view(np.uint8).reshape(h, w * 6).astype(np.int32)
    raw = bytearray()
    prev = np.zeros(w * 6, np.int32)
    for y in range(h):
        cur = rows[y]
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if ftype == 0:
            enc = cur
        elif ftype == 1:
            enc = cur - a
        elif ftype == 2:
            enc = cur - prev
        elif ftype == 3:
            enc = cur - ((a + prev) >> 1)
        else:
            p = a + prev - c
            pa, pb, pc = np.abs(p - a), np.abs(p - prev), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
            enc = cur - pred
        raw.append(ftype)
        raw.extend((enc & 255).astype(np.uint8).tobytes())
        prev = cur
    ch = lambda t, d: struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    fn = str(tmp_path / "x.png")
    with open(fn, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + ch(b"IHDR", struct.pack(">IIBBBBB", w, h, 16, 2, 0, 0, 0)))
        z = zlib.compress(bytes(raw))
        f.write(ch(b"IDAT", z[:7]) + ch(b"IDAT", z[7:]) + ch(b"IEND", b""))   # split IDAT, as real encoders do
    assert np.array_equal(F.read_png16_rgb(fn), img)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 3])
def test_eval_losses_gpu_vs_oracle(cuda, seed):
    from irr_b200 import losses
    out, tgt = LO.synthetic_eval_case(seed, B=3, H=37, W=53)
    dev = lambda d: {k: v.to(cuda) for k, v in d.items()}
    with torch.no_grad():
        a = losses.MultiScaleEPE_PWC_Bi_Occ_upsample(None).eval()(dev(out), dev(tgt))
        b = losses.MultiScaleEPE_PWC_Bi_Occ_upsample_KITTI(None).eval()(dev(out), dev(tgt))
        c = losses.MultiScaleEPE_PWC(None).eval()(dev(out), dev(tgt))
    x, y = LO.eval_pwc_bi_occ_upsample(out, tgt), LO.eval_pwc_bi_occ_upsample_kitti(out, tgt)
    assert abs(a["epe"].item() - x["epe"].item()) <= 1e-5 * x["epe"].item()
    assert abs(a["F1"].item() - x["F1"].item()) <= 1e-6
    assert abs(b["epe"].item() - y["epe"].item()) <= 1e-5 * y["epe"].item()
    assert abs(b["outlier"].item() - y["outlier"].item()) <= 1e-6
    assert abs(c["epe"].item() - x["epe"].item()) <= 1e-5 * x["epe"].item()
    assert a["epe"].is_cuda and a["epe"].dim() == 0


@pytest.mark.gpu
def test_harness_evaluate_matches_reference_moving_average(cuda):
    """harness.evaluate == EvaluationEpoch's batch-size-weighted averages (runtime.py:417-431), with the oracle's forward
    + oracle losses standing in for the reference on the CPU; ragged last batch; host (un-pinned) example dicts."""
    import irr_b200
    from irr_b200 import harness, losses
    p = O.synthetic_params("IRR_PWC", seed=1234, gain=0.7)
    m = irr_b200.IRR_PWC(None)
    irr_b200.load_state_dict_strict(m, p)
    m = m.to(cuda).eval()
    mal = harness.ModelAndLoss(None, m, None, losses.MultiScaleEPE_PWC_Bi_Occ_upsample(None))
    batches, want, n = [], {"epe": 0.0, "F1": 0.0}, 0
    for bi, B in enumerate((2, 2, 1)):
        i1, i2, gt = O.synthetic_pair(B, 64, 96, seed=20 + bi, max_flow=4.0)
        tocc = (torch.rand(B, 1, 64, 96, generator=torch.Generator().manual_seed(bi)) < 0.3).float()
        ex = {"input1": i1, "input2": i2, "target1": gt, "target_occ1": tocc, "index": bi, "basename": [f"b{bi}"] * B}
        batches.append(ex)
        with torch.no_grad():
            o = O.irr_pwc_forward(p, i1, i2)
            l = LO.eval_pwc_bi_occ_upsample(o, ex)
        for k in want:
            want[k] += l[k].item() * B
        n += B
    got = harness.evaluate(mal, batches)
    assert set(got) == {"epe", "F1"}
    assert abs(got["epe"] - want["epe"] / n) <= 2e-3          # forward parity tolerance (EPE of the flows <= 2e-2 px)
    assert abs(got["F1"] - want["F1"] / n) <= 2e-2
    assert harness.evaluate(mal, []) == {}


@pytest.mark.gpu
def test_pipelined_inference_matches_direct_calls(cuda):
    """harness.PipelinedInference: a stream of DIFFERENT batches through the double-buffered, overlapped loop returns, in
    order, exactly what a direct model call returns for each batch (graph replay == eager launches, same kernels)."""
    import irr_b200
    from irr_b200 import harness
    p = O.synthetic_params("IRR_PWC", seed=1234, gain=0.7)
    m = irr_b200.IRR_PWC(None)
    irr_b200.load_state_dict_strict(m, p)
    m = m.to(cuda).eval()
    H, W, B, n = 64, 96, 2, 5
    runner = harness.PipelinedInference(m, B, H, W, cuda)
    hosts, outs = [], []
    for i in range(n):
        i1, i2, _ = O.synthetic_pair(B, H, W, seed=40 + i, max_flow=4.0)
        hosts.append((i1.pin_memory(), i2.pin_memory()))
        outs.append((torch.empty(B, 2, H, W).pin_memory(), torch.empty(B, 1, H, W).pin_memory()))
    for (h1, h2), (of, oo) in zip(hosts, outs):
        runner.submit(h1, h2, of, oo)
    runner.sync()
    for (h1, h2), (of, oo) in zip(hosts, outs):
        with torch.no_grad():
            ref = m({"input1": h1.to(cuda), "input2": h2.to(cuda)})
        # not torch.equal: the rolling conv kernel's three MMA issuers accumulate into one TMEM accumulator in issue
        # order, which varies run to run (last-ulp differences); a wrong / stale batch would be O(1) off
        assert (of - ref["flow"].cpu()).abs().max().item() <= 1e-4
        assert (oo - ref["occ"].cpu()).abs().max().item() <= 1e-4
    assert (outs[0][0] - outs[1][0]).abs().max().item() > 1e-2   # the batches really differ
