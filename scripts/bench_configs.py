"""Throughput of the OTHER BASELINE.json configs on one B200 (the bench line is config 3; these are parity-test cases
whose speed is recorded for completeness): 1 correlation op alone, 2 PWCNet 256x256 b1, 4 PWCNet_irr_occ_bi 436x1024
(4 pairs per GPU = batch 32 over 8 GPUs), 5 IRR_PWC KITTI shape 375x1242 (4 pairs per GPU) with fp32 and bf16 features.
CUDA-graph replay, inputs resident, CUDA events; one JSON line per config."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irr_b200
from irr_b200 import ops, synthetic as S

dev = torch.device("cuda:0")


def timed(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def model_case(cfg, name, B, H, W, feat="fp32"):
    m = irr_b200.MODELS[name](None)
    irr_b200.load_state_dict_strict(m, S.synthetic_params(name, seed=1234, gain=0.7))
    m = m.to(dev).eval()
    if feat != "fp32":
        m.set_feature_dtype(feat)
    i1, i2, _ = S.synthetic_pair(B, H, W, seed=cfg, max_flow=12.0)
    inp = {"input1": i1.to(dev), "input2": i2.to(dev)}
    with torch.no_grad():
        for _ in range(2):
            m(inp)
        torch.cuda.synchronize()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            m(inp)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            m(inp)
    ms = timed(g.replay)
    print(json.dumps({"config": cfg, "model": name, "B_H_W": [B, H, W], "features": feat, "ms_per_step": round(ms, 3),
                      "pairs_per_s": round(B / (ms * 1e-3), 1)}), flush=True)


# config 1: correlation op alone, 2 x (1, 64, 64, 128)
f1 = torch.randn(1, 64, 64, 128, device=dev); f2 = torch.randn(1, 64, 64, 128, device=dev); out = torch.empty(1, 81, 64, 128, device=dev)
ms = timed(lambda: ops.correlation(f1, f2, out=out), steps=50, warm=5)
by = 64 * 128 * (8 * 64 + 324)
print(json.dumps({"config": 1, "op": "correlation 2x(1,64,64,128)", "us": round(ms * 1e3, 2), "algorithmic_GBps": round(by / (ms * 1e-3) / 1e9, 1),
                  "note": "one 8-tile-row x 4-tile-column launch: 32 CTAs, latency-bound"}), flush=True)
model_case(2, "PWCNet", 1, 256, 256)
model_case(4, "PWCNet_irr_occ_bi", 4, 436, 1024)
model_case(5, "IRR_PWC", 4, 375, 1242, "fp32")
model_case(5, "IRR_PWC", 4, 375, 1242, "bf16")
