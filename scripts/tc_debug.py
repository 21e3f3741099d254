"""Decode how the tcgen05 conv kernel maps (n, k) — run on the GPU box under `timeout`.  Prints, for 1x1 convs with
w[n][c] = n*32 + c (exact in tf32) and x = indicator of channel c0, the decoded (n', c') at a few (n, m)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from irr_b200 import ops

dev = torch.device("cuda:0")
for mode, math in (("tf32", ops.MATH_TC_TF32), ("3xtf32", ops.MATH_TC_3XTF32)):
    Cin, Cout, H, W = 32, 32, 8, 32
    w = (torch.arange(Cout).view(-1, 1) * 32 + torch.arange(Cin).view(1, -1)).float().view(Cout, Cin, 1, 1)
    packed = ops.pack_weights(w.to(dev), math)
    bias = torch.zeros(Cout, device=dev)
    bad = 0
    for c0 in range(Cin):
        x = torch.zeros(1, Cin, H, W); x[:, c0] = 1.0
        y = ops.conv2d(x.to(dev), packed, bias, Cout, 1, slope=1.0, math=math).cpu()
        exp = (torch.arange(Cout) * 32 + c0).float().view(1, Cout, 1, 1).expand_as(y)
        if not torch.equal(y, exp):
            bad += 1
            if bad <= 6:
                yy = y[0].view(Cout, -1)
                print(f"[{mode}] c0={c0}: n'=", (yy[:8, 0] // 32).int().tolist(), "c'=", (yy[:8, 0] % 32).int().tolist(),
                      " m-variation:", yy[3, [0, 1, 31, 32, 127, 128, 255]].tolist())
    print(f"[{mode}] 1x1 indicator test: {Cin - bad}/{Cin} channels exact")
    # random 3x3 sanity with error stats
    torch.manual_seed(0)
    x = torch.randn(2, 64, 20, 36); w = torch.randn(48, 64, 3, 3) * 0.05; b = torch.randn(48) * 0.1
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1)
    y = ops.conv2d(x.to(dev), ops.pack_weights(w.to(dev), math), b.to(dev), 48, 3, slope=1.0, math=math).cpu().double()
    print(f"[{mode}] 3x3 64->48 max-abs err {float((y - ref).abs().max()):.3e}  (ref max {float(ref.abs().max()):.2f})")
torch.cuda.synchronize()
print("tc_debug done")
