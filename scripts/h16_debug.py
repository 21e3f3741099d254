"""First-contact check of the 3xF16 conv kernel on the GPU box: a few shapes, each synchronised, errors printed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irr_b200 import ops

dev = torch.device("cuda:0")
CASES = [(1, 32, 16, 32, 32, 1, 1, 1), (1, 32, 16, 32, 32, 3, 1, 1), (1, 32, 16, 128, 32, 3, 1, 1),
         (2, 115, 28, 64, 128, 3, 1, 1), (1, 128, 28, 64, 96, 3, 1, 8), (2, 16, 47, 78, 32, 3, 2, 1),
         (1, 64, 109, 256, 64, 3, 1, 1), (1, 35, 13, 39, 128, 3, 1, 1), (1, 128, 14, 32, 196, 3, 1, 1)]
for (B, Cin, H, W, Cout, k, s, d) in CASES:
    torch.manual_seed(1)
    x = torch.randn(B, Cin, H, W)
    w = torch.randn(Cout, Cin, k, k) * (2.0 / (Cin * k * k)) ** 0.5
    b = torch.randn(Cout) * 0.1
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=s,
                                                                    padding=((k - 1) * d) // 2, dilation=d), 0.1)
    pk = ops.pack_weights(w.to(dev), ops.MATH_TC_3XF16)
    torch.cuda.synchronize()
    print("packed hdr", pk[:3].tolist(), flush=True)
    y = ops.conv2d(x.to(dev), pk, b.to(dev), Cout, k, s, d, math=ops.MATH_TC_3XF16)
    torch.cuda.synchronize()
    e = (y.cpu().double() - ref).abs()
    print((B, Cin, H, W, Cout, k, s, d), "max-abs %.3e" % e.max().item(), "mean-abs %.3e" % e.mean().item(),
          "refmax %.2f" % ref.abs().max().item(), flush=True)
    if e.max().item() > 1e-3:
        bad = (e > 1e-3).nonzero()
        print("  bad count", bad.shape[0], "first", bad[:5].tolist(), flush=True)
