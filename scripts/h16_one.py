"""One conv shape through the 3xF16 kernel (own process: a device trap poisons the context).
usage: python scripts/h16_one.py B Cin H W Cout k s d"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irr_b200 import ops
B, Cin, H, W, Cout, k, s, d = [int(a) for a in sys.argv[1:9]]
dev = torch.device("cuda:0")
torch.manual_seed(1)
x = torch.randn(B, Cin, H, W)
w = torch.randn(Cout, Cin, k, k) * (2.0 / (Cin * k * k)) ** 0.5
b = torch.randn(Cout) * 0.1
ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x.double(), w.double(), b.double(), stride=s,
                                                                padding=((k - 1) * d) // 2, dilation=d), 0.1)
pk = ops.pack_weights(w.to(dev), ops.MATH_TC_3XF16)
y = ops.conv2d(x.to(dev), pk, b.to(dev), Cout, k, s, d, math=ops.MATH_TC_3XF16)
torch.cuda.synchronize()
e = (y.cpu().double() - ref).abs()
print(sys.argv[1:9], "gather" if os.environ.get("IRR_CONV_GATHER") == "1" else "staged?",
      "max-abs %.3e mean-abs %.3e refmax %.2f" % (e.max().item(), e.mean().item(), ref.abs().max().item()), flush=True)
if e.max().item() > 1e-3:
    bad = (e > 1e-3).nonzero()
    print("  bad count", bad.shape[0], "of", e.numel(), "first", bad[:6].tolist(), "last", bad[-3:].tolist(), flush=True)
