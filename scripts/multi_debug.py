"""Debug: irr_conv2d_fwd_multi per-segment errors on several shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from irr_b200 import ops
cuda = torch.device("cuda:0")
def rs(seed, shape): return np.random.RandomState(seed).standard_normal(shape).astype("float32")
for shape in [(16, 467, 28, 64, 2), (4, 467, 28, 64, 2), (16, 64, 28, 64, 2), (3, 467, 109, 256, 2), (16, 467, 28, 64, 1)]:
    for variant in ("full", "nopre", "noinplace"):
        B, Cin, H, W, co = shape
        x = torch.from_numpy(rs(81, (B, Cin, H, W)))
        w = torch.from_numpy(rs(82, (96 + co, Cin, 3, 3))) * float(np.sqrt(2.0 / (Cin * 9)))
        b = torch.from_numpy(rs(83, (96 + co,))) * 0.1
        pre = torch.from_numpy(rs(84, (B, 64, H, W)))
        add3 = torch.from_numpy(rs(85, (B, co, H, W)))
        full = torch.nn.functional.conv2d(x.double(), w.double(), b.double(), padding=1)
        use_pre = variant != "nopre"
        ref1 = 0.5 * torch.nn.functional.leaky_relu(full[:, :64] + (pre.double() if use_pre else 0), 0.1) + (0 if use_pre else pre.double())
        ref2 = full[:, 64:96]
        ref3 = full[:, 96:] + add3.double()
        o1 = torch.zeros((B, 70, H, W), device=cuda)
        o2 = torch.empty((B, 32, H, W), device=cuda)
        a3 = add3.to(cuda)
        o3 = a3.clone() if variant != "noinplace" else torch.empty_like(a3)
        ops.conv2d_multi(x.to(cuda), ops.pack_weights(w.to(cuda), ops.MATH_TC_3XF16), b.to(cuda), 96 + co, 3, [
            dict(n_begin=0, out=o1[:, 3:67], slope=0.1, alpha=0.5, addend=pre.to(cuda), pre=use_pre),
            dict(n_begin=64, out=o2, slope=1.0),
            dict(n_begin=96, out=o3, slope=1.0, addend=(o3 if variant != "noinplace" else a3))])
        torch.cuda.synchronize()
        e1 = (o1[:, 3:67].cpu().double() - ref1).abs()
        e2 = (o2.cpu().double() - ref2).abs()
        e3 = (o3.cpu().double() - ref3).abs()
        print(shape, variant, f"seg0 {e1.max():.2e} seg1 {e2.max():.2e} seg2 {e3.max():.2e}", flush=True)
        for nm, e in (("seg0", e1), ("seg1", e2), ("seg2", e3)):
            if e.max() > 1e-4:
                bad = (e > 1e-4).nonzero()
                print("   ", nm, "bad count", len(bad), "first", bad[:3].tolist(), "last", bad[-1].tolist(),
                      "per-batch bad", [(int((e[i] > 1e-4).sum())) for i in range(B)][:16])
