"""Micro-benchmark of the correlation kernel (run on the GPU box): plain / fused with smooth small flow / fused with
wild flow, at the five cfg-3 level shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irr_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
PEAK = 6581.2
for (B, C, H, W) in [(16, 32, 109, 256), (16, 64, 55, 128), (16, 96, 28, 64), (16, 128, 14, 32), (16, 196, 7, 16), (1, 64, 64, 128)]:
    f = torch.randn(B, C, H, W, device=dev)
    lo = torch.randn(B, 2, 3, 4, device=dev)
    smooth = torch.nn.functional.interpolate(lo, size=[H, W], mode="bicubic", align_corners=True) * 0.05 * 4.0 * torch.tensor([1024 / W, 436 / H], device=dev).view(1, 2, 1, 1)
    wild = torch.randn(B, 2, H, W, device=dev) * 0.05 * 30.0 * torch.tensor([1024 / W, 436 / H], device=dev).view(1, 2, 1, 1)
    out = torch.empty(B, 81, H, W, device=dev)
    byts = B * H * W * (8 * C + 324)
    line = f"{(B, C, H, W)}: "
    for name, fn in [("plain", lambda: ops.correlation(f, f, out=out, shift=B // 2, slope=0.1)),
                     ("fused/smooth", lambda: ops.warp_correlation(f, f, smooth, 436, 1024, 0.05, out=out, shift=B // 2, slope=0.1)),
                     ("fused/wild", lambda: ops.warp_correlation(f, f, wild, 436, 1024, 0.05, out=out, shift=B // 2, slope=0.1))]:
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 5
        line += f"{name} {us:7.1f} us {byts / us / 1e3:6.0f} GB/s ({byts / us / 1e3 / PEAK * 100:4.1f}%) | "
    print(line, flush=True)
