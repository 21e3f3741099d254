"""Micro-benchmark of the conv kernels on the layer shapes that dominate cfg 3 (run on the GPU box).
usage: python scripts/bench_conv.py [mode ...]   modes: fp32 3xtf32 tf32     env IRR_CONV_ONLY=<idx> to run one shape (ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irr_b200 import ops

SHAPES = [  # (B, Cin, H, W, Cout, k, stride, dil)
    (16, 565, 109, 256, 128, 3, 1, 1), (16, 243, 109, 256, 128, 3, 1, 1), (16, 128, 109, 256, 128, 3, 1, 2),
    (16, 371, 109, 256, 96, 3, 1, 1), (16, 467, 109, 256, 64, 3, 1, 1), (16, 531, 109, 256, 32, 3, 1, 1),
    (16, 563, 109, 256, 2, 3, 1, 1), (16, 32, 436, 1024, 32, 3, 1, 1), (16, 11, 436, 1024, 32, 3, 1, 1),
    (16, 32, 436, 1024, 1, 3, 1, 1), (16, 64, 109, 256, 32, 3, 1, 1), (16, 3, 436, 1024, 16, 3, 2, 1),
    (16, 16, 218, 512, 16, 3, 1, 1), (16, 128, 55, 128, 128, 3, 1, 1), (16, 196, 7, 16, 32, 1, 1, 1),
]
MODES = {"fp32": ops.MATH_FP32_SIMT, "3xtf32": ops.MATH_TC_3XTF32, "tf32": ops.MATH_TC_TF32, "3xf16": ops.MATH_TC_3XF16}
modes = [m for m in sys.argv[1:] if m in MODES] or ["3xtf32"]
only = os.environ.get("IRR_CONV_ONLY")
dev = torch.device("cuda:0")
torch.manual_seed(0)
for i, (B, Cin, H, W, Cout, k, s, d) in enumerate(SHAPES):
    if only is not None and int(only) != i:
        continue
    x = torch.randn(B, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, k, k, device=dev) * (2.0 / (Cin * k * k)) ** 0.5
    b = torch.randn(Cout, device=dev) * 0.1
    Ho, Wo = ops.conv_out_hw(H, W, k, s, d)
    fl = 2.0 * B * Ho * Wo * Cout * Cin * k * k
    byts = 4.0 * B * (Cin * H * W + Cout * Ho * Wo)
    line = f"[{i:2d}] {Cin:4d}->{Cout:3d} k{k} s{s} d{d} {Ho}x{Wo}: "
    ref = None
    for m in modes:
        math = MODES[m]
        if math != 0 and not ops.tc_supported(Cout, Cin, k, s, d, math):
            line += f"{m}: n/a  "
            continue
        pk = ops.pack_weights(w, math)
        add = torch.randn(B, Cout, Ho, Wo, device=dev) if os.environ.get("IRR_CONV_ADDEND") else None
        y = ops.conv2d(x, pk, b, Cout, k, s, d, math=math, addend=add, alpha=0.1 if add is not None else 1.0)
        torch.cuda.synchronize()
        n = 3 if only is None else 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            ops.conv2d(x, pk, b, Cout, k, s, d, math=math, out=y, addend=add, alpha=0.1 if add is not None else 1.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        err = ""
        if ref is None:
            ref = y.clone()
        else:
            err = f" d={float((y - ref).abs().max()):.1e}"
        line += f"{m}: {ms:7.3f} ms {fl / ms / 1e9:6.1f} TF/s {byts / ms / 1e6:6.0f} GB/s{err} | "
    print(line, flush=True)
