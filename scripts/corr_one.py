"""One correlation launch (debug aid): python scripts/corr_one.py B C H W [fused] [bf16]
bf16: f1 / f2 are read from packed bf16 storage (irr_warp_correlation_fwd_dt) and compared with the fp32-storage kernel on the
same rounded values (must be bit-identical)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irr_b200 import ops
B, C, H, W = map(int, sys.argv[1:5]); fused = "fused" in sys.argv[5:]; bf16 = "bf16" in sys.argv[5:]
dev = torch.device("cuda:0"); torch.manual_seed(0)
f1 = torch.randn(B, C, H, W, device=dev); f2 = torch.randn(B, C, H, W, device=dev)
flow = torch.randn(B, 2, H, W, device=dev) * 0.05
if bf16:
    a16, b16 = ops.round_bf16_store(f1), ops.round_bf16_store(f2)   # f1 / f2 are rounded in place as well
    if fused:
        out = ops.warp_correlation(a16, b16, flow, H * 4, W * 4, 0.05, shift=B // 2, slope=0.1)
        ref = ops.warp_correlation(f1, f2, flow, H * 4, W * 4, 0.05, shift=B // 2, slope=0.1)
    else:
        out = ops.correlation(a16, b16, shift=B // 2, slope=0.1)
        ref = ops.correlation(f1, f2, shift=B // 2, slope=0.1)
    torch.cuda.synchronize()
    print(sys.argv[1:], "max |bf16 storage - fp32 storage| =", (out - ref).abs().max().item(), flush=True)
    sys.exit(0)
if fused:
    out = ops.warp_correlation(f1, f2, flow, H * 4, W * 4, 0.05, shift=B // 2, slope=0.1)
    os.environ["IRR_CORR_NO_TMA"] = "1"
    ref = ops.warp_correlation(f1, f2, flow, H * 4, W * 4, 0.05, shift=B // 2, slope=0.1)
else:
    out = ops.correlation(f1, f2, shift=B // 2, slope=0.1)
    os.environ["IRR_CORR_NO_TMA"] = "1"
    ref = ops.correlation(f1, f2, shift=B // 2, slope=0.1)
torch.cuda.synchronize()
print(sys.argv[1:], "max |tma - cp.async| =", (out - ref).abs().max().item(), flush=True)
