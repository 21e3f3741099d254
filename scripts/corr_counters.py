"""Role cycle counters of the TMA correlation kernel (CTA 0), IRR_CORR_CTR=1.  Run on the GPU box."""
import ctypes, os, sys
os.environ["IRR_CORR_CTR"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irr_b200 import ops, _lib
lib = _lib.load(); dev = torch.device("cuda:0")
lib.irrdbg_corr_counters.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
for (B, C, H, W) in [(16, 32, 109, 256), (16, 64, 55, 128), (16, 196, 7, 16)]:
    f = torch.randn(B, C, H, W, device=dev); flow = torch.randn(B, 2, H, W, device=dev) * 0.05
    for fused in (True, False):
        fn = (lambda: ops.warp_correlation(f, f, flow, 436, 1024, 0.05, shift=B // 2, slope=0.1)) if fused else (lambda: ops.correlation(f, f, shift=B // 2, slope=0.1))
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 32)(); lib.irrdbg_corr_counters(buf)
        d = list(buf)
        print(f"{(B,C,H,W)} fused={fused}: {e0.elapsed_time(e1)*1e3:.1f} us | compute: wait_full {d[0]} setup {d[1]} epilogue {d[2]} total {d[3]}"
              f" | issuer: wait_tab {d[8]} wait_empty {d[9]} wait_fpempty {d[10]} total {d[11]}"
              f" | sampler: wait_tab {d[16]} wait_fpfull {d[17]} wait_empty {d[18]} total {d[19]}", flush=True)
