import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irr_b200 import ops, _lib
lib = _lib.load(); dev = torch.device("cuda:0")
lib.irr_debug_corr_counters.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
for (B, C, H, W) in [(16, 32, 109, 256), (16, 64, 55, 128), (16, 196, 7, 16)]:
    f = torch.randn(B, C, H, W, device=dev); flow = torch.randn(B, 2, H, W, device=dev) * 0.05
    for fused in (True, False):
        fn = (lambda: ops.warp_correlation(f, f, flow, 436, 1024, 0.05, shift=B // 2, slope=0.1)) if fused else (lambda: ops.correlation(f, f, shift=B // 2, slope=0.1))
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 16)(); lib.irr_debug_corr_counters(buf)
        d = list(buf)
        print(f"{(B,C,H,W)} fused={fused}: {e0.elapsed_time(e1)*1e3:.1f} us | producer t0: wait_empty {d[0]} total {d[1]} chunks {d[2]} | compute t0: wait_full {d[4]} epilogue {d[5]} total {d[6]} chunks {d[7]}")
