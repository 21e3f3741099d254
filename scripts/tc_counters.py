"""Per-role cycle counters of the tcgen05 conv kernel (CTA 0) for a few shapes — run on the GPU box."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irr_b200 import ops, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
buf = torch.zeros(16, dtype=torch.int64, device=dev)
lib.irr_debug_conv_counters.argtypes = [ctypes.c_void_p]
lib.irr_debug_conv_counters(buf.data_ptr())
for (B, Cin, H, W, Cout, k) in [(16, 563, 109, 256, 2, 3), (16, 565, 109, 256, 128, 3), (16, 32, 436, 1024, 32, 3), (16, 128, 109, 256, 128, 3)]:
    for mode, math in (("3xtf32", ops.MATH_TC_3XTF32), ("tf32", ops.MATH_TC_TF32)):
        x = torch.randn(B, Cin, H, W, device=dev); w = torch.randn(Cout, Cin, k, k, device=dev) * 0.02; b = torch.zeros(Cout, device=dev)
        pk = ops.pack_weights(w, math)
        ops.conv2d(x, pk, b, Cout, k, math=math); torch.cuda.synchronize()
        buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.conv2d(x, pk, b, Cout, k, math=math); e1.record(); torch.cuda.synchronize()
        d = buf.cpu().tolist()
        nk = max(d[3], 1); nm = max(d[8], 1)
        print(f"{Cin}->{Cout} {H}x{W} {mode}: {e0.elapsed_time(e1):.3f} ms | producer(w0) per kb: wait_empty {d[0]/nk:.0f} cvt+st {d[1]/nk:.0f} wait::st+arrive {d[2]/nk:.0f} (kb={d[3]}) | "
              f"mma per kb: wait_b {d[4]/nm:.0f} wait_a {d[5]/nm:.0f} issue {d[6]/nm:.0f} acc_wait_total {d[7]} (kb={d[8]})")
lib.irr_debug_conv_counters(None)
