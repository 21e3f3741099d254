"""Per-role cycle counters of the row-rolling conv kernel (CTA 0) — run on the GPU box."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irr_b200 import ops, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
buf = torch.zeros(32, dtype=torch.int64, device=dev)
lib.irrdbg_conv_counters.argtypes = [ctypes.c_void_p]
lib.irrdbg_conv_counters(buf.data_ptr())
for (B, Cin, H, W, Cout) in [(16, 32, 436, 1024, 32), (16, 32, 436, 1024, 1), (16, 16, 218, 512, 16)]:
    x = torch.randn(B, Cin, H, W, device=dev); w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05; b = torch.zeros(Cout, device=dev)
    pk = ops.pack_weights(w, ops.MATH_TC_3XF16)
    ops.conv2d(x, pk, b, Cout, 3, math=ops.MATH_TC_3XF16); torch.cuda.synchronize()
    buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv2d(x, pk, b, Cout, 3, math=ops.MATH_TC_3XF16); e1.record(); torch.cuda.synchronize()
    c = buf.cpu().tolist()
    nr = max(c[5], 1)
    print(f"{Cin}->{Cout} {H}x{W}: {e0.elapsed_time(e1):.3f} ms   (per input row, {c[5]} rows)\n"
          f"  producer w0: x_wait {c[0]/nr:.0f}  convert+bar {c[1]/nr:.0f}  tap_lds {c[2]/nr:.0f}  a_empty_wait {c[3]/nr:.0f}  st+arrive {c[4]/nr:.0f}")
    for ky in range(3):
        o = 8 + 6 * ky; n = max(c[o + 5], 1) / 3
        print(f"  issuer ky={ky}: a_wait {c[o]/n:.0f}  acc_empty_wait {c[o+1]/n:.0f}  issue {c[o+2]/n:.0f}")
    n = max(c[31], 1)
    print(f"  epilogue w8 (per output row, {c[31]}): acc_full_wait {c[26]/n:.0f}  drain+zero {c[27]/n:.0f}  math+stores {c[28]/n:.0f}", flush=True)
lib.irrdbg_conv_counters(None)
