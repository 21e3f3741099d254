"""Per-role cycle counters of the 3xF16 conv kernel (CTA 0) for the shapes that dominate cfg 3 — run on the GPU box."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from irr_b200 import ops, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
buf = torch.zeros(32, dtype=torch.int64, device=dev)
lib.irrdbg_conv_counters.argtypes = [ctypes.c_void_p]
lib.irrdbg_conv_counters(buf.data_ptr())
ADD = os.environ.get('IRR_CONV_ADDEND')
SHAPES = [(16, 531, 109, 256, 32, 3, 1), (16, 565, 109, 256, 128, 3, 1), (16, 32, 436, 1024, 32, 3, 1), (16, 128, 109, 256, 128, 3, 2),
          (16, 467, 109, 256, 64, 3, 1)]
for (B, Cin, H, W, Cout, k, d) in SHAPES:
    x = torch.randn(B, Cin, H, W, device=dev); w = torch.randn(Cout, Cin, k, k, device=dev) * 0.02; b = torch.zeros(Cout, device=dev)
    pk = ops.pack_weights(w, ops.MATH_TC_3XF16)
    ops.conv2d(x, pk, b, Cout, k, 1, d, math=ops.MATH_TC_3XF16); torch.cuda.synchronize()
    buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv2d(x, pk, b, Cout, k, 1, d, math=ops.MATH_TC_3XF16); e1.record(); torch.cuda.synchronize()
    c = buf.cpu().tolist()
    nk = max(c[5], 1); nm = max(c[13], 1); nx = max(c[17], 1); ne = max(c[26], 1)
    print(f"{Cin}->{Cout} d{d} {H}x{W}: {e0.elapsed_time(e1):.3f} ms\n"
          f"  producer w0 per own kb ({c[5]} kb): x_wait {c[0]/nk:.0f}  convert+bar {c[1]/nk:.0f}  tap_lds {c[2]/nk:.0f}  a_empty_wait {c[3]/nk:.0f}  st+arrive {c[4]/nk:.0f}\n"
          f"  mma w8 per kb ({c[13]} kb): b_wait {c[9]/nm:.0f}  a_wait {c[10]/nm:.0f}  issue {c[11]/nm:.0f}  acc_empty_wait total {c[8]}\n"
          f"  x loader: x_empty_wait per stage {c[16]/nx:.0f} ({c[17]} stages)   epilogue w12: acc_full wait per item {c[24]/ne:.0f}  other {c[25]/ne:.0f}  tmem_ld {c[27]/ne:.0f}  math+stores {c[28]/ne:.0f} ({c[26]} items)", flush=True)
lib.irrdbg_conv_counters(None)
