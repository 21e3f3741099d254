"""End-to-end parity of every model class against the oracle, next to the reference's own noise floors
(VERDICT r1 'next' #1/#2): for each (class, size) of the GPU test-suite, with the deterministic synthetic weights,
  ours      EPE(irr_b200 3xF16, oracle on the host)                      [and mean |d occ logit|]
  floor_gc  EPE(oracle's torch-op sequence on THIS GPU, oracle on host)   — what tests/test_models_gpu.py gates against
  floor_64  EPE(oracle on host in fp64, oracle on host fp32)
  floor_1t  EPE(oracle on host with 1 thread, oracle on host all threads)
Writes gpurun_out/parity_floor.txt (committed as profiles/r02_parity_floor.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irr_b200
from irr_b200 import ops, pwc_modules
from oracle import irr_oracle as O

cuda = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
pwc_modules.set_conv_math(ops.MATH_TC_3XF16)
CASES = [("IRR_PWC", 128, 192, 7, 6.0), ("IRR_PWC", 94, 156, 7, 6.0), ("PWCNet", 128, 128, 7, 6.0),
         ("PWCNet_irr_occ_bi", 128, 192, 7, 6.0)]
for n in sorted(O.FAMILY):
    CASES += [(n, 64, 128, 11, 5.0), (n, 94, 156, 12, 5.0)]
CASES += [("IRR_PWC", 436, 1024, 3, 20.0), ("PWCNet_irr_occ_bi", 436, 1024, 3, 20.0), ("IRR_PWC", 375, 1242, 5, 10.0)]
out = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_floor.txt"), "w")
def note(s):
    print(s, flush=True); out.write(s + "\n"); out.flush()
note(f"{'class':18s} {'size':>9s} | {'ours':>9s} {'floor_gc':>9s} {'floor_64':>9s} {'floor_1t':>9s} | ours/max(floor) | occ: ours floor_gc")
nt = torch.get_num_threads()
for name, H, W, seed, mf in CASES:
    p = O.synthetic_params(name, seed=1234, gain=0.7)
    i1, i2, gt = O.synthetic_pair(1, H, W, seed=seed, max_flow=mf)
    with torch.no_grad():
        ref = O.FORWARDS[name](p, i1, i2)
        g = O.FORWARDS[name]({k: v.to(cuda) for k, v in p.items()}, i1.to(cuda), i2.to(cuda))
        r64 = O.FORWARDS[name]({k: v.double() for k, v in p.items()}, i1.double(), i2.double())
        torch.set_num_threads(1)
        r1t = O.FORWARDS[name](p, i1, i2) if H * W <= 128 * 192 else None
        torch.set_num_threads(nt)
        m = irr_b200.MODELS[name](None); irr_b200.load_state_dict_strict(m, p); m = m.to(cuda).eval()
        got = m({"input1": i1.to(cuda), "input2": i2.to(cuda)})
    e = O.epe(got["flow"].cpu(), ref["flow"]).item()
    fg = O.epe(g["flow"].cpu(), ref["flow"]).item()
    f64 = O.epe(r64["flow"].float(), ref["flow"]).item()
    f1 = O.epe(r1t["flow"], ref["flow"]).item() if r1t is not None else float("nan")
    mx = max(fg, f64, f1 if f1 == f1 else 0.0, 1e-9)
    occ = ""
    if "occ" in ref:
        occ = f"{(got['occ'].cpu() - ref['occ']).abs().mean().item():.2e} {(g['occ'].cpu() - ref['occ']).abs().mean().item():.2e}"
    note(f"{name:18s} {H:4d}x{W:<4d} | {e:9.2e} {fg:9.2e} {f64:9.2e} {f1:9.2e} | {e / mx:8.2f} (vs gc {e / max(fg, 1e-9):.2f}) | {occ}")
