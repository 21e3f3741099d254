"""Run-to-run determinism probe: eager vs eager, graph vs graph, eager vs graph, with / without the side stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, irr_b200
from irr_b200 import synthetic as O   # parameter / input generators
dev = torch.device("cuda:0")
H, W, B = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 96, 2)
m = irr_b200.IRR_PWC(None); irr_b200.load_state_dict_strict(m, O.synthetic_params("IRR_PWC", seed=1234, gain=0.7)); m = m.to(dev).eval()
i1, i2, _ = O.synthetic_pair(B, H, W, seed=40, max_flow=4.0)
inp = {"input1": i1.to(dev), "input2": i2.to(dev)}
mod = sys.modules["irr_b200.IRR_PWC"]
def diff(a, b): return {k: float((a[k] - b[k]).abs().max()) for k in a}
for side in (True, False):
    mod.set_side_stream(side)
    with torch.no_grad():
        outs = [{k: v.clone() for k, v in m(inp).items()} for _ in range(4)]
    torch.cuda.synchronize()
    print("side", side, "eager run-to-run:", [diff(outs[0], o) for o in outs[1:]], flush=True)
    rec = [dict() for _ in range(2)]
    with torch.no_grad():
        m(inp, record=rec[0]); m(inp, record=rec[1])
    torch.cuda.synchronize()
    for l in rec[0]:
        d = {k: float((rec[0][l][k] - rec[1][l][k]).abs().max()) for k in rec[0][l]}
        bad = {k: v for k, v in d.items() if v != 0.0}
        if bad: print("   first differing level", l, bad); break
