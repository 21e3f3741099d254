"""Experiment: two independent batch-8 steps in flight on two streams (the launch-bound coarse levels of one step overlap
the fat levels of the other) vs. the same steps back to back.  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irr_b200
from irr_b200 import ops, pwc_modules
from irr_b200 import synthetic as S

dev = torch.device("cuda:0")
pwc_modules.set_conv_math(ops.MATH_TC_3XF16)
B, H, W = 8, 436, 1024
model = irr_b200.IRR_PWC(None)
irr_b200.load_state_dict_strict(model, S.synthetic_params("IRR_PWC", seed=1234, gain=0.7))
model = model.to(dev).eval()
graphs, outs, inps = [], [], []   # keep the inputs alive: torch.cuda.graph() empties the allocator cache before a capture
for k in range(2):
    i1, i2, _ = S.synthetic_pair(B, H, W, seed=7 + k, max_flow=8.0)
    inp = {"input1": i1.to(dev), "input2": i2.to(dev)}
    with torch.no_grad():
        for _ in range(2):
            model(inp)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            model(inp)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = model(inp)
    graphs.append(g); outs.append(out); inps.append(inp)
torch.cuda.synchronize()
K = 12
def seq():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(K):
        graphs[i & 1].replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
def par():
    s = [torch.cuda.Stream(), torch.cuda.Stream()]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for st in s:
        st.wait_stream(torch.cuda.current_stream())
    for i in range(K):
        with torch.cuda.stream(s[i & 1]):
            graphs[i & 1].replay()
    for st in s:
        torch.cuda.current_stream().wait_stream(st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
for g in graphs:
    g.replay()
torch.cuda.synchronize()
ref = [o.clone() for o in (outs[0]["flow"], outs[1]["flow"])]   # outputs of one replay each (capture does not execute)
for name, fn in (("sequential", seq), ("two in flight", par), ("sequential", seq), ("two in flight", par)):
    fn()
    t = fn()
    same = all(torch.equal(r, o["flow"]) for r, o in zip(ref, outs))
    print(f"{name:14s}: {t:.3f} ms per step = {B / t * 1e3:.1f} pairs/s   outputs identical: {same}", flush=True)
