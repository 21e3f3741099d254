"""One eager IRR_PWC forward (cfg 3: B=8, 1024x436) bracketed by cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python scripts/ncu_forward.py
(a number taken under ncu is never a bench value: the launch list gives each kernel's SHARE of the step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irr_b200
from irr_b200 import ops, pwc_modules
from irr_b200 import synthetic as O   # parameter / input generators

dev = torch.device("cuda:0")
pwc_modules.set_conv_math({"fp32": 0, "3xtf32": 1, "tf32": 2, "3xf16": 3}[os.environ.get("IRR_MATH", "3xf16")])
m = irr_b200.IRR_PWC(None)
irr_b200.load_state_dict_strict(m, O.synthetic_params("IRR_PWC", seed=1234, gain=0.7))
m = m.to(dev).eval()
B = int(os.environ.get("IRR_BATCH", "8"))
i1, i2, _ = O.synthetic_pair(B, 436, 1024, seed=3, max_flow=20.0)
inp = {"input1": i1.to(dev), "input2": i2.to(dev)}
sys.modules["irr_b200.IRR_PWC"].set_side_stream(False)
for _ in range(2):
    m(inp)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m(inp)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
