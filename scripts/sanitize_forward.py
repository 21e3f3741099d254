"""Small forwards of the three model classes at sizes whose pyramid levels are NOT multiples of 4 (row-pitched path), with
fp32 and bf16 features, for compute-sanitizer memcheck:
   compute-sanitizer --tool memcheck python scripts/sanitize_forward.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import irr_b200
from irr_b200 import synthetic as S

dev = torch.device("cuda:0")
for name, (H, W), feat in (("IRR_PWC", (94, 156), "fp32"), ("IRR_PWC", (125, 414), "bf16"), ("PWCNet", (94, 156), "fp32"),
                           ("PWCNet_irr_occ_bi", (77, 205), "fp32")):
    m = irr_b200.MODELS[name](None)
    irr_b200.load_state_dict_strict(m, S.synthetic_params(name, seed=1234, gain=0.7))
    m = m.to(dev).eval()
    if feat != "fp32":
        m.set_feature_dtype(feat)
    i1, i2, _ = S.synthetic_pair(2, H, W, seed=11, max_flow=6.0)
    out = m({"input1": i1.to(dev), "input2": i2.to(dev)})
    torch.cuda.synchronize()
    print(name, (H, W), feat, {k: (tuple(v.shape), bool(torch.isfinite(v).all())) for k, v in out.items()}, flush=True)
print("forwards ok")
