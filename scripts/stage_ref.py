#!/usr/bin/env python
"""Stage the UNMODIFIED reference model package and its trained checkpoints under baseline/_ref/ (git-ignored, NOT
gpurun-ignored: it travels to the GPU box with the snapshot, it never enters history).

    python scripts/stage_ref.py [--ref /root/reference]

What is staged:  models/*.py (the PWC / FlowNet python modules — `import models` needs nothing but torch; the
correlation_package C++/CUDA extension is not staged, no PWC model imports it) and
saved_check_point/pwcnet/{IRR-PWC_sintel,IRR-PWC_kitti,PWCNet,PWCNet-irr}/checkpoint_best.ckpt.
Used by: tests/test_trained_gpu.py (trained-checkpoint parity, same-GPU noise floor, install() drop-in test),
bench.py's `torch_gpu_baseline` / `--impl reference` legs (cpu_baseline.kind "reference" when present).
Nothing under irr_b200/ reads baseline/_ref.
"""
import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPTS = ["IRR-PWC_sintel", "IRR-PWC_kitti", "PWCNet", "PWCNet-irr"]


def stage(ref="/root/reference", dst=os.path.join(ROOT, "baseline", "_ref")):
    if not os.path.isdir(os.path.join(ref, "models")):
        return False
    os.makedirs(os.path.join(dst, "models"), exist_ok=True)
    for fn in sorted(os.listdir(os.path.join(ref, "models"))):
        if fn.endswith(".py"):
            shutil.copyfile(os.path.join(ref, "models", fn), os.path.join(dst, "models", fn))
    for name in CKPTS:
        src = os.path.join(ref, "saved_check_point", "pwcnet", name, "checkpoint_best.ckpt")
        if os.path.isfile(src):
            d = os.path.join(dst, "saved_check_point", "pwcnet", name)
            os.makedirs(d, exist_ok=True)
            out = os.path.join(d, "checkpoint_best.ckpt")
            if not (os.path.isfile(out) and os.path.getsize(out) == os.path.getsize(src)):
                shutil.copyfile(src, out)
    with open(os.path.join(dst, "STAGED_FROM"), "w") as f:
        f.write(ref + "\n")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    ok = stage(a.ref)
    print("staged" if ok else "reference not mounted: nothing staged")
    sys.exit(0)
