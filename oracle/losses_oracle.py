"""CPU restatement (torch) of the reference's EVAL-mode losses — TEST INFRASTRUCTURE ONLY (see oracle/README or
DESIGN.md §2): nothing under irr_b200/ imports this.  Pinned against the unmodified reference ``losses.py`` (live check in
tests/test_losses.py where /root/reference is mounted, and tests/golden/losses.npz generated from it by
oracle/gen_golden.py).
"""
import torch


def elementwise_epe(flow, target):
    """losses.py:8-10."""
    return torch.norm(target - flow, p=2, dim=1, keepdim=True)


def f1_score(y_true, y_pred, eps=1e-8):
    """losses.py:24-37 with beta = 1: per-image sums over H, W; mean over batch (and channel)."""
    y_pred, y_true = y_pred.float(), y_true.float()
    tp = (y_pred * y_true).sum(dim=2).sum(dim=2)
    precision = tp / (y_pred.sum(dim=2).sum(dim=2) + eps)
    recall = tp / (y_true.sum(dim=2).sum(dim=2) + eps)
    return torch.mean(precision * recall / (precision * 1 + recall + eps) * 2)


def eval_pwc_bi_occ_upsample(output_dict, target_dict):
    """MultiScaleEPE_PWC_Bi_Occ_upsample.forward, eval branch (losses.py:634-636)."""
    return {"epe": elementwise_epe(output_dict["flow"], target_dict["target1"]).mean(),
            "F1": f1_score(target_dict["target_occ1"], torch.round(torch.sigmoid(output_dict["occ"])))}


def eval_pwc_bi_occ_upsample_kitti(output_dict, target_dict):
    """MultiScaleEPE_PWC_Bi_Occ_upsample_KITTI.forward, eval branch (losses.py:688-697)."""
    valid = target_dict["input_valid"]
    b = target_dict["target1"].size(0)
    mag = torch.norm(target_dict["target1"], p=2, dim=1, keepdim=True) + 1e-8
    epe = elementwise_epe(output_dict["flow"], target_dict["target1"]) * valid
    epe_img = epe.view(b, -1).sum(1) / valid.view(b, -1).sum(1)
    outl = (epe > 3).float() * ((epe / mag) > 0.05).float() * valid
    return {"epe": epe_img.mean(), "outlier": (outl.view(b, -1).sum(1) / valid.view(b, -1).sum(1)).mean()}


def synthetic_eval_case(seed, B=3, H=37, W=53):
    """Deterministic eval inputs: a flow prediction near its target (with a few gross outliers), occlusion logits and a
    binary occlusion target, a KITTI-style sparse valid mask."""
    g = torch.Generator().manual_seed(seed)
    target = torch.randn(B, 2, H, W, generator=g) * 8.0
    flow = target + torch.randn(B, 2, H, W, generator=g) * 0.7
    big = torch.rand(B, 1, H, W, generator=g) < 0.05
    flow = flow + big.float() * torch.randn(B, 2, H, W, generator=g) * 12.0
    occ = torch.randn(B, 1, H, W, generator=g) * 3.0
    tocc = (torch.rand(B, 1, H, W, generator=g) < 0.3).float()
    valid = (torch.rand(B, 1, H, W, generator=g) < 0.4).float()
    return ({"flow": flow, "occ": occ}, {"target1": target, "target_occ1": tocc, "input_valid": valid})
