"""Validation metrics of script/validate_model.py:24-95 (hard Dice / IoU with eps 1e-8) on thresholded predictions.

`calculate_dice_score` / `calculate_iou` keep the reference's signatures (masks in, float out).  `validate` runs the
whole loop on the device: eval forward, threshold and the three per-case counts |P&T|, |P|, |T| in one kernel
(b200_seg_counts, exact int64), one host read per batch instead of two `.item()` per case."""
import torch

from . import ops
from .data import DevicePrefetcher


def _ratio(num, den, eps):
    return (num + eps) / (den + eps)


def dice_from_counts(inter, n_pred, n_target, eps=1e-8):
    return _ratio(2.0 * inter, n_pred + n_target, eps)


def iou_from_counts(inter, n_pred, n_target, eps=1e-8):
    return _ratio(float(inter), n_pred + n_target - inter, eps)


def _counts(pred_mask, target_mask):
    pred_mask, target_mask = torch.as_tensor(pred_mask), torch.as_tensor(target_mask)
    if not pred_mask.is_cuda:  # host masks (the reference passes numpy / CPU tensors) are staged, the counting is CUDA
        pred_mask = pred_mask.cuda()
    p = pred_mask.float().contiguous().reshape(1, -1)
    t = target_mask.to(p.device).float().contiguous().reshape(1, -1)
    if p.shape != t.shape:
        raise ValueError(f"mask shapes differ: {tuple(pred_mask.shape)} vs {tuple(target_mask.shape)}")
    return [int(v) for v in ops.seg_counts(p, t, 0.5)[0].tolist()]


def calculate_dice_score(pred_mask, target_mask, eps=1e-8):
    return dice_from_counts(*_counts(pred_mask, target_mask), eps=eps)


def calculate_iou(pred_mask, target_mask, eps=1e-8):
    return iou_from_counts(*_counts(pred_mask, target_mask), eps=eps)


@torch.no_grad()
def validate(model, loader, device, threshold=0.5):
    """per-case Dice / IoU of model.predict(x) > threshold (ModelValidator.validate, validate_model.py:216-248)"""
    model.eval()
    rows = []
    for batch in DevicePrefetcher(loader, device):
        probs = model.predict(batch["image"])
        counts = ops.seg_counts(probs, batch["label"].float().contiguous(), threshold).tolist()
        for cid, (i, p, t) in zip(batch["case_id"], counts):
            rows.append({"case_id": cid, "dice": dice_from_counts(i, p, t), "iou": iou_from_counts(i, p, t)})
    return rows
