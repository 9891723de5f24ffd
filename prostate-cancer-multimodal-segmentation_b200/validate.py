"""Validation metrics of script/validate_model.py:24-95 (hard Dice / IoU with eps 1e-8) on thresholded predictions."""
import torch


def calculate_dice_score(pred_mask, target_mask, eps=1e-8):
    p, t = pred_mask.float().reshape(-1), target_mask.float().reshape(-1)
    inter = (p * t).sum()
    return ((2 * inter + eps) / (p.sum() + t.sum() + eps)).item()


def calculate_iou(pred_mask, target_mask, eps=1e-8):
    p, t = pred_mask.float().reshape(-1), target_mask.float().reshape(-1)
    inter = (p * t).sum()
    union = p.sum() + t.sum() - inter
    return ((inter + eps) / (union + eps)).item()


@torch.no_grad()
def validate(model, loader, device, threshold=0.5):
    """per-case Dice / IoU of model.predict(x) > threshold (ModelValidator.validate, validate_model.py:216-248)"""
    model.eval()
    rows = []
    for batch in loader:
        x, y = batch["image"].to(device), batch["label"].to(device)
        mask = (model.predict(x) > threshold).float()
        for i in range(x.shape[0]):
            rows.append({"case_id": batch["case_id"][i], "dice": calculate_dice_score(mask[i], y[i]),
                         "iou": calculate_iou(mask[i], y[i])})
    return rows
