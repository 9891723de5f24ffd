"""Drop-in for the reference's ``utils/losses.py``: DiceLoss and BCEDiceLoss with the same constructors, attributes
and error behaviour; forward and backward each run one fused reduction / elementwise kernel pair on the B200."""
import torch
import torch.nn as nn

from . import ops
from ._lib import B200Error


class _SegLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, bce_w, dice_w, smooth):
        dev = pred.device
        z = pred.detach()
        if z.dtype != torch.float32 or not z.is_contiguous():
            z = z.float().contiguous()
        t = target.detach()
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.float().contiguous()
        ws = torch.empty(4 * 1024, device=dev, dtype=torch.float32)
        sums = torch.empty(4, device=dev, dtype=torch.float32)
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        ops.loss_fwd(z, t, bce_w, dice_w, smooth, ws, sums, loss)
        ctx.save_for_backward(z, t, sums)
        ctx.cfg = (bce_w, dice_w, smooth)
        return loss.view(())

    @staticmethod
    def backward(ctx, gout):
        z, t, sums = ctx.saved_tensors
        bce_w, dice_w, smooth = ctx.cfg
        g = gout.detach().float().reshape(1).contiguous()
        dz = torch.empty_like(z)
        ops.loss_bwd(z, t, bce_w, dice_w, smooth, sums, g, dz)
        return dz, None, None, None, None


def _seg_loss(pred, target, bce_w, dice_w, smooth):
    if pred.shape != target.shape:
        # same exception type and content as utils/losses.py:67-68 (both shapes in the message)
        raise ValueError(f"预测值和目标值的形状不匹配: pred.shape={pred.shape}, target.shape={target.shape}")
    if not pred.is_cuda:
        raise B200Error("B200 losses run on CUDA tensors only: there is no CPU path")
    with torch.autocast(device_type="cuda", enabled=False):
        return _SegLossFunction.apply(pred, target, float(bce_w), float(dice_w), float(smooth))


class DiceLoss(nn.Module):
    """1 - (2*sum(sigmoid(p)*t) + smooth) / (sum(sigmoid(p)) + sum(t) + smooth), over the whole flattened batch
    (reference utils/losses.py:16-92)"""

    def __init__(self, smooth=1.0):
        super().__init__()
        self.smooth = smooth

    def forward(self, pred, target):
        return _seg_loss(pred, target, 0.0, 1.0, self.smooth)


class BCEDiceLoss(nn.Module):
    """bce_weight * BCEWithLogits(mean) + dice_weight * DiceLoss() (reference utils/losses.py:95-152)"""

    def __init__(self, bce_weight=0.5, dice_weight=0.5):
        super().__init__()
        self.bce_weight = bce_weight
        self.dice_weight = dice_weight
        self.bce_loss = nn.BCEWithLogitsLoss()  # kept as attributes for API parity; never called
        self.dice_loss = DiceLoss()

    def forward(self, pred, target):
        return _seg_loss(pred, target, self.bce_weight, self.dice_weight, self.dice_loss.smooth)
