"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU (or any-device) fp32 restatement, in plain torch functional ops, of the reference hot path:
  * UNet3D graph                  — reference models/unet3d.py:27-40, 78-83, 120-158, 247-296
  * predict / inference           — reference models/unet3d.py:298-344
  * DiceLoss / BCEDiceLoss        — reference utils/losses.py:44-92, 107-152
  * one training step             — reference utils/trainer.py:177-195 (zero_grad, forward, loss, backward, Adam.step)
  * Adam update                   — torch.optim.Adam as constructed at utils/trainer.py:113-117
  * data-parallel step (new capability): per-shard forward/backward, mean of gradients, one Adam step
  * sliding-window inference (new capability, SURVEY.md 8d cfg #4): uniform averaging of window logits

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Pinning: the reference ships no golden vectors or assertions for this path.  This restatement is
pinned against the *reference itself*: oracle/make_golden.py imports /root/reference (models/unet3d.py,
utils/losses.py) in the build container and writes tests/golden/*.pt; tests/test_oracle_cpu.py checks this file
against those vectors.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ---- optional "bf16 storage" mode -----------------------------------------------------------------------------
# The same graph with every tensor that the B200 engine keeps in HBM as bf16 rounded to bf16 at that point (forward
# value and, through the straight-through backward, its gradient), all arithmetic still fp32.  With store=None (the
# default) nothing is rounded: that is the pinned fp32 oracle.  The per-layer 2e-2 bound of north_star ("fp32-
# accumulated bf16 outputs and gradients") is checked against this mode; end-to-end fp32 comparisons are reported too.
class _StoreBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


class _OperandBF16(torch.autograd.Function):
    """weights are fp32 masters read through a bf16 shadow; their gradients stay fp32"""

    @staticmethod
    def forward(ctx, w):
        return w.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


def store_bf16(x):
    return _StoreBF16.apply(x)


def _ident(x):
    return x


def _double_conv(x, sd, prefix, training, taps, store=None):
    """Conv3d(3,p1)+BN+ReLU twice.  sd: state-dict-like mapping; running stats are updated in place when training."""
    wop = _OperandBF16.apply if store is not None else _ident
    store = store or _ident
    for conv_i, bn_i in ((0, 1), (3, 4)):
        x = store(F.conv3d(x, wop(sd[f"{prefix}.{conv_i}.weight"]), sd[f"{prefix}.{conv_i}.bias"], padding=1))
        if taps is not None:
            taps[f"{prefix}.{conv_i}"] = x
        rm, rv = sd[f"{prefix}.{bn_i}.running_mean"], sd[f"{prefix}.{bn_i}.running_var"]
        x = F.batch_norm(x, rm, rv, sd[f"{prefix}.{bn_i}.weight"], sd[f"{prefix}.{bn_i}.bias"], training,
                         BN_MOMENTUM, BN_EPS)
        if training:
            sd[f"{prefix}.{bn_i}.num_batches_tracked"] += 1
        x = store(F.relu(x))
        if taps is not None:
            taps[f"{prefix}.{bn_i + 1}"] = x
    return x


def unet3d_forward(x, sd, training=False, taps=None, store=None):
    """logits = UNet3D(x).  `sd` has the reference's 136 state_dict keys; `taps` (dict) collects per-layer outputs
    keyed by the reference module path (conv outputs at '.0'/'.3', post-ReLU at '.2'/'.5').  store: see above."""
    wop = _OperandBF16.apply if store is not None else _ident
    st = store or _ident
    skips = []
    h = _double_conv(st(x), sd, "inc.conv", training, taps, store)
    skips.append(h)
    for k in (1, 2, 3, 4):
        h = F.max_pool3d(h, 2)
        h = _double_conv(h, sd, f"down{k}.maxpool_conv.1.conv", training, taps, store)
        if k < 4:
            skips.append(h)
    for j in (1, 2, 3, 4):
        skip = skips[4 - j]
        h = st(F.conv_transpose3d(h, wop(sd[f"up{j}.up.weight"]), sd[f"up{j}.up.bias"], stride=2))
        if taps is not None:
            taps[f"up{j}.up"] = h
        dz, dy, dx = (skip.shape[2] - h.shape[2], skip.shape[3] - h.shape[3], skip.shape[4] - h.shape[4])
        h = F.pad(h, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2, dz // 2, dz - dz // 2])
        h = torch.cat([skip, h], dim=1)
        h = _double_conv(h, sd, f"up{j}.conv.conv", training, taps, store)
    return F.conv3d(h, sd["outc.weight"], sd["outc.bias"])


def predict(x, sd):
    with torch.no_grad():
        return torch.sigmoid(unet3d_forward(x, sd, training=False))


def inference(x, sd, threshold=0.5):
    return (predict(x, sd) > threshold).float()


def dice_loss(pred, target, smooth=1.0):
    if pred.shape != target.shape:
        raise ValueError(f"shape mismatch: pred.shape={pred.shape}, target.shape={target.shape}")
    p = torch.sigmoid(pred).reshape(-1)
    t = target.reshape(-1)
    inter = (p * t).sum()
    return 1 - (2.0 * inter + smooth) / (p.sum() + t.sum() + smooth)


def bce_dice_loss(pred, target, bce_weight=0.5, dice_weight=0.5, smooth=1.0):
    bce = F.binary_cross_entropy_with_logits(pred, target)
    return bce_weight * bce + dice_weight * dice_loss(pred, target, smooth)


def loss_grad_closed_form(pred, target, bce_weight=0.5, dice_weight=0.5, smooth=1.0):
    """dL/dz in closed form — cross-check of autograd in the CPU tests."""
    s = torch.sigmoid(pred)
    n = pred.numel()
    inter, p_sum, t_sum = (s * target).sum(), s.sum(), target.sum()
    b = p_sum + t_sum + smooth
    a = 2 * inter + smooth
    return bce_weight * (s - target) / n + dice_weight * s * (1 - s) * (-2 * target / b + a / (b * b))


def param_names(sd):
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


def adam_update(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """in-place Adam with coupled L2 decay, as torch.optim.Adam(amsgrad=False)"""
    b1, b2 = betas
    if weight_decay:
        g = g + weight_decay * p
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def train_step(sd, opt_state, x, y, lr=1e-4, weight_decay=1e-5, loss="bce_dice", taps=None):
    """zero_grad -> forward -> loss -> backward -> Adam.step on `sd` in place.  Returns (loss, grads, logits)."""
    names = param_names(sd)
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    work = dict(sd)
    work.update(leaves)
    logits = unet3d_forward(x, work, training=True, taps=taps)
    lval = bce_dice_loss(logits, y) if loss == "bce_dice" else dice_loss(logits, y)
    grads = dict(zip(names, torch.autograd.grad(lval, [leaves[k] for k in names])))
    apply_adam(sd, opt_state, grads, lr, weight_decay)
    return lval.detach(), grads, logits.detach()


def apply_adam(sd, opt_state, grads, lr=1e-4, weight_decay=1e-5):
    opt_state["step"] = opt_state.get("step", 0) + 1
    with torch.no_grad():
        for k, g in grads.items():
            m = opt_state.setdefault(("m", k), torch.zeros_like(sd[k]))
            v = opt_state.setdefault(("v", k), torch.zeros_like(sd[k]))
            adam_update(sd[k], g, m, v, opt_state["step"], lr, weight_decay=weight_decay)


def dp_train_step(sd, opt_state, x_shards, y_shards, lr=1e-4, weight_decay=1e-5, loss="bce_dice", store=None):
    """Data-parallel semantics: every rank runs forward/backward on its shard from the same weights
    with rank-local BatchNorm statistics and rank-local loss; parameter gradients are averaged; one Adam step.
    Running statistics follow rank 0 (DDP broadcast_buffers behaviour).  Returns (mean loss, averaged grads)."""
    names = param_names(sd)
    total, losses = None, []
    sd_rank0 = None
    for r, (xs, ys) in enumerate(zip(x_shards, y_shards)):
        work = {k: (v.clone() if not k in names else v) for k, v in sd.items()}
        leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
        work.update(leaves)
        logits = unet3d_forward(xs, work, training=True, store=store)
        lval = bce_dice_loss(logits, ys) if loss == "bce_dice" else dice_loss(logits, ys)
        gr = torch.autograd.grad(lval, [leaves[k] for k in names])
        total = list(gr) if total is None else [a + b for a, b in zip(total, gr)]
        losses.append(lval.detach())
        if r == 0:
            sd_rank0 = work
    world = len(x_shards)
    grads = {k: g / world for k, g in zip(names, total)}
    for k in sd:
        if k not in names:
            sd[k].copy_(sd_rank0[k])
    apply_adam(sd, opt_state, grads, lr, weight_decay)
    return torch.stack(losses).mean(), grads


def window_origins(extent, window, stride):
    """start offsets of sliding windows along one axis: regular stride, last window flush with the end"""
    if extent <= window:
        return [0]
    o = list(range(0, extent - window + 1, stride))
    if o[-1] != extent - window:
        o.append(extent - window)
    return o


def sliding_window_logits(x, sd, window, stride):
    """uniform-weight average of per-window logits, windows visited in (d, h, w) order (SURVEY.md 8d cfg #4)"""
    n, _, D, H, W = x.shape
    wd, wh, ww = (min(window[0], D), min(window[1], H), min(window[2], W))
    out = None
    cnt = torch.zeros(1, 1, D, H, W, dtype=x.dtype, device=x.device)
    with torch.no_grad():
        for d0 in window_origins(D, wd, stride[0]):
            for h0 in window_origins(H, wh, stride[1]):
                for w0 in window_origins(W, ww, stride[2]):
                    lg = unet3d_forward(x[:, :, d0:d0 + wd, h0:h0 + wh, w0:w0 + ww], sd, training=False)
                    if out is None:
                        out = torch.zeros(n, lg.shape[1], D, H, W, dtype=x.dtype, device=x.device)
                    out[:, :, d0:d0 + wd, h0:h0 + wh, w0:w0 + ww] += lg
                    cnt[:, :, d0:d0 + wd, h0:h0 + wh, w0:w0 + ww] += 1
    return out / cnt


# ------------------------------------------------------------------------------------------------ "next" rows (8f)
def resample3d_itk(x, size, nearest=False, binarize=False):
    """numpy restatement of the reference's resample-to-target (script/data_loader.py:257-279 image, :389-406 label):
    sitk.ResampleImageFilter with identity transform, the input's origin/direction and output spacing
    in_size * spacing / out_size.  PARITY UNPINNED: SimpleITK (an un-vendored dependency, requirements.txt) is not
    installable offline, so this follows ITK's published algorithm — ResampleImageFilter maps output index i to the
    continuous input index i * in/out; IsInsideBuffer is index < size - 0.5 (else DefaultPixelValue 0);
    LinearInterpolateImageFunction takes floor + fraction and clamps the upper neighbour to the last index;
    NearestNeighborInterpolateImageFunction rounds half up.  x: (..., D, H, W) float array."""
    import numpy as np
    x = np.asarray(x, dtype=np.float32)
    di, hi, wi = x.shape[-3:]
    do, ho, wo = size
    out = np.zeros(x.shape[:-3] + (do, ho, wo), dtype=np.float32)

    def axis(n_in, n_out):
        c = np.arange(n_out, dtype=np.float64) * (n_in / n_out)
        inside = c < n_in - 0.5
        if nearest:
            i0 = np.minimum(np.floor(c + 0.5).astype(np.int64), n_in - 1)
            return inside, i0, i0, np.zeros(n_out, dtype=np.float32)
        i0 = np.floor(c).astype(np.int64)
        return inside, np.minimum(i0, n_in - 1), np.minimum(i0 + 1, n_in - 1), (c - i0).astype(np.float32)

    ind, d0, d1, fd = axis(di, do)
    inh, h0, h1, fh = axis(hi, ho)
    inw, w0, w1, fw = axis(wi, wo)
    fd, fh, fw = fd[:, None, None], fh[None, :, None], fw[None, None, :]

    def g(a, b, c):
        return x[..., a[:, None, None], b[None, :, None], c[None, None, :]]

    if nearest:
        v = g(d0, h0, w0)
    else:
        c00 = g(d0, h0, w0) + fw * (g(d0, h0, w1) - g(d0, h0, w0))
        c01 = g(d0, h1, w0) + fw * (g(d0, h1, w1) - g(d0, h1, w0))
        c10 = g(d1, h0, w0) + fw * (g(d1, h0, w1) - g(d1, h0, w0))
        c11 = g(d1, h1, w0) + fw * (g(d1, h1, w1) - g(d1, h1, w0))
        c0 = c00 + fh * (c01 - c00)
        c1 = c10 + fh * (c11 - c10)
        v = c0 + fd * (c1 - c0)
    inside = ind[:, None, None] & inh[None, :, None] & inw[None, None, :]
    out[...] = np.where(inside, v, 0.0)
    if binarize:
        out = (out > 0).astype(np.float32)
    return out


def minmax_normalize(image):
    """per-modality (x - min) / (max - min), constant volume -> zeros (script/predict.py:69-75)"""
    import numpy as np
    out = np.zeros_like(image, dtype=np.float32)
    for m in range(image.shape[0]):
        lo, hi = np.float32(image[m].min()), np.float32(image[m].max())
        if hi > lo:
            out[m] = (image[m].astype(np.float32) - lo) / (hi - lo)
    return out


def hard_dice_iou(pred_mask, target_mask, eps=1e-8):
    """script/validate_model.py:24-95 on float 0/1 masks, in double"""
    import numpy as np
    p = np.asarray(pred_mask, dtype=np.float64).reshape(-1)
    t = np.asarray(target_mask, dtype=np.float64).reshape(-1)
    inter = float((p * t).sum())
    dice = (2 * inter + eps) / (p.sum() + t.sum() + eps)
    iou = (inter + eps) / (p.sum() + t.sum() - inter + eps)
    return float(dice), float(iou)
