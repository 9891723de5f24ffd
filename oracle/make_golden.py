"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference: models/unet3d.py,
utils/losses.py, torch.optim.Adam as built at utils/trainer.py:113-117) on seeded synthetic inputs.

Run in the build container only (the GPU box has no /root/reference):   python oracle/make_golden.py
Weights are not stored: they are the reference's own initialisation under torch.manual_seed(SEED), which the
tests reproduce (same torch build -> same CPU RNG stream), so the fixtures stay small.
"""
import os
import sys

import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 0
HEAD = 8  # leading values of each gradient / parameter kept next to its norm


def synth(shape, seed):
    g = torch.Generator().manual_seed(seed)
    n, c, d, h, w = shape
    x = torch.randn(n, c, d, h, w, generator=g)
    y = (torch.rand(n, 1, d, h, w, generator=g) > 0.9).float()
    return x, y


def summarize(t):
    f = t.detach().flatten().double()
    return {"norm": f.norm().item(), "sum": f.sum().item(), "head": f[:HEAD].float().clone()}


def main():
    sys.path.insert(0, REF)
    from models.unet3d import UNet3D
    from utils.losses import BCEDiceLoss, DiceLoss
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)

    # ---- case A: one full training step, 1x5x32^3, n_classes=1 (trainer configuration, utils/trainer.py:86-89)
    torch.manual_seed(SEED)
    model = UNet3D(5, 1)
    x, y = synth((1, 5, 32, 32, 32), 1234)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-5)
    model.train()
    opt.zero_grad()
    acts = {}
    hooks = [m.register_forward_hook(lambda mod, i, o, name=name: acts.__setitem__(name, summarize(o)))
             for name, m in model.named_modules()
             if isinstance(m, (torch.nn.Conv3d, torch.nn.ConvTranspose3d, torch.nn.ReLU, torch.nn.MaxPool3d))]
    logits = model(x)
    for h_ in hooks:
        h_.remove()
    loss = BCEDiceLoss()(logits, y)
    dice = DiceLoss()(logits, y)
    loss.backward()
    grads = {k: summarize(p.grad) for k, p in model.named_parameters()}
    opt.step()
    params_after = {k: summarize(p) for k, p in model.named_parameters()}
    buffers_after = {k: (b.clone() if b.numel() <= 1024 else summarize(b)) for k, b in model.named_buffers()}
    model.eval()
    with torch.no_grad():
        logits_eval = model(x)
        probs = model.predict(x)
        mask = model.inference(x)
    torch.save({"seed": SEED, "x_seed": 1234, "shape": (1, 5, 32, 32, 32), "n_classes": 1,
                "logits_train": logits.detach(), "bce_dice": loss.item(), "dice": dice.item(), "acts": acts,
                "grads": grads, "params_after_adam": params_after, "buffers_after": buffers_after,
                "logits_eval_after_step": logits_eval, "probs": probs, "mask": mask.to(torch.uint8),
                "torch": torch.__version__}, os.path.join(OUT, "step_32cube.pt"))
    print("case A: loss", loss.item(), "dice", dice.item())

    # ---- case B: pad path (extents not multiples of 16) and n_classes=2 (run.py:130), forward only, train-mode BN
    torch.manual_seed(SEED)
    model2 = UNet3D(5, 2)
    x2, _ = synth((1, 5, 20, 36, 18), 4321)
    model2.train()
    with torch.no_grad():
        lg2 = model2(x2)
    torch.save({"seed": SEED, "x_seed": 4321, "shape": (1, 5, 20, 36, 18), "n_classes": 2, "logits_train": lg2,
                "torch": torch.__version__}, os.path.join(OUT, "fwd_pad_2class.pt"))
    print("case B: logits norm", lg2.norm().item())

    # ---- case C: loss known answers on a fixed small tensor (utils/losses.py), incl. the ValueError contract
    g = torch.Generator().manual_seed(99)
    z = torch.randn(2, 1, 4, 5, 6, generator=g) * 3
    t = (torch.rand(2, 1, 4, 5, 6, generator=g) > 0.7).float()
    zr = z.clone().requires_grad_(True)
    lb = BCEDiceLoss(0.3, 0.7)(zr, t)
    lb.backward()
    ld = DiceLoss(smooth=2.0)(z, t)
    try:
        DiceLoss()(z, t[:, :, :3])
        err = None
    except ValueError as e:
        err = str(e)
    torch.save({"z": z, "t": t, "bce_dice_03_07": lb.item(), "grad_03_07": zr.grad.clone(), "dice_smooth2": ld.item(),
                "value_error": err}, os.path.join(OUT, "losses.pt"))
    print("case C:", lb.item(), ld.item(), err)


if __name__ == "__main__":
    main()
