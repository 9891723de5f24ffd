"""CPU suite: pins the oracle (oracle/unet3d_oracle.py) against golden vectors generated from the unmodified reference
(oracle/make_golden.py), checks the drop-in boundary (state_dict keys, seed-identical init, error behaviour) and the
C-ABI library's export table.  No GPU compute here."""
import copy
import os
import re
import sys

import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unet3d_oracle as oracle  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def synth(shape, seed):
    g = torch.Generator().manual_seed(seed)
    n, c, d, h, w = shape
    x = torch.randn(n, c, d, h, w, generator=g)
    y = (torch.rand(n, 1, d, h, w, generator=g) > 0.9).float()
    return x, y


def seeded_state_dict(pkg, seed, n_classes):
    torch.manual_seed(seed)
    model = pkg.UNet3D(5, n_classes)
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def check_summary(t, ref, rtol=2e-4, atol=1e-6):
    f = t.detach().flatten().double()
    assert abs(f.norm().item() - ref["norm"]) <= rtol * abs(ref["norm"]) + atol
    assert torch.allclose(f[:ref["head"].numel()].float(), ref["head"], rtol=5e-3, atol=1e-5 + 1e-4 * ref["norm"] / max(1, f.numel()) ** 0.5)


def test_oracle_training_step_matches_reference_golden(pkg):
    gold = torch.load(os.path.join(GOLD, "step_32cube.pt"), weights_only=False)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    sd = seeded_state_dict(pkg, gold["seed"], gold["n_classes"])
    assert list(gold["grads"].keys()) == oracle.param_names(sd)
    x, y = synth(gold["shape"], gold["x_seed"])
    taps = {}
    opt_state = {}
    loss, grads, logits = oracle.train_step(sd, opt_state, x, y, lr=1e-4, weight_decay=1e-5, taps=taps)
    assert torch.allclose(logits, gold["logits_train"], rtol=1e-3, atol=1e-4)
    assert abs(loss.item() - gold["bce_dice"]) < 1e-5
    assert abs(oracle.dice_loss(logits, y).item() - gold["dice"]) < 1e-5
    # per-layer activations (reference forward hooks on Conv3d / ReLU / ConvTranspose3d)
    for name, t in taps.items():
        assert name in gold["acts"], name
        check_summary(t, gold["acts"][name], rtol=5e-4)
    for k, g in grads.items():
        check_summary(g, gold["grads"][k], rtol=2e-3, atol=1e-7)
    for k, ref in gold["params_after_adam"].items():
        check_summary(sd[k], ref, rtol=1e-5)
    for k, ref in gold["buffers_after"].items():
        if isinstance(ref, dict):
            check_summary(sd[k], ref, rtol=1e-4)
        else:
            assert torch.allclose(sd[k].double(), ref.double(), rtol=1e-4, atol=1e-6), k
    # eval mode after the step: running statistics + updated weights
    logits_eval = oracle.unet3d_forward(x, sd, training=False)
    assert torch.allclose(logits_eval, gold["logits_eval_after_step"], rtol=1e-3, atol=1e-3)
    assert torch.allclose(oracle.predict(x, sd), gold["probs"], atol=1e-4)
    mism = (oracle.inference(x, sd).to(torch.uint8) != gold["mask"]).sum().item()
    assert mism <= 2  # identical up to fp32 ties at the 0.5 threshold


def test_oracle_pad_path_two_classes(pkg):
    gold = torch.load(os.path.join(GOLD, "fwd_pad_2class.pt"), weights_only=False)
    sd = seeded_state_dict(pkg, gold["seed"], gold["n_classes"])
    x, _ = synth(gold["shape"], gold["x_seed"])
    with torch.no_grad():
        logits = oracle.unet3d_forward(x, sd, training=True)
    assert logits.shape == gold["logits_train"].shape == (1, 2, 20, 36, 18)
    assert torch.allclose(logits, gold["logits_train"], rtol=1e-3, atol=1e-3)


def test_oracle_losses_known_answers():
    gold = torch.load(os.path.join(GOLD, "losses.pt"), weights_only=False)
    z, t = gold["z"], gold["t"]
    zr = z.clone().requires_grad_(True)
    l = oracle.bce_dice_loss(zr, t, 0.3, 0.7)
    l.backward()
    assert abs(l.item() - gold["bce_dice_03_07"]) < 1e-6
    assert torch.allclose(zr.grad, gold["grad_03_07"], rtol=1e-5, atol=1e-8)
    assert torch.allclose(oracle.loss_grad_closed_form(z, t, 0.3, 0.7), gold["grad_03_07"], rtol=1e-4, atol=1e-8)
    assert abs(oracle.dice_loss(z, t, smooth=2.0).item() - gold["dice_smooth2"]) < 1e-6
    with pytest.raises(ValueError):
        oracle.dice_loss(z, t[:, :, :3])
    assert "pred.shape" in gold["value_error"] and "target.shape" in gold["value_error"]


def test_oracle_adam_matches_torch_optim():
    g = torch.Generator().manual_seed(5)
    p0 = torch.randn(1000, generator=g)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3, weight_decay=1e-5)
    p, m, v = p0.clone(), torch.zeros(1000), torch.zeros(1000)
    for step in range(1, 5):
        grad = torch.randn(1000, generator=g)
        pr.grad = grad.clone()
        opt.step()
        oracle.adam_update(p, grad, m, v, step, 1e-3, weight_decay=1e-5)
    assert torch.allclose(p, pr.detach(), rtol=1e-6, atol=1e-7)


def test_window_origins():
    assert oracle.window_origins(256, 128, 64) == [0, 64, 128]
    assert oracle.window_origins(64, 64, 64) == [0]
    assert oracle.window_origins(100, 64, 64) == [0, 36]
    assert oracle.window_origins(40, 64, 64) == [0]


# ------------------------------------------------------------------------------------------------ boundary
def test_state_dict_contract(pkg):
    torch.manual_seed(0)
    m = pkg.UNet3D(5, 1)
    sd = m.state_dict()
    assert len(sd) == 136
    assert sum(p.numel() for p in m.parameters()) == 90311361
    assert len(list(m.parameters())) == 82 and len(list(m.buffers())) == 54
    for k in ("inc.conv.0.weight", "inc.conv.1.running_var", "inc.conv.4.num_batches_tracked",
              "down4.maxpool_conv.1.conv.3.bias", "up1.up.weight", "up4.conv.conv.4.weight", "outc.bias"):
        assert k in sd
    assert sd["inc.conv.0.weight"].shape == (64, 5, 3, 3, 3)
    assert sd["up1.up.weight"].shape == (1024, 512, 2, 2, 2)
    assert sd["outc.weight"].shape == (1, 64, 1, 1, 1)
    assert (m.n_modalities, m.n_classes, m.init_features) == (5, 1, 64)
    assert pkg.UNet3D().n_classes == 2
    # reverse-forward flat order covers every parameter exactly once
    names = [n for n, _ in m.engine.ordered_params()]
    assert len(names) == 82 and set(names) == {n for n, _ in m.named_parameters()}
    assert names[0] == "outc.weight" and names[-1] == "inc.conv.0.bias"
    # base-32 variant (BASELINE cfg #5)
    m32 = pkg.UNet3D(5, 1, init_features=32)
    assert m32.state_dict()["down4.maxpool_conv.1.conv.3.weight"].shape == (512, 512, 3, 3, 3)
    with pytest.raises(ValueError):
        pkg.UNet3D(5, 1, init_features=24)


def test_no_cpu_fallback(pkg):
    m = pkg.UNet3D(5, 1, init_features=16)
    with pytest.raises(pkg.B200Error, match="no CPU path"):
        m(torch.randn(1, 5, 16, 16, 16))
    with pytest.raises(pkg.B200Error, match="no CPU path"):
        pkg.DiceLoss()(torch.randn(1, 1, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4))
    # shape mismatch is a ValueError carrying both shapes (utils/losses.py:67-68), checked before anything else
    with pytest.raises(ValueError, match="pred.shape"):
        pkg.BCEDiceLoss()(torch.randn(1, 2, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4))
    assert pkg.DiceLoss().smooth == 1.0
    l = pkg.BCEDiceLoss()
    assert (l.bce_weight, l.dice_weight) == (0.5, 0.5) and hasattr(l, "bce_loss") and hasattr(l, "dice_loss")


def test_c_abi_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "b200_unet3d.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    lib = pkg.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/b200_unet3d.h but not exported"
    from importlib import import_module
    sigs = import_module(pkg.__name__ + "._lib").SIGNATURES
    assert declared == set(sigs), declared ^ set(sigs)
    assert lib.b200_abi_version() == import_module(pkg.__name__ + "._lib").ABI_VERSION
    # development-only entry points (micro-probes, ablation switches) are not in the product header or library
    assert not any("probe" in n or "_dev_" in n for n in declared)
    assert not hasattr(lib, "b200_probe_mma") and not hasattr(lib, "b200_dev_set_ablation")


def test_flops_accounting(pkg):
    from importlib import import_module
    eng = import_module(pkg.__name__ + ".engine")
    fwd, fb = eng.total_flops_per_voxel(64, 5, 1)
    assert abs(fwd / 1e6 - 1.8245) < 2e-3 and abs(fb / 1e6 - 5.4563) < 5e-3
    fwd32, fb32 = eng.total_flops_per_voxel(32, 5, 1)
    assert abs(fwd32 / 1e6 - 0.4605) < 2e-3 and abs(fb32 / 1e6 - 1.3728) < 5e-3


def test_resample_restatement_known_answers():
    """hand-computed values of the ITK index mapping (i * in/out, clamp, outside -> 0) on a ramp"""
    import numpy as np
    x = np.arange(4, dtype=np.float32).reshape(1, 1, 4) * 10.0         # W axis ramp 0,10,20,30
    up = oracle.resample3d_itk(x, (1, 1, 8))                            # continuous index 0,.5,...,3.5
    assert np.allclose(up[0, 0], [0, 5, 10, 15, 20, 25, 30, 0])         # 3.5 >= 4 - 0.5 -> outside -> 0
    down = oracle.resample3d_itk(x, (1, 1, 2))                          # index 0, 2
    assert np.allclose(down[0, 0], [0, 20])
    near = oracle.resample3d_itk(x, (1, 1, 8), nearest=True)            # round half up: 0,1,1,2,2,3,3,outside
    assert np.allclose(near[0, 0], [0, 10, 10, 20, 20, 30, 30, 0])
    third = oracle.resample3d_itk(x, (1, 1, 3))                         # index 0, 4/3, 8/3
    assert np.allclose(third[0, 0], [0, 40 / 3, 80 / 3], atol=1e-5)
    assert np.array_equal(oracle.resample3d_itk(x, (1, 1, 4)), x)
    assert np.array_equal(oracle.resample3d_itk(x, (1, 1, 2), nearest=True, binarize=True)[0, 0], [0, 1])


def test_metric_and_normalise_restatements():
    import numpy as np
    p = np.array([1, 1, 0, 0, 1], dtype=np.float32)
    t = np.array([1, 0, 0, 1, 1], dtype=np.float32)
    d, j = oracle.hard_dice_iou(p, t)
    assert abs(d - (4 + 1e-8) / (6 + 1e-8)) < 1e-12 and abs(j - (2 + 1e-8) / (4 + 1e-8)) < 1e-12
    d0, j0 = oracle.hard_dice_iou(np.zeros(4), np.zeros(4))             # both empty -> eps/eps = 1
    assert d0 == 1.0 and j0 == 1.0
    img = np.stack([np.array([[[2.0, 4.0, 6.0]]]), np.full((1, 1, 3), 7.0)]).astype(np.float32)
    out = oracle.minmax_normalize(img)
    assert np.allclose(out[0], [[[0, 0.5, 1]]]) and np.array_equal(out[1], np.zeros((1, 1, 3)))
