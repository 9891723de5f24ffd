"""Full-size parity harness shared by tests/test_fullsize_gpu.py and tools/parity_fullsize.py.

One training step (forward, BCE+Dice, backward) of the B200 engine at a BASELINE.json shape is compared, on the same
GPU, with the oracle (oracle/unet3d_oracle.py = the reference graph as stock torch fp32 ops, TF32 off) in three forms:

  fp32      : nothing rounded — the reference's arithmetic (models/unet3d.py:247-296, utils/losses.py:107-152)
  storage   : the same graph with every tensor the engine keeps in HBM as bf16 rounded to bf16 at that point (fp32
              arithmetic everywhere) — what the engine computes up to accumulation order
  autocast  : torch's own bf16 autocast of the same graph (the reference's AMP loop, train_bph_optimized.py:269)

Reported per layer: relative L2 of every kept activation (raw conv outputs, post-ReLU outputs, transposed-conv outputs)
and of all parameter gradients.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unet3d_oracle as oracle  # noqa: E402


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    den = b.norm().item()
    return a.norm().item() if den == 0.0 else ((a - b).norm() / den).item()


def is_dead_bias(name):
    """conv bias feeding a train-mode BatchNorm: its gradient is exactly zero in exact arithmetic"""
    return name.endswith(".bias") and (".conv.0." in name or ".conv.3." in name)


def synth_batch(batch, size, seed, device, zero_fill=False):
    """synthetic volumes as SURVEY.md 8(d): randn image, Bernoulli(0.1) label; zero_fill: each of modalities 1..4 is a
    whole channel of zeros with p = 0.2 per sample (script/data_loader.py:320-322), ADC always present"""
    g = torch.Generator().manual_seed(seed)
    d, h, w = size
    x = torch.randn(batch, 5, d, h, w, generator=g)
    y = (torch.rand(batch, 1, d, h, w, generator=g) < 0.1).float()
    if zero_fill:
        present = torch.rand(batch, 5, generator=g) >= 0.2
        present[:, 0] = True
        if bool(present.all()):
            present[0, 2] = False  # make sure the case is exercised
        x = x * present[:, :, None, None, None].float()
    return x.to(device), y.to(device)


def _oracle_pass(sd, x, y, store=None, autocast=False, engine_taps=None):
    """returns (loss, grads, logits, {layer: rel-L2 of the engine's activation vs this oracle's})"""
    names = oracle.param_names(sd)
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    work = {k: v.clone() for k, v in sd.items()}
    work.update(leaves)
    taps = {} if engine_taps is not None else None
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits = oracle.unet3d_forward(x, work, training=True, taps=taps, store=store)
    acts = {}
    if taps is not None:
        for name, view in engine_taps.items():
            ref = taps.pop(name).detach()
            ours = view.to_ncdhw()
            if ours.shape != ref.shape:   # pad path: the engine's tensor holds the F.pad border as well
                dd, dh, dw = (ours.shape[i] - ref.shape[i] for i in (2, 3, 4))
                ours = ours[:, :, dd // 2:dd // 2 + ref.shape[2], dh // 2:dh // 2 + ref.shape[3],
                            dw // 2:dw // 2 + ref.shape[4]]
            acts[name] = rel_l2(ours, ref)
            del ours, ref
        taps.clear()
    loss = oracle.bce_dice_loss(logits.float(), y)
    grads = dict(zip(names, torch.autograd.grad(loss, [leaves[k] for k in names])))
    return loss.detach(), grads, logits.detach().float(), acts, work


def train_step_parity(pkg, device, batch, size, base=64, n_classes=1, zero_fill=False, seed=0, x_seed=1234,
                      with_autocast=True):
    """runs the comparison; returns a dict with 'loss', 'logits', 'acts', 'grads' sub-dicts (see write_report)"""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(seed)
    model = pkg.UNet3D(5, n_classes, init_features=base)
    sd = {k: v.detach().clone().to(device) for k, v in model.state_dict().items()}
    model = model.to(device).train()
    x, y = synth_batch(batch, size, x_seed, device, zero_fill)
    if n_classes > 1:
        y = y.expand(-1, n_classes, -1, -1, -1).contiguous()
    eng = model.engine
    eng.keep_tape = True
    model.zero_grad()
    logits = model(x)
    engine_taps = dict(eng.layer_outputs(eng.last_tape))
    eng.keep_tape, eng.last_tape = False, None
    loss = pkg.BCEDiceLoss()(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    ours_grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    ours_logits = logits.detach()
    ours_bn = {k: v.detach().clone() for k, v in model.state_dict().items() if "running" in k}

    out = {"config": {"batch": batch, "size": list(size), "base": base, "n_classes": n_classes,
                      "zero_fill": zero_fill}, "loss": {"ours": loss.item()}, "logits": {}, "acts": {}, "grads": {},
           "bn_buffers": {}}
    forms = [("fp32", None, False), ("storage", oracle.store_bf16, False)]
    if with_autocast:
        forms.append(("autocast", None, True))
    ref_grads = {}
    for form, store, ac in forms:
        o_loss, o_grads, o_logits, acts, work = _oracle_pass(sd, x, y, store=store, autocast=ac,
                                                             engine_taps=engine_taps if form != "autocast" else None)
        out["loss"][form] = o_loss.item()
        if form == "fp32":
            fp32_logits = o_logits
            ref_grads = o_grads
            out["bn_buffers"] = {k: rel_l2(ours_bn[k], work[k]) for k in ours_bn}
            out["logits"]["ours_vs_fp32"] = rel_l2(ours_logits, o_logits)
            out["mask_mismatch_fp32"] = float(((ours_logits > 0) != (o_logits > 0)).float().mean().item())
            sure = o_logits.abs() > 0.05 * o_logits.abs().mean()
            out["mask_mismatch_fp32_sure"] = float(((ours_logits > 0) != (o_logits > 0))[sure].float().sum().item())
        elif form == "storage":
            out["logits"]["ours_vs_storage"] = rel_l2(ours_logits, o_logits)
            out["logits"]["storage_vs_fp32"] = rel_l2(o_logits, fp32_logits)
        else:
            out["logits"]["autocast_vs_fp32"] = rel_l2(o_logits, fp32_logits)
        if acts:
            out["acts"][form] = acts
        for n, g in ours_grads.items():
            if is_dead_bias(n):
                continue
            row = out["grads"].setdefault(n, {})
            if form == "fp32":
                row["ours_vs_fp32"] = rel_l2(g, o_grads[n])
            elif form == "storage":
                row["ours_vs_storage"] = rel_l2(g, o_grads[n])
                row["storage_vs_fp32"] = rel_l2(o_grads[n], ref_grads[n])
            else:
                row["autocast_vs_fp32"] = rel_l2(o_grads[n], ref_grads[n])
        del o_grads, o_logits, work
        torch.cuda.empty_cache()
    dead = {}
    for n, g in ours_grads.items():
        if is_dead_bias(n):
            wn = ref_grads[n.replace(".bias", ".weight")].norm().item()
            dead[n] = g.norm().item() / max(wn, 1e-12)
    out["dead_bias_ratio_max"] = max(dead.values()) if dead else 0.0
    return out


def write_report(res, path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    c = res["config"]
    with open(path, "w") as f:
        f.write(f"# one training step, batch {c['batch']} x 5 x {c['size']} base {c['base']} classes {c['n_classes']}"
                f"{' zero_fill' if c['zero_fill'] else ''}: B200 engine vs the oracle on the same GPU (rel-L2)\n")
        f.write("# forms: fp32 = reference arithmetic; storage = fp32 arithmetic, tensors rounded to bf16 where the "
                "engine stores bf16; autocast = torch bf16 autocast\n")
        f.write("loss " + " ".join(f"{k}={v:.6f}" for k, v in res["loss"].items()) + "\n")
        f.write("logits " + " ".join(f"{k}={v:.3e}" for k, v in res["logits"].items()) + "\n")
        f.write(f"mask_mismatch_fraction_vs_fp32 {res['mask_mismatch_fp32']:.3e} "
                f"(where |fp32 logit| > 5% of mean: {res['mask_mismatch_fp32_sure']:.0f} voxels)\n")
        f.write(f"bn_running_buffers_max_rel_l2 {max(res['bn_buffers'].values()):.3e}\n")
        f.write(f"dead_bias_grad_over_weight_grad_max {res['dead_bias_ratio_max']:.3e}\n")
        f.write("# activation  ours_vs_fp32  ours_vs_storage\n")
        for name in res["acts"].get("fp32", {}):
            f.write(f"act {name} {res['acts']['fp32'][name]:.3e} {res['acts'].get('storage', {}).get(name, float('nan')):.3e}\n")
        f.write("# gradient  ours_vs_fp32  ours_vs_storage  storage_vs_fp32  autocast_vs_fp32\n")
        for name, row in res["grads"].items():
            f.write(f"grad {name} {row.get('ours_vs_fp32', float('nan')):.3e} {row.get('ours_vs_storage', float('nan')):.3e} "
                    f"{row.get('storage_vs_fp32', float('nan')):.3e} {row.get('autocast_vs_fp32', float('nan')):.3e}\n")


def summarize(res):
    g = res["grads"]
    mx = lambda key: max((r[key] for r in g.values() if key in r), default=float("nan"))  # noqa: E731
    return {"loss_abs_err": abs(res["loss"]["ours"] - res["loss"]["fp32"]),
            "logits_vs_fp32": res["logits"]["ours_vs_fp32"],
            "act_max_vs_fp32": max(res["acts"]["fp32"].values()),
            "act_max_vs_storage": max(res["acts"]["storage"].values()),
            "grad_max_vs_fp32": mx("ours_vs_fp32"), "grad_max_vs_storage": mx("ours_vs_storage"),
            "grad_max_storage_vs_fp32": mx("storage_vs_fp32"), "grad_max_autocast_vs_fp32": mx("autocast_vs_fp32")}
