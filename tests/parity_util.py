"""Full-size parity harness shared by tests/test_fullsize_gpu.py and tools/parity_fullsize.py.

One training step (forward, BCE+Dice, backward) of the B200 engine at a BASELINE.json shape is compared, on the same
GPU, with the oracle (oracle/unet3d_oracle.py = the reference graph as stock torch fp32 ops, TF32 off) in three forms:

  fp32      : nothing rounded — the reference's arithmetic (models/unet3d.py:247-296, utils/losses.py:107-152)
  storage   : the same graph with every tensor the engine keeps in HBM as bf16 rounded to bf16 at that point (fp32
              arithmetic everywhere) — what the engine computes up to accumulation order
  autocast  : torch's own bf16 autocast of the same graph (the reference's AMP loop, train_bph_optimized.py:269)
  storage~  : the storage form again with the input and every weight multiplied by (1 + 2^-21 u), u ~ U(-1, 1) — a
              perturbation of the size of ONE fp32 rounding.  How far this moves the storage form's own gradients is the
              yardstick for "equal up to accumulation order": a network of 23 BatchNorm'd layers with bf16-rounded
              tensors amplifies fp32-level differences (a different but equally valid summation order flips bf16
              roundings, those flip ReLU masks downstream), so two correct bf16-storage implementations differ end
              to end by this much.

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


def _perturbed(t, gen):
    u = torch.rand(t.shape, generator=gen, device=t.device, dtype=torch.float32) * 2 - 1
    return t * (1 + u * 2.0 ** -21)


def _oracle_pass(sd, x, y, store=None, autocast=False, engine_taps=None, perturb_seed=None, keep_taps=False,
                 ref_taps=None):
    """returns (loss, grads, logits, {layer: rel-L2 of the engine's activation vs this oracle's}, state, taps);
    with ref_taps the per-layer numbers are this pass's activations against ref_taps instead"""
    names = oracle.param_names(sd)
    if perturb_seed is not None:
        gen = torch.Generator(device=x.device).manual_seed(perturb_seed)
        x = _perturbed(x, gen)
        sd = {k: (_perturbed(v, gen) if k in names else v) for k, v in sd.items()}
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    work = {k: v.clone() for k, v in sd.items()}
    work.update(leaves)
    taps = {} if (engine_taps is not None or ref_taps is not None) else None
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits = oracle.unet3d_forward(x, work, training=True, taps=taps, store=store)
    acts = {}
    kept = None
    if ref_taps is not None:
        acts = {name: rel_l2(taps[name].detach(), ref) for name, ref in ref_taps.items()}
        taps.clear()
    elif taps is not None:
        if keep_taps:
            kept = {k: v.detach() for k, v in taps.items()}
        for name, view in engine_taps.items():
            if name.endswith(".in"):
                continue
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
    return loss.detach(), grads, logits.detach().float(), acts, work, kept


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
    forms = [("fp32", None, False), ("storage", oracle.store_bf16, False), ("storage~", oracle.store_bf16, False)]
    if with_autocast:
        forms.append(("autocast", None, True))
    ref_grads, st_grads, st_taps = {}, {}, None
    for form, store, ac in forms:
        o_loss, o_grads, o_logits, acts, work, kept = _oracle_pass(
            sd, x, y, store=store, autocast=ac, engine_taps=engine_taps if form in ("fp32", "storage") else None,
            perturb_seed=77 if form == "storage~" else None, keep_taps=(form == "storage"),
            ref_taps=st_taps if form == "storage~" else None)
        if form == "storage":
            st_taps = kept
        elif form == "storage~":
            st_taps = None
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
            st_logits, st_grads = o_logits, o_grads
            out["logits"]["ours_vs_storage"] = rel_l2(ours_logits, o_logits)
            out["logits"]["storage_vs_fp32"] = rel_l2(o_logits, fp32_logits)
        elif form == "storage~":
            out["logits"]["storage~_vs_storage"] = rel_l2(o_logits, st_logits)
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
            elif form == "storage~":
                row["storage~_vs_storage"] = rel_l2(o_grads[n], st_grads[n])
            else:
                row["autocast_vs_fp32"] = rel_l2(o_grads[n], ref_grads[n])
        if form != "storage":
            del o_grads, o_logits
        del work
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
        f.write("# activation  ours_vs_fp32  ours_vs_storage  storage~_vs_storage\n")
        for name in res["acts"].get("fp32", {}):
            f.write(f"act {name} {res['acts']['fp32'][name]:.3e} "
                    f"{res['acts'].get('storage', {}).get(name, float('nan')):.3e} "
                    f"{res['acts'].get('storage~', {}).get(name, float('nan')):.3e}\n")
        f.write("# gradient  ours_vs_fp32  ours_vs_storage  storage~_vs_storage  storage_vs_fp32  autocast_vs_fp32\n")
        nan = float("nan")
        for name, row in res["grads"].items():
            f.write(f"grad {name} {row.get('ours_vs_fp32', nan):.3e} {row.get('ours_vs_storage', nan):.3e} "
                    f"{row.get('storage~_vs_storage', nan):.3e} {row.get('storage_vs_fp32', nan):.3e} "
                    f"{row.get('autocast_vs_fp32', nan):.3e}\n")


def summarize(res):
    g = res["grads"]
    mx = lambda key: max((r[key] for r in g.values() if key in r), default=float("nan"))  # noqa: E731
    return {"loss_abs_err": abs(res["loss"]["ours"] - res["loss"]["fp32"]),
            "logits_vs_fp32": res["logits"]["ours_vs_fp32"],
            "act_max_vs_fp32": max(res["acts"]["fp32"].values()),
            "act_max_vs_storage": max(res["acts"]["storage"].values()),
            "act_max_storage~_vs_storage": max(res["acts"]["storage~"].values()),
            "grad_max_vs_fp32": mx("ours_vs_fp32"), "grad_max_vs_storage": mx("ours_vs_storage"),
            "grad_max_storage~_vs_storage": mx("storage~_vs_storage"),
            "grad_max_storage_vs_fp32": mx("storage_vs_fp32"), "grad_max_autocast_vs_fp32": mx("autocast_vs_fp32")}


# ------------------------------------------------------------------------------------------------ layer by layer
def layerwise_check(pkg, device, batch, size, base=64, n_classes=1, zero_fill=False, seed=0, x_seed=1234):
    """Every layer of one training step checked IN ISOLATION at the given (BASELINE) shape: the engine's own input
    tensors of each op (bf16 activations / gradients exactly as it stored them, bf16-rounded weights) are fed to the
    torch fp32 op the reference calls at that point (F.conv3d, F.batch_norm + relu, F.max_pool3d, F.conv_transpose3d,
    the 1x1x1 head, BCE+Dice) and to torch autograd of it; the engine's outputs are compared with the result.  This is
    the north_star bound proper — "fp32-accumulated bf16 outputs and gradients within 2e-2 relative L2 per layer" —
    with identical inputs, forward and backward, at full size.  Returns [(op, quantity, rel-L2)], exact-match rows
    carry the number of mismatching elements instead."""
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(seed)
    model = pkg.UNet3D(5, n_classes, init_features=base).to(device).train()
    x, y = synth_batch(batch, size, x_seed, device, zero_fill)
    if n_classes > 1:
        y = y.expand(-1, n_classes, -1, -1, -1).contiguous()
    eng = model.engine
    eng.keep_tape, eng.grad_taps = True, {}
    model.zero_grad()
    logits = model(x)
    tape = eng.last_tape
    fw = dict(eng.layer_outputs(tape))
    skips = [pkg.ops.ActView(tape.cats[k], 0, tape.cats[k].shape[-1] // 2) for k in range(4)]
    pooled = list(tape.pooled)
    last = tape.last
    loss = pkg.BCEDiceLoss()(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    gt = eng.grad_taps
    eng.keep_tape, eng.last_tape, eng.grad_taps = False, None, None
    pg = {n: p.grad.detach() for n, p in model.named_parameters()}
    par = {n: p.detach() for n, p in model.named_parameters()}
    rows = []
    f32 = lambda v: v.to_ncdhw()                                   # noqa: E731  (exact bf16 values as fp32 NCDHW)
    rb = lambda w: w.to(torch.bfloat16).to(torch.float32)          # noqa: E731  (the operand the conv kernels read)

    def add(op, what, a, b):
        rows.append((op, what, rel_l2(a, b)))

    # ---- loss and head
    zl = logits.detach().clone().requires_grad_(True)
    ref_loss = oracle.bce_dice_loss(zl, y)
    (dz_ref,) = torch.autograd.grad(ref_loss, [zl])
    rows.append(("loss", "value_abs_err", abs(loss.item() - ref_loss.item())))
    hx = f32(last).requires_grad_(True)
    hw, hb = par["outc.weight"].clone().requires_grad_(True), par["outc.bias"].clone().requires_grad_(True)
    href = F.conv3d(hx, hw, hb)
    add("outc", "fwd", logits.detach(), href.detach())
    gx, gw, gb = torch.autograd.grad(href, [hx, hw, hb], dz_ref)
    add("outc", "dx", f32(gt["head"]["dx"]), gx)
    add("outc", "dw", pg["outc.weight"], gw)
    add("outc", "db", pg["outc.bias"], gb)
    del hx, href, gx, zl

    # ---- DoubleConv blocks
    prefix = {"inc": "inc.conv"}
    for k in (1, 2, 3, 4):
        prefix[f"down{k}"] = f"down{k}.maxpool_conv.1.conv"
    for j in (1, 2, 3, 4):
        prefix[f"up{j}"] = f"up{j}.conv.conv"
    for name, pre in prefix.items():
        g = gt[name]
        xin = rb(x) if name == "inc" else f32(fw[f"{pre}.in"])
        for ci, bi, yk, ak, dyk, dxk in ((0, 1, "0", "2", "dy1", "dx"), (3, 4, "3", "5", "dy2", "da1")):
            w = rb(par[f"{pre}.{ci}.weight"]).requires_grad_(True)
            a_in = (xin if ci == 0 else f32(fw[f"{pre}.2"])).requires_grad_(name != "inc" or ci == 3)
            yref = F.conv3d(a_in, w, par[f"{pre}.{ci}.bias"], padding=1)
            add(f"{pre}.{ci}", "fwd", f32(fw[f"{pre}.{yk}"]), yref.detach())
            dy = f32(g[dyk])
            wanted = [w] + ([a_in] if a_in.requires_grad else [])
            gr = torch.autograd.grad(yref, wanted, dy)
            add(f"{pre}.{ci}", "dw", pg[f"{pre}.{ci}.weight"], gr[0])
            if a_in.requires_grad:
                got = g[dxk]
                if ci == 0 and name.startswith("up"):
                    # the lower (skip) half of this buffer receives the pool gradient later in the same backward:
                    # its value at this point was kept by the pool tap
                    kk = 4 - int(name[2])
                    c = got.shape[-1] // 2
                    ours = torch.cat([gt[f"pool{kk + 1}"]["dskip_before"], f32(got)[:, c:]], dim=1)
                else:
                    ours = f32(got)
                add(f"{pre}.{ci}", "dx", ours, gr[1])
                del ours
            del yref, gr, a_in, w
            # BatchNorm (batch statistics) + ReLU on the engine's own conv output
            yv = f32(fw[f"{pre}.{yk}"]).requires_grad_(True)
            gam = par[f"{pre}.{bi}.weight"].clone().requires_grad_(True)
            bet = par[f"{pre}.{bi}.bias"].clone().requires_grad_(True)
            zref = F.batch_norm(yv, None, None, gam, bet, True, 0.1, 1e-5)
            a_ours = f32(fw[f"{pre}.{ak}"])
            add(f"{pre}.{bi}", "fwd", a_ours, F.relu(zref).detach())
            # backward through the ReLU decisions the engine's forward took: a pre-activation within one fp32 rounding of
            # zero (the batch mean summed in another order) may land on either side — invisible in the forward row,
            # but one such voxel in a 256-voxel channel of the bottom level moves that channel's dbeta by its whole dout
            keep = a_ours > 0
            rows.append((f"{pre}.{bi}", "relu_decisions_differing_from_torch(info)",
                         float((keep != (zref.detach() > 0)).sum().item())))
            aref = zref * keep
            dout = f32(g["dout"] if ci == 3 else g["da1"])
            gy, gg, gbb = torch.autograd.grad(aref, [yv, gam, bet], dout)
            per_channel = yv.numel() // yv.shape[1]
            # BatchNorm over a handful of samples is degenerate (2 samples: x-hat = +-1, the input gradient cancels to
            # rounding noise, a zero variance puts every output ON the ReLU threshold): reported, not comparable
            tag = "" if per_channel >= 16 else f"(degenerate:{per_channel}_samples_per_channel)"
            add(f"{pre}.{bi}", "dx" + tag, dy, gy)
            add(f"{pre}.{bi}", "dgamma" + tag, pg[f"{pre}.{bi}.weight"], gg)
            add(f"{pre}.{bi}", "dbeta" + tag, pg[f"{pre}.{bi}.bias"], gbb)
            del yv, aref, dout, gy, dy
        del xin
        torch.cuda.empty_cache()

    # ---- max pooling: bit-exact forward, first-maximum routing backward (+ the skip gradient added in place)
    for k in (1, 2, 3, 4):
        g = gt[f"pool{k}"]
        xs = f32(skips[k - 1]).requires_grad_(True)
        pref = F.max_pool3d(xs, 2)
        rows.append((f"down{k}.maxpool", "fwd_mismatches", float((f32(pooled[k - 1]) != pref.detach()).sum().item())))
        (gxs,) = torch.autograd.grad(pref, [xs], f32(g["dy"]))
        add(f"down{k}.maxpool", "dx(+skip)", f32(g["dx"]), rb(g["dskip_before"] + gxs))
        del xs, pref, gxs

    # ---- transposed convolutions
    for j in (1, 2, 3, 4):
        g = gt[f"up{j}.up"]
        xi = f32(g["x"]).requires_grad_(True)
        w = rb(par[f"up{j}.up.weight"]).requires_grad_(True)
        b = par[f"up{j}.up.bias"].clone().requires_grad_(True)
        uref = F.conv_transpose3d(xi, w, b, stride=2)
        ours_up, dup = f32(fw[f"up{j}.up"]), f32(g["dout"])
        pd, ph, pw = g["pad"]
        sl = (slice(None), slice(None), slice(pd, pd + uref.shape[2]), slice(ph, ph + uref.shape[3]),
              slice(pw, pw + uref.shape[4]))
        add(f"up{j}.up", "fwd", ours_up[sl], uref.detach())
        gx, gw, gb = torch.autograd.grad(uref, [xi, w, b], dup[sl].contiguous())
        add(f"up{j}.up", "dx", f32(g["dx"]), gx)
        add(f"up{j}.up", "dw", pg[f"up{j}.up.weight"], gw)
        add(f"up{j}.up", "db", pg[f"up{j}.up.bias"], gb)
        del xi, uref, gx, ours_up, dup
    return rows


def write_layerwise(rows, path, header=""):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write("# every op of one training step against the torch fp32 op / autograd on the engine's own inputs "
                "(identical inputs, rel-L2; *_mismatches = element count)\n")
        if header:
            f.write("# " + header + "\n")
        for op, what, v in rows:
            f.write(f"{op} {what} {v:.3e}\n")
