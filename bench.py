#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native 3D U-Net hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload train|infer|cv] [--global-batch G]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

workloads (BASELINE.json configs):
  train (default, the headline)  configs[1]/[2]: one training step = forward, BCE+Dice, backward, fused Adam on synthetic
                                 5x128^3 volumes, batch 2 per GPU (weak scaling; N=8 is the global batch 16 of
                                 configs[2]); --global-batch 16 splits that batch over the ranks instead (8/4/2 per GPU)
  infer                          configs[3]: sliding-window inference, one 5x256x256x64 volume per GPU, windows
                                 128x128x64 stride 64, each rank keeps its own volume (no exchange)
  cv                             configs[4]: zero_fill training step at 5x160^3, base channels 32 and 64 back to back,
                                 data-parallel over the ranks; `value` is the base-64 rate

Metric: spatial voxels (N*D*H*W, not x5 channels) per second, whole job.  ONE JSON line on rank 0; DESIGN.md
"Measurement" says how each field is obtained.  `--impl reference` times the reference's own modules (unmodified, from
baseline/_ref) on the host cores.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "prostate-cancer-multimodal-segmentation_b200"

VOLUME = (128, 128, 128)
PER_GPU_BATCH = 2
CPU_SAMPLE_SHAPE = (1, 5, 64, 64, 64)      # bounded sample of the training workload for the CPU legs (configs[0])
CPU_WINDOW_SHAPE = (1, 5, 128, 128, 64)    # one sliding window: 1/9 of a configs[3] volume
PROFILE_PASSES = 5


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default 2 = BASELINE configs[1])")
    ap.add_argument("--global-batch", type=int, default=None,
                    help="fixed global batch split over the ranks (configs[2]: 16 -> 8/4/2 per GPU at N=2/4/8)")
    ap.add_argument("--size", type=int, nargs=3, default=None)
    ap.add_argument("--base", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile-pass", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch eagerly (no CUDA-graph replay)")
    ap.add_argument("--dump-kernels", default=None, help="write the per-launch GEMM timing table of the profile passes")
    ap.add_argument("--workload", default="train", choices=["train", "infer", "cv"])
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ reference modules
def reference_modules():
    """(UNet3D, BCEDiceLoss, DiceLoss) of the UNMODIFIED reference, from baseline/_ref (staged by __graft_entry__.build()
    in the build container; git-ignored, travels with the snapshot), or None when it is not there"""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.exists(os.path.join(ref, "models", "unet3d.py")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        um = importlib.import_module("models.unet3d")
        lm = importlib.import_module("utils.losses")
        return um.UNet3D, lm.BCEDiceLoss, lm.DiceLoss
    except Exception:
        return None


def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    return importlib.import_module("unet3d_oracle")


# ------------------------------------------------------------------------------------------------ CPU legs
def cpu_train_steps(steps, warmup, threads=None):
    """The reference training step — utils/trainer.py:177-195: zero_grad, forward, BCEDiceLoss, backward,
    Adam(lr=1e-4, weight_decay=1e-5).step() — on the host cores, on CPU_SAMPLE_SHAPE.  With baseline/_ref it runs the
    reference's own UNet3D / BCEDiceLoss modules (kind "reference"), else the oracle port of them (kind "port").
    Returns (voxels/s, ms/step, cores, kind)."""
    import torch
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(*CPU_SAMPLE_SHAPE, generator=g)
    y = (torch.rand(CPU_SAMPLE_SHAPE[0], 1, *CPU_SAMPLE_SHAPE[2:], generator=g) < 0.1).float()
    ref = reference_modules()
    torch.manual_seed(0)
    if ref is not None:
        UNet3D, BCEDiceLoss, _ = ref
        model = UNet3D(5, 1).train()
        crit = BCEDiceLoss()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-5)

        def step():
            opt.zero_grad()
            loss = crit(model(x), y)
            loss.backward()
            opt.step()
        kind = "reference"
    else:
        oracle = _oracle()
        pkg = importlib.import_module(PKG)
        sd = {k: v.detach().clone() for k, v in pkg.UNet3D(5, 1).state_dict().items()}
        state = {}

        def step():
            oracle.train_step(sd, state, x, y)
        kind = "port"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    total = time.perf_counter() - t0
    vox = CPU_SAMPLE_SHAPE[0] * CPU_SAMPLE_SHAPE[2] * CPU_SAMPLE_SHAPE[3] * CPU_SAMPLE_SHAPE[4]
    return vox * steps / total, 1e3 * total / steps, cores, kind


def cpu_infer_windows(steps, warmup, threads=None):
    """UNet3D.predict (models/unet3d.py:298-318) of one 128x128x64 window on the host cores; a configs[3] volume is 9
    such windows for 256*256*64 output voxels.  Returns (output voxels/s, ms/window, cores, kind)."""
    import torch
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(99)
    x = torch.rand(*CPU_WINDOW_SHAPE, generator=g)
    ref = reference_modules()
    torch.manual_seed(0)
    if ref is not None:
        model = ref[0](5, 1)
        run, kind = (lambda: model.predict(x)), "reference"
    else:
        oracle = _oracle()
        pkg = importlib.import_module(PKG)
        sd = {k: v.detach().clone() for k, v in pkg.UNet3D(5, 1).state_dict().items()}
        run, kind = (lambda: oracle.predict(x, sd)), "port"
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    per = (time.perf_counter() - t0) / steps
    return 256 * 256 * 64 / (9 * per), 1e3 * per, cores, kind


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if args.workload == "infer":
        v, ms, cores, kind = cpu_infer_windows(max(1, args.steps), max(1, min(args.warmup, 2)))
        metric, wl = "infer_voxels_per_s", "UNet3D sliding-window inference (BASELINE configs[3]), CPU bounded sample"
        sample = (f"UNet3D.predict of one 1x5x128x128x64 window per step (1/9 of a 5x256x256x64 volume; voxels/s = "
                  f"256*256*64 / (9 x window time)); {ms:.0f} ms/window")
        shape = CPU_WINDOW_SHAPE
    else:
        v, ms, cores, kind = cpu_train_steps(args.steps, max(1, min(args.warmup, 2)))
        metric, wl = "train_voxels_per_s", "UNet3D(5->1, base 64) training step fwd+BCEDice+bwd+Adam, CPU bounded sample"
        sample = (f"training step on {CPU_SAMPLE_SHAPE[0]}x5x{CPU_SAMPLE_SHAPE[2]}^3 fp32 per step (BASELINE configs[0]: "
                  f"1/16 of one rank's 2x5x128^3 batch); {ms:.0f} ms/step")
        shape = CPU_SAMPLE_SHAPE
    what = ("the reference's own modules (unmodified models/unet3d.py, utils/losses.py from baseline/_ref), stock torch "
            "CPU ops" if kind == "reference" else "oracle port (oracle/unet3d_oracle.py) of the reference modules")
    line = {
        "impl": "reference", "metric": metric, "value": v, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "sample_shape": list(shape), "code": what},
        "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t_wait = time.time()
            while not self.rows and time.time() - t_wait < 5.0:  # first nvidia-smi sample can take a second
                time.sleep(0.05)
            self.rows.clear()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        thr = 0.5 * max(pw)   # "under load" = samples drawing more than half of the peak observed power
        load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw)}


# ------------------------------------------------------------------------------------------------ shared helpers
def peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pk = {}
    if "bf16_tflops_sustained" in pk:
        return pk["bf16_tflops_sustained"], "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    return 1400.0, "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"


def profile_passes(ops, run_once, passes, ms_step, dump=None):
    """`passes` eager executions of the step with CUDA events around every GEMM launch (ops.profile_hook).  Returns the
    `roofline` object: dominant GEMM kernel by time, achieved = algorithmic FLOPs / summed duration, MEDIAN over the
    passes (a single pass swings by +-15 % with the clocks), plus the per-kernel and per-layer tables."""
    import torch
    per_pass, layer_ms = [], {}
    last = []
    for _ in range(passes):
        recs = []
        ops.profile_hook = lambda k, tag, fl, a, b, shape=None: recs.append((k, tag, fl, a, b, shape))
        try:
            run_once()
            torch.cuda.synchronize()
        finally:
            ops.profile_hook = None
        agg = {}
        for k, tag, fl, a, b, shape in recs:
            ms = a.elapsed_time(b)
            d = agg.setdefault(k, {"ms": 0.0, "flops": 0.0, "launches": 0})
            d["ms"] += ms; d["flops"] += fl; d["launches"] += 1
            if shape is not None:
                layer_ms.setdefault((tag, tuple(shape)), []).append(ms)
        per_pass.append(agg)
        last = recs
    if dump:
        with open(dump, "w") as f:
            f.write("kernel tag gflop ms tflops voxels cin cout   (last of the profile passes)\n")
            for k, tag, fl, a, b, shape in last:
                ms = a.elapsed_time(b)
                f.write(f"{k} {tag} {fl / 1e9:.1f} {ms:.4f} {fl / (ms * 1e-3) / 1e12:.1f} "
                        f"{' '.join(str(v) for v in (shape or ()))}\n")
    names = sorted(per_pass[0])
    med = {}
    for k in names:
        ms = statistics.median(p[k]["ms"] for p in per_pass if k in p)
        fl = per_pass[0][k]["flops"]
        med[k] = {"ms": ms, "flops": fl, "launches": per_pass[0][k]["launches"],
                  "tflops_passes": [round(p[k]["flops"] / (p[k]["ms"] * 1e-3) / 1e12, 1) for p in per_pass if k in p]}
    top = max(med, key=lambda k: med[k]["ms"])
    peak, peak_src = peaks()
    a = med[top]
    achieved = statistics.median(a["tflops_passes"])
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(top)
    except Exception:
        pass
    roof = {"bound": "tensor", "kernel": top, "achieved": round(achieved, 2), "peak": peak, "unit": "TFLOP/s",
            "frac": round(achieved / peak, 4),
            "traffic": traffic.get("dram_bytes") if isinstance(traffic, dict) else None, "traffic_detail": traffic,
            "peak_source": peak_src, "passes": len(per_pass), "achieved_per_pass": a["tflops_passes"],
            "avg_launch_ms": round(a["ms"] / a["launches"], 4),
            "algorithmic_flops_per_launch": a["flops"] / a["launches"],
            "gemm_share_of_step": round(sum(v["ms"] for v in med.values()) / ms_step, 3),
            "kernels": {k: {"ms_per_step": round(v["ms"], 3), "launches": v["launches"],
                            "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1)} for k, v in med.items()},
            "frac_of_nominal_2250": round(achieved / 2250.0, 4)}
    layers = {key: statistics.median(v) for key, v in layer_ms.items()}
    return roof, layers


def library_baseline(dev, batch, size, base, layers, steps=4, zero_fill=False):
    """The existing Blackwell library path on the same GPU (SURVEY.md 8d "GPU library baseline"): the training step of
    the reference's own modules (baseline/_ref; the oracle's graph when absent) through stock torch — cuDNN 3-D
    convolutions, ATen BatchNorm / pooling / loss, torch.optim.Adam — in fp32 and in its best configuration (bf16
    autocast + channels_last_3d); and per 3x3x3 layer cuDNN forward / backward (dgrad + wgrad) next to this path's
    kernels for the same layer (`layers` = per-launch medians of the profile passes)."""
    import torch
    import torch.nn.functional as F
    D, H, W = size
    out = {"config": f"batch {batch} x 5 x {D}x{H}x{W}, base {base}"}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(batch, 5, D, H, W, generator=g).to(dev)
    y = (torch.rand(batch, 1, D, H, W, generator=g) < 0.1).float().to(dev)
    ref = reference_modules() if base == 64 else None   # the reference hard-codes 64 base channels
    out["code"] = "reference modules (baseline/_ref)" if ref is not None else "oracle graph (stock torch functional ops)"
    for label, autocast, cl3d in (("bf16_autocast_channels_last_3d", True, True), ("fp32", False, False)):
        try:
            torch.manual_seed(0)
            xx = x.contiguous(memory_format=torch.channels_last_3d) if cl3d else x
            if ref is not None:
                model = ref[0](5, 1).to(dev).train()
                if cl3d:
                    model = model.to(memory_format=torch.channels_last_3d)
                crit = ref[1]()
                opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-5)

                def step():
                    opt.zero_grad()
                    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                        loss = crit(model(xx), y)
                    loss.backward()
                    opt.step()
            else:
                oracle = _oracle()
                pkg = importlib.import_module(PKG)
                work = {k: v.detach().clone().to(dev) for k, v in pkg.UNet3D(5, 1, init_features=base).state_dict().items()}
                if cl3d:
                    work = {k: (v.contiguous(memory_format=torch.channels_last_3d) if v.dim() == 5 else v)
                            for k, v in work.items()}
                state = {}

                def step():
                    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                        oracle.train_step(work, state, xx, y)
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[label] = {"ms_per_step": round(ms, 3), "voxels_per_s": batch * D * H * W / (ms * 1e-3)}
        except Exception as e:   # a side measurement: report (e.g. out of memory), do not fail the bench
            out[label] = {"error": repr(e)[:200]}
        model = opt = work = state = None  # noqa: F841
        torch.cuda.empty_cache()
    # ---- per 3x3x3 layer: cuDNN (bf16, channels_last_3d) vs this path's kernels
    per_layer = {}
    shapes = sorted({s for (tag, s) in layers if tag == "conv3d_fprop"} | {s for (tag, s) in layers if tag == "conv1_fprop"})
    vox_total = batch * D * H * W
    for (vox, cin, cout) in shapes:
        lvl = round((vox_total / vox) ** (1 / 3))
        d, h, w = D // lvl, H // lvl, W // lvl
        if batch * d * h * w != vox:
            continue
        try:
            xi = torch.randn(batch, cin, d, h, w, device=dev, dtype=torch.bfloat16).contiguous(
                memory_format=torch.channels_last_3d).requires_grad_(cin != 5)
            wt = (torch.randn(cout, cin, 3, 3, 3, device=dev, dtype=torch.bfloat16) * 0.05).contiguous(
                memory_format=torch.channels_last_3d).requires_grad_(True)
            gy = None
            tf, tb = [], []
            for it in range(4):
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                yy = F.conv3d(xi, wt, None, padding=1)
                e[1].record()
                if gy is None:
                    gy = torch.randn_like(yy)
                yy.backward(gy)
                e[2].record()
                torch.cuda.synchronize()
                xi.grad = None; wt.grad = None
                if it:
                    tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
            first = cin == 5
            ours_f = layers.get(("conv1_fprop" if first else "conv3d_fprop", (vox, cin, cout)))
            ours_b = (layers.get(("conv3d_dgrad", (vox, cin, cout)), 0.0) if not first else 0.0) + \
                layers.get(("conv1_wgrad" if first else "conv3d_wgrad", (vox, cin, cout)), 0.0)
            per_layer[f"{cin}->{cout}@{d}x{h}x{w}"] = {
                "cudnn_fwd_ms": round(statistics.median(tf), 4), "cudnn_bwd_ms": round(statistics.median(tb), 4),
                "b200_fwd_ms": round(ours_f, 4) if ours_f else None, "b200_bwd_ms": round(ours_b, 4) if ours_b else None}
        except Exception as e:
            per_layer[f"{cin}->{cout}@{d}x{h}x{w}"] = {"error": repr(e)[:120]}
        xi = wt = gy = yy = None  # noqa: F841
        torch.cuda.empty_cache()
    out["per_layer_3x3x3"] = per_layer
    ok = [v for v in per_layer.values() if v.get("b200_fwd_ms") and v.get("b200_bwd_ms")]
    if ok:
        out["per_layer_sum_ms"] = {"cudnn": round(sum(v["cudnn_fwd_ms"] + v["cudnn_bwd_ms"] for v in ok), 3),
                                   "b200": round(sum(v["b200_fwd_ms"] + v["b200_bwd_ms"] for v in ok), 3),
                                   "note": "one instance of every distinct layer shape"}
    return out


def init_dist(dev):
    import torch
    import torch.distributed as dist
    # stdout carries the one JSON line: NCCL prints its version banner (NCCL_DEBUG=VERSION/WARN) with printf while the
    # communicator is created, so file descriptor 1 points at stderr until the first collective has run
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        warm = torch.zeros(1, device=dev)
        dist.all_reduce(warm)
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)


class Ctx:
    """timing helpers shared by the workloads: barrier + synchronize on both sides, CUDA events, max over ranks"""

    def __init__(self, dev, world, rank):
        import torch
        self.torch, self.dev, self.world, self.rank = torch, dev, world, rank

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world == 1:
            return ms
        import torch.distributed as dist
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def timed(self, fn, steps):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)), out


# ------------------------------------------------------------------------------------------------ inference workload
def run_infer(args, pkg, par, ctx):
    """BASELINE configs[3]: sliding-window inference, `world` volumes of 5x256x256x64, one per rank (the window
    schedule is dealt in contiguous blocks, so a volume's 9 windows stay on one rank: no exchange), sigmoid +
    threshold.  voxels/s = output voxels of all volumes / time."""
    import torch
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    ops = pkg.ops
    torch.manual_seed(0)
    model = pkg.UNet3D(5, 1, init_features=args.base).to(dev).eval()
    g = torch.Generator().manual_seed(99)
    x_host = torch.rand(world, 5, 256, 256, 64, generator=g).pin_memory()
    x = x_host.to(dev)
    window, stride = (128, 128, 64), (64, 64, 64)
    mine = par.owned_volumes(tuple(x.shape), window, stride, rank, world)
    assert mine == par.rank_volumes(tuple(x.shape), window, stride, rank, world) == [rank], "one volume per rank"

    def step():
        return par.sliding_window_predict(model, x, window, stride, rank=rank, world=world)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count
    ms, _ = ctx.timed(step, args.steps)
    launches = ops.launch_count - l0
    # end to end: every rank uploads its volume from pinned host memory and downloads its mask, every step
    m_host = torch.empty(world, 1, 256, 256, 64, dtype=torch.uint8).pin_memory()

    def e2e_step():
        for v in mine:
            x[v].copy_(x_host[v], non_blocking=True)
        _, mask = step()
        m8 = mask.to(torch.uint8)
        for v in mine:
            m_host[v].copy_(m8[v], non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the host holds this step's masks before the next one starts

    e2e_step()
    ms2, _ = ctx.timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None   # sampled over both timed regions (100 ms period)
    vox = world * 256 * 256 * 64
    ms_step = ms / args.steps
    roofline = None
    if not args.no_profile_pass and rank == 0:
        roofline, _ = profile_passes(ops, step, PROFILE_PASSES, ms_step, args.dump_kernels)
    ctx.barrier()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms_cpu, cores, kind = cpu_infer_windows(2, 1)
        cpu = {"value": v, "unit": "voxels/s", "cores": cores, "kind": kind,
               "sample": f"UNet3D.predict of 2 windows 1x5x128x128x64 (a volume is 9); {ms_cpu:.0f} ms/window"}
    if rank == 0:
        fwd, _ = importlib.import_module(PKG + ".engine").total_flops_per_voxel(args.base, 5, 1)
        win_vox = 9 * 128 * 128 * 64 * world
        print(json.dumps({
            "metric": "infer_voxels_per_s", "value": vox / (ms_step * 1e-3), "unit": "voxels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "UNet3D sliding-window inference, 5x256x256x64 volumes (BASELINE configs[3]), "
                                   "window 128x128x64 stride 64 (9 windows/volume, evaluated as one batch), one "
                                   "volume per GPU", "volumes": world,
                       "parallelism": f"windows sharded over {world} rank(s), no exchange",
                       "l2": "a window batch touches GBs of activations (> 126 MB L2); no explicit flush"},
            "clocks": clocks,
            "e2e": {"value": vox / (ms2 / args.steps * 1e-3), "unit": "voxels/s",
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": int(m_host.numel()),
                    "note": "per step every rank uploads its volume from pinned memory and downloads its mask"},
            "gpu_launches": launches,
            "model_tflops": round(win_vox * fwd / (ms_step * 1e-3) / 1e12, 1),
            "roofline": roofline, "cpu_baseline": cpu}), flush=True)


# ------------------------------------------------------------------------------------------------ training workloads
def make_batch(B, size, seed, zero_fill):
    import torch
    D, H, W = size
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 5, D, H, W, generator=g)
    y = (torch.rand(B, 1, D, H, W, generator=g) < 0.1).float()
    if zero_fill:   # each of modalities 1..4 absent with p = 0.2: a whole channel of zeros (script/data_loader.py:320)
        present = torch.rand(B, 5, generator=g) >= 0.2
        present[:, 0] = True
        x = x * present[:, :, None, None, None].float()
    return x.pin_memory(), y.pin_memory()


def train_arm(args, pkg, par, ctx, B, size, base, crit_name, zero_fill, want_profile, want_library):
    """times the training step of one configuration; returns the measurements as a dict"""
    import torch
    import gc
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    ops = pkg.ops
    D, H, W = size
    torch.manual_seed(0)
    model = pkg.UNet3D(5, 1, init_features=base).to(dev)
    crit = pkg.BCEDiceLoss() if crit_name == "bce_dice" else pkg.DiceLoss()
    opt = pkg.FusedAdam(model, lr=1e-4, weight_decay=1e-5)
    sync = par.make_data_parallel(model, opt) if world > 1 else None
    model.train()
    x_host, y_host = make_batch(B, size, 1234 + rank, zero_fill)
    x, y = x_host.to(dev), y_host.to(dev)
    vox_step = B * D * H * W * world

    def eager_step(xx=x, yy=y):
        opt.zero_grad()
        loss = crit(model(xx), yy)
        loss.backward()
        opt.step()
        return loss

    # what BaseTrainer._step runs: the step replayed from a CUDA graph after two eager steps (graph.GraphedTrainStep;
    # same kernels, same arithmetic); data-parallel ranks record their NCCL bucket all-reduces into the graph as well
    # (B200_GRAPH_DP=0: eager launches on data-parallel ranks)
    graph_dp = os.environ.get("B200_GRAPH_DP", "1") != "0"
    use_graph = not args.no_graph and (world == 1 or graph_dp)
    graphed = pkg.GraphedTrainStep(model, crit, opt, capture_collectives=world > 1) if use_graph else None
    step = graphed if graphed is not None else eager_step
    for _ in range(max(args.warmup, 3)):
        step(x, y)
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count
    ms_total, loss = ctx.timed(lambda: step(x, y), args.steps)
    launches = ops.launch_count - l0
    ms_step = ms_total / args.steps
    res = {"value": vox_step / (ms_step * 1e-3), "ms_step": ms_step, "loss": loss.item(), "launches": launches,
           "vox_step": vox_step,
           "launch": ("CUDA-graph replay of the step" + (" incl. the NCCL bucket all-reduces" if world > 1 else "")
                      if graphed is not None and graphed.replays > 0
                      else "eager launches" + (f" ({graphed.disabled})" if graphed is not None and graphed.disabled else ""))}
    # ---- end to end through the public API with host buffers: the trainer's own input path (trainer.py:train_epoch):
    # every step's image+label go pinned host -> device through data.DevicePrefetcher (copy of step i+1 on a copy stream
    # under the kernels of step i), every step's loss comes back through data.AsyncScalarReader (read one step late)
    prefetcher = pkg.data.DevicePrefetcher([], dev)
    for batch in prefetcher.over([{"image": x_host, "label": y_host}] * 3):
        step(batch["image"], batch["label"])
    ctx.barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e2.record()
    losses = pkg.data.AsyncScalarReader()
    for batch in prefetcher.over([{"image": x_host, "label": y_host}] * args.steps):
        losses.push(step(batch["image"], batch["label"]))
    assert len(losses.finish()) == args.steps
    e3.record()
    ctx.barrier()
    ms_e2e = ctx.max_over_ranks(max(e2.elapsed_time(e3), (time.perf_counter() - t_host0) * 1e3 if world == 1 else 0.0))
    res["clocks"] = sampler.stop() if rank == 0 else None   # sampled over both timed regions (100 ms period)
    res["e2e"] = {"value": vox_step / (ms_e2e / args.steps * 1e-3), "unit": "voxels/s",
                  "h2d_bytes_per_step": (x_host.numel() + y_host.numel()) * 4 * world, "d2h_bytes_per_step": 4 * world,
                  "ms_per_step": ms_e2e / args.steps}
    if sync is not None:
        res["buckets"] = len(sync._buckets or [])
    # ---- per-kernel passes: CUDA events around every GEMM launch of PROFILE_PASSES more (eager) steps
    layers = {}
    if want_profile:
        model.engine.overlap_wgrad = False   # per-kernel durations: no concurrent side-stream kernels in these passes
        sync_saved, model.engine.grad_sync = model.engine.grad_sync, None   # rank-local passes (no collective inside)
        try:
            if rank == 0 or world == 1:
                res["roofline"], layers = profile_passes(ops, eager_step, PROFILE_PASSES, ms_step, args.dump_kernels)
        finally:
            model.engine.overlap_wgrad = True
            model.engine.grad_sync = sync_saved
        if world > 1:   # the passes above changed rank 0's replica only: bring the replicas back together
            import torch.distributed as dist
            ctx.barrier()
            dist.broadcast(model.engine.flat_param, src=0)
    # recorded NCCL kernels keep the communicator busy until the graph object is gone
    graphed = step = None  # noqa: F841
    del model, opt, crit, prefetcher
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if want_library and rank == 0 and world == 1:
        res["library"] = library_baseline(dev, B, size, base, layers, zero_fill=zero_fill)
    return res


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    par = importlib.import_module(PKG + ".parallel")
    eng_mod = importlib.import_module(PKG + ".engine")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_dist(dev)
    pkg.load_library()
    ctx = Ctx(dev, world, rank)
    try:
        if args.workload == "infer":
            run_infer(args, pkg, par, ctx)
            return
        peak, _ = peaks()
        if args.workload == "cv":
            size = tuple(args.size) if args.size else (160, 160, 160)
            B = args.batch or 1
            arms = {}
            for base in (32, 64):
                arms[base] = train_arm(args, pkg, par, ctx, B, size, base, "dice", True,
                                       want_profile=not args.no_profile_pass, want_library=False)
            r = arms[64]
            workload = (f"UNet3D zero_fill training step (fwd + DiceLoss + bwd + Adam) at 5x{size[0]}x{size[1]}x{size[2]}, "
                        f"batch {B}/GPU, base channels 32 and 64 back to back (BASELINE configs[4]); value = base 64")
            extra = {"cv": {f"base{b}": {"voxels_per_s": a["value"], "ms_per_step": a["ms_step"],
                                         "e2e_voxels_per_s": a["e2e"]["value"],
                                         "model_tflops": round(a["value"] * eng_mod.total_flops_per_voxel(b, 5, 1)[1] / 1e12, 1),
                                         "kernels": (a.get("roofline") or {}).get("kernels")}
                            for b, a in arms.items()}}
            base = 64
        else:
            size = tuple(args.size) if args.size else VOLUME
            if args.global_batch:
                if args.global_batch % world:
                    raise SystemExit(f"--global-batch {args.global_batch} is not a multiple of {world} ranks")
                B = args.global_batch // world
            else:
                B = args.batch or PER_GPU_BATCH
            r = train_arm(args, pkg, par, ctx, B, size, args.base, "bce_dice", False,
                          want_profile=not args.no_profile_pass, want_library=not args.no_library_baseline)
            workload = (f"UNet3D(5->1, base {args.base}) training step fwd+BCEDice+bwd+Adam, batch {B}/GPU, "
                        f"5x{size[0]}x{size[1]}x{size[2]} (BASELINE configs[1]; N=8 -> configs[2])")
            extra = {}
            base = args.base
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            v, ms_cpu, cores, kind = cpu_train_steps(3, 1)
            cpu = {"value": v, "unit": "voxels/s", "cores": cores, "kind": kind,
                   "sample": f"3 steps of the reference training step on {CPU_SAMPLE_SHAPE[0]}x5x{CPU_SAMPLE_SHAPE[2]}^3 "
                             f"fp32 (BASELINE configs[0]); {ms_cpu:.0f} ms/step"}
        if rank == 0:
            _, fb = eng_mod.total_flops_per_voxel(base, 5, 1)
            line = {
                "metric": "train_voxels_per_s", "value": r["value"], "unit": "voxels/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_step"],
                "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload, "global_batch": r["vox_step"] // (size[0] * size[1] * size[2]),
                           "per_gpu_batch": B, "volume": list(size), "parallelism": f"dp{world}", "launch": r["launch"],
                           "l2": "per-step working set (GBs of activations) exceeds the 126 MB L2; no explicit flush",
                           "loss": r["loss"]},
                "clocks": r["clocks"], "e2e": r["e2e"], "gpu_launches": r["launches"],
                "model_tflops": round(r["value"] * fb / 1e12, 2),
                "roofline": r.get("roofline"), "cpu_baseline": cpu,
            }
            if "library" in r:
                lib = r["library"]
                best = lib.get("bf16_autocast_channels_last_3d", {})
                if "ms_per_step" in best:
                    lib["speedup_vs_best_library_config"] = round(best["ms_per_step"] / r["ms_step"], 3)
                line["gpu_library_baseline"] = lib
            if "buckets" in r:
                line["config"]["allreduce_buckets_per_step"] = r["buckets"]
            line.update(extra)
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            import gc
            gc.collect()
            torch.cuda.synchronize()
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
