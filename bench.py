#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native 3D U-Net hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one training step (H2D excluded for `value`, included for `e2e`): forward, BCE+Dice loss, backward, fused
Adam, on batch 2 per GPU of synthetic 5x128^3 volumes (BASELINE.json configs[1]; at N=8 the global batch is 16 =
configs[2]).  Metric: spatial voxels (N*D*H*W, not x5 channels) per second, whole job.

Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for how each field is obtained.
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
CPU_SAMPLE_SHAPE = (1, 5, 64, 64, 64)  # bounded sample of the workload for the CPU legs (BASELINE configs[0])


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch (default: BASELINE configs[1])")
    ap.add_argument("--size", type=int, nargs=3, default=list(VOLUME))
    ap.add_argument("--base", type=int, default=64)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile-pass", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch eagerly (no CUDA-graph replay)")
    ap.add_argument("--dump-kernels", default=None, help="write the per-launch GEMM timing table of the profile pass")
    ap.add_argument("--workload", default="train", choices=["train", "infer"],
                    help="train: BASELINE configs[1]/[2] (default, the headline); infer: configs[3] sliding-window "
                         "inference on 5x256x256x64 volumes, one volume per GPU, windows 128x128x64 stride 64")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def cpu_oracle_steps(steps, warmup, threads=None):
    """times the oracle's restatement of the reference training step (utils/trainer.py:177-195 with BCEDiceLoss and
    Adam(lr=1e-4, weight_decay=1e-5)) on the host cores, on CPU_SAMPLE_SHAPE.  Returns (voxels/s, ms/step, cores)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import unet3d_oracle as oracle
    pkg = importlib.import_module(PKG)
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = pkg.UNet3D(5, 1)  # parameter container only: seed-identical to the reference's init
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    del model
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(*CPU_SAMPLE_SHAPE, generator=g)
    y = (torch.rand(CPU_SAMPLE_SHAPE[0], 1, *CPU_SAMPLE_SHAPE[2:], generator=g) < 0.1).float()
    state = {}
    for _ in range(warmup):
        oracle.train_step(sd, state, x, y)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        oracle.train_step(sd, state, x, y)
        times.append(time.perf_counter() - t0)
    vox = CPU_SAMPLE_SHAPE[0] * CPU_SAMPLE_SHAPE[2] * CPU_SAMPLE_SHAPE[3] * CPU_SAMPLE_SHAPE[4]
    total = sum(times)
    return vox * steps / total, 1e3 * total / steps, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, ms, cores = cpu_oracle_steps(args.steps, max(1, min(args.warmup, 2)))
    sample = (f"oracle port (oracle/unet3d_oracle.py, torch fp32 CPU ops) of the reference training step on "
              f"{CPU_SAMPLE_SHAPE[0]}x5x{CPU_SAMPLE_SHAPE[2]}^3 (1/16 of one rank's 2x5x128^3 batch) per step")
    line = {
        "impl": "reference", "metric": "train_voxels_per_s", "value": v, "unit": "voxels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "UNet3D(5->1, base 64) training step fwd+BCEDice+bwd+Adam, CPU bounded sample",
                   "sample_shape": list(CPU_SAMPLE_SHAPE)},
        "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample},
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
        # "under load" = samples drawing more than half of the peak observed power
        thr = 0.5 * max(pw)
        load = [s for s, p in zip(sm, pw) if p >= thr] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw)}


# ------------------------------------------------------------------------------------------------ inference workload
def run_infer(args, pkg, par, dev, world, rank):
    """BASELINE configs[3]: sliding-window inference, `world` volumes of 5x256x256x64, the 9*world windows dealt
    round-robin to the ranks, one all-reduce of the logit volumes, sigmoid + threshold.  voxels/s = output voxels."""
    import torch
    import torch.distributed as dist
    torch.manual_seed(0)
    model = pkg.UNet3D(5, 1, init_features=args.base).to(dev).eval()
    g = torch.Generator().manual_seed(99)
    x_host = torch.rand(world, 5, 256, 256, 64, generator=g).pin_memory()
    x = x_host.to(dev)
    window, stride = (128, 128, 64), (64, 64, 64)

    def step():
        return par.sliding_window_predict(model, x, window, stride, rank=rank, world=world)

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = pkg.ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        probs, mask = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    launches = pkg.ops.launch_count - l0
    # end to end: host volume in, host mask out
    # every rank uploads only the volumes its windows read and downloads the masks of those volumes
    vols = par.rank_volumes(tuple(x.shape), window, stride, rank, world)
    m_host = torch.empty(world, 1, 256, 256, 64, dtype=torch.uint8).pin_memory()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2.record()
    for _ in range(args.steps):
        for v in vols:
            x[v].copy_(x_host[v], non_blocking=True)
        probs, mask = step()
        m8 = mask.to(torch.uint8)
        for v in vols:
            m_host[v].copy_(m8[v], non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the host holds this step's masks before the next one starts
    e3.record()
    torch.cuda.synchronize()
    ms2 = e2.elapsed_time(e3)
    if world > 1:
        t = torch.tensor([ms2], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2 = t.item()
    vox = world * 256 * 256 * 64
    if rank == 0:
        fwd, _ = importlib.import_module(PKG + ".engine").total_flops_per_voxel(args.base, 5, 1)
        win_vox = 9 * 128 * 128 * 64 * world
        print(json.dumps({
            "metric": "infer_voxels_per_s", "value": vox / (ms / args.steps * 1e-3), "unit": "voxels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "UNet3D sliding-window inference, 5x256x256x64 volumes (BASELINE configs[3]), "
                                   "window 128x128x64 stride 64 (9 windows/volume), one volume per GPU",
                       "volumes": world, "parallelism": f"windows sharded over {world} rank(s)"},
            "e2e": {"value": vox / (ms2 / args.steps * 1e-3), "unit": "voxels/s",
                    "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": int(m_host.numel()),
                    "note": "per step every rank uploads the volumes its windows read and downloads their masks"},
            "gpu_launches": launches,
            "model_tflops": round(win_vox * fwd / (ms / args.steps * 1e-3) / 1e12, 1)}), flush=True)


# ------------------------------------------------------------------------------------------------ main arm
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    ops = pkg.ops
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
        # stdout carries the one JSON line: NCCL prints its version banner (NCCL_DEBUG=VERSION/WARN) with printf while
        # the communicator is created, so file descriptor 1 points at stderr until the first collective has run
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
    pkg.load_library()

    if args.workload == "infer":
        run_infer(args, pkg, par, dev, world, rank)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    D, H, W = args.size
    B = args.batch
    torch.manual_seed(0)
    model = pkg.UNet3D(5, 1, init_features=args.base).to(dev)
    crit = pkg.BCEDiceLoss()
    opt = pkg.FusedAdam(model, lr=1e-4, weight_decay=1e-5)
    sync = par.make_data_parallel(model, opt) if world > 1 else None
    model.train()

    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, 5, D, H, W, generator=g).pin_memory()
    y_host = (torch.rand(B, 1, D, H, W, generator=g) < 0.1).float().pin_memory()
    x = x_host.to(dev)
    y = y_host.to(dev)
    vox_step = B * D * H * W * world

    def eager_step(xx, yy):
        opt.zero_grad()
        out = model(xx)
        loss = crit(out, yy)
        loss.backward()
        opt.step()
        return loss

    # what BaseTrainer._step runs: on one process the step is replayed from a CUDA graph after two eager steps
    # (graph.GraphedTrainStep; same kernels, same arithmetic); data-parallel ranks issue it eagerly
    graph_dp = os.environ.get("B200_GRAPH_DP") == "1"   # experiment: record the NCCL all-reduces too
    graphed = (pkg.GraphedTrainStep(model, crit, opt, capture_collectives=graph_dp)
               if ((world == 1 or graph_dp) and not args.no_graph) else None)
    step = graphed if graphed is not None else eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    for _ in range(args.warmup):
        step(x, y)
    barrier()

    # ---- device-resident throughput (`value`)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 5.0:  # first nvidia-smi sample can take a second
            time.sleep(0.05)
        sampler.rows.clear()
    barrier()
    l0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step(x, y)
    e1.record()
    barrier()
    launches = ops.launch_count - l0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = vox_step / (ms_step * 1e-3)
    final_loss = loss.item()

    # ---- end to end through the public API with host buffers (`e2e`)
    # warm the staging path (device slots of the prefetcher, copy stream) like the compute path: untimed steps
    prefetcher = pkg.data.DevicePrefetcher([], dev)
    for batch in prefetcher.over([{"image": x_host, "label": y_host}] * max(2, min(args.warmup, 3))):
        step(batch["image"], batch["label"])
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e2.record()
    # the trainer's own input path (trainer.py:train_epoch): every step's image+label go pinned host -> device through
    # data.DevicePrefetcher, which issues the copy of step i+1 on a copy stream under the kernels of step i
    host_batches = [{"image": x_host, "label": y_host}] * args.steps
    losses = pkg.data.AsyncScalarReader()  # every step's loss comes back to the host, read one step late
    for batch in prefetcher.over(host_batches):
        losses.push(step(batch["image"], batch["label"]))
    e2e_losses = losses.finish()
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(max(e2.elapsed_time(e3), (time.perf_counter() - t_host0) * 1e3 if world == 1 else 0.0))
    e2e_value = vox_step / (ms_e2e / args.steps * 1e-3)
    h2d = (x_host.numel() + y_host.numel()) * 4 * world
    d2h = 4 * world
    assert len(e2e_losses) == args.steps

    # ---- per-kernel pass: CUDA events around every GEMM launch of one more step (dominant-kernel roofline)
    roofline = None
    if not args.no_profile_pass:
        recs = []
        ops.profile_hook = lambda k, tag, fl, a, b: recs.append((k, tag, fl, a, b))
        model.engine.overlap_wgrad = False   # per-kernel durations: no concurrent side-stream kernels in this pass
        eager_step(x, y)
        torch.cuda.synchronize()
        model.engine.overlap_wgrad = True
        ops.profile_hook = None
        if args.dump_kernels and rank == 0:
            with open(args.dump_kernels, "w") as f:
                f.write("kernel tag gflop ms tflops\n")
                for k, tag, fl, a, b in recs:
                    ms = a.elapsed_time(b)
                    f.write(f"{k} {tag} {fl / 1e9:.1f} {ms:.4f} {fl / (ms * 1e-3) / 1e12:.1f}\n")
        agg = {}
        for k, tag, fl, a, b in recs:
            d = agg.setdefault(k, {"ms": 0.0, "flops": 0.0, "launches": 0})
            d["ms"] += a.elapsed_time(b); d["flops"] += fl; d["launches"] += 1
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" \
            if "bf16_tflops_sustained" in peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        top = max(agg, key=lambda k: agg[k]["ms"])
        kern = {k: {"ms_per_step": round(v["ms"], 3), "launches": v["launches"],
                    "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1)} for k, v in agg.items()}
        a = agg[top]
        achieved = a["flops"] / (a["ms"] * 1e-3) / 1e12
        # DRAM bytes of one captured launch of the dominant kernel (committed ncu --set full summary), if present
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(top)
        except Exception:
            pass
        roofline = {"bound": "tensor", "kernel": top, "achieved": round(achieved, 2), "peak": peak,
                    "unit": "TFLOP/s", "frac": round(achieved / peak, 4),
                    "traffic": traffic.get("dram_bytes") if isinstance(traffic, dict) else None,
                    "traffic_detail": traffic, "peak_source": peak_src,
                    "avg_launch_ms": round(a["ms"] / a["launches"], 4),
                    "algorithmic_flops_per_launch": a["flops"] / a["launches"],
                    "gemm_share_of_step": round(sum(v["ms"] for v in agg.values()) / ms_step, 3),
                    "kernels": kern, "frac_of_nominal_2250": round(achieved / 2250.0, 4)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms_cpu, cores = cpu_oracle_steps(3, 1)
        cpu = {"value": v, "unit": "voxels/s", "cores": cores, "kind": "port",
               "sample": f"3 steps of the oracle training step on {CPU_SAMPLE_SHAPE[0]}x5x{CPU_SAMPLE_SHAPE[2]}^3 fp32 "
                         f"(BASELINE configs[0]); {ms_cpu:.0f} ms/step"}

    if rank == 0:
        fwd, fb = eng_mod.total_flops_per_voxel(args.base, 5, 1)
        line = {
            "metric": "train_voxels_per_s", "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"UNet3D(5->1, base {args.base}) training step fwd+BCEDice+bwd+Adam, "
                                   f"batch {B}/GPU, 5x{D}x{H}x{W} (BASELINE configs[1]; N=8 -> configs[2])",
                       "global_batch": B * world, "per_gpu_batch": B, "volume": [D, H, W],
                       "parallelism": f"dp{world}",
                       "launch": ("CUDA-graph replay of the step" if graphed is not None and graphed.replays > 0
                                  else "eager launches" + (f" ({graphed.disabled})" if graphed is not None and
                                                           graphed.disabled else "")),
                       "l2": "per-step working set (GBs of activations) exceeds the "
                                                           "126 MB L2; no explicit flush",
                       "loss": final_loss},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "voxels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "model_tflops": round(value * fb / 1e12, 2),
            "roofline": roofline, "cpu_baseline": cpu,
        }
        if sync is not None:
            line["config"]["allreduce_buckets_per_step"] = sync.launched // max(1, args.steps * 2 + args.warmup + 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        # recorded NCCL kernels (B200_GRAPH_DP=1) keep the communicator busy until the graph object is gone
        graphed = step = None  # noqa: F841
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
