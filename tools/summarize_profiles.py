"""Turns the raw gpurun_out/ artefacts of tools/collect_profiles.sh into the tracked summaries under profiles/.
usage: python tools/summarize_profiles.py TAG        (run in the build container; needs `ncu` for the .ncu-rep files)"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

for name in ("bench.json", "bench_reference_arm.json", "bench_infer.json", "bench_cv.json", "gemm_launches.txt",
             "timeline_events.txt", "launches.csv", "bn_passes_timing.txt", "ab_fused.txt", "ab_pdl.txt",
             "dmarch2_bench.txt", "conv1_march_bench.txt"):
    src = os.path.join(G, f"{tag}_{name}")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f"{tag}_{name}"))

# ---- launch list -> per-kernel shares of one step
src = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(src):
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v if r[ui] == "us" else v * 1e3)
        seq.append((r[ki].split("(")[0][:60], v))
    adam = [i for i, s in enumerate(seq) if "adam" in s[0]]
    if len(adam) >= 2:
        a, b = adam[0] + 1, adam[1] + 1
        agg = collections.OrderedDict()
        for n, v in seq[a:b]:
            agg.setdefault(n, [0, 0.0])
            agg[n][0] += 1
            agg[n][1] += v
        tot = sum(v for _, v in agg.values())
        with open(os.path.join(P, f"{tag}_launch_summary.txt"), "w") as f:
            f.write(f"# one training step (2x5x128^3, base 64) from profiles/{tag}_launches.csv: "
                    "ncu --metrics gpu__time_duration.sum --clock-control none\n"
                    "# (per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)\n"
                    f"# launches in step {b - a}, sum of kernel durations {tot / 1e3:.3f} ms\n")
            for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"{n:64s} {c:4d} {v:10.1f} us {100 * v / tot:5.1f}%\n")

# ---- full captures -> key metrics
WANT = ["gpu__time_duration.sum", "launch__grid_size", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum"]
traffic = {}
WANT += ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for name in ("igemm_pair", "dmarch_pair", "dmarch2", "wgrad_halo", "igemm", "dmarch", "igemm_im2col5", "wgrad_im2col5",
             "conv1_march", "conv1_march_wgrad", "bn_passes"):
    rep = os.path.join(G, f"{tag}_prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    with open(os.path.join(P, f"{tag}_ncu_{name}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, kernel {name}, from gpurun_out/{tag}_prof_{name}.ncu-rep\n")
        for li, r in enumerate(rows[2:]):
            f.write(f"## captured launch {li}: {r[h.index('Kernel Name')]} grid {r[h.index('Grid Size')]}\n")
            vals = {}
            for j, col in enumerate(h):
                short = col.split(".", 2)[-1] if col.startswith(("TPC.", "SM_", "LTS.", "SYSLTS.")) else col
                for wname in WANT:
                    if short == wname or col == wname:
                        vals[wname] = (r[j], units[j])
            for wname in WANT:
                if wname in vals:
                    f.write(f"{wname:90s} {vals[wname][0]:>16s} {vals[wname][1]}\n")

            def num(k):
                v, u = vals[k]
                v = float(v.replace(",", ""))
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
            if li == 0 and "dram__bytes_read.sum" in vals and name != "bn_passes":
                kname = name + "_kernel"
                traffic[kname] = {"dram_bytes": int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")),
                                  "duration_ms_under_ncu": float(vals["gpu__time_duration.sum"][0]) *
                                  {"ms": 1.0, "us": 1e-3, "ns": 1e-6}[vals["gpu__time_duration.sum"][1]],
                                  "source": f"profiles/{tag}_ncu_{name}.txt, captured launch 0"}
if traffic:
    path = os.path.join(P, "ncu_traffic.json")
    old = {}
    if os.path.exists(path):
        old = json.load(open(path))
    for k, v in traffic.items():
        prev = old.get(k, {}) if isinstance(old.get(k), dict) else {}
        prev.update(v)
        old[k] = prev
    old["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one captured launch per kernel "
                       "(ncu --set full, profiles/*_ncu_*.txt); `launch` / `algorithmic_bytes` are filled in by hand")
    json.dump(old, open(path, "w"), indent=1)
print("profiles updated:", sorted(os.listdir(P)))
