"""Prints the markdown tables of profiles/rN_notes.md from profiles/rN_bench*.json (so the notes quote the files).
usage: python tools/notes_tables.py r2"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
d = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_bench.json")))
r = d["roofline"]
print(f"value {d['value'] / 1e6:.1f} M voxels/s ({d['ms_per_step']:.2f} ms/step), e2e {d['e2e']['value'] / 1e6:.1f} M "
      f"({d['e2e']['ms_per_step']:.2f} ms), clocks {d['clocks']}, launches/step {d['gpu_launches'] // d['steps']}, "
      f"model TFLOP/s {d['model_tflops']}")
print(f"roofline: {r['kernel']} {r['achieved']} TFLOP/s, frac {r['frac']} of {r['peak']}, per pass {r['achieved_per_pass']}, "
      f"gemm share {r['gemm_share_of_step']}, of nominal {r['frac_of_nominal_2250']}")
print("\n| kernel | launches | ms / step | TFLOP/s |\n|---|---|---|---|")
for k, v in sorted(r["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print(f"| `{k}` | {v['launches']} | {v['ms_per_step']:.2f} | {v['tflops']:.0f} |")
g = d["gpu_library_baseline"]
print(f"\nlibrary: bf16 {g['bf16_autocast_channels_last_3d']['ms_per_step']:.1f} ms, fp32 {g['fp32']['ms_per_step']:.1f} ms, "
      f"speedup {g['speedup_vs_best_library_config']}, per-layer sums {g['per_layer_sum_ms']}")
print("\n| layer (Cin→Cout @ D×H×W, batch 2) | cuDNN fwd | cuDNN bwd | B200 fwd | B200 bwd | speed-up |\n|---|---|---|---|---|---|")
for k, v in g["per_layer_3x3x3"].items():
    sp = (v["cudnn_fwd_ms"] + v["cudnn_bwd_ms"]) / (v["b200_fwd_ms"] + v["b200_bwd_ms"])
    print(f"| {k} | {v['cudnn_fwd_ms']:.3f} | {v['cudnn_bwd_ms']:.3f} | {v['b200_fwd_ms']:.3f} | {v['b200_bwd_ms']:.3f} | "
          f"{sp:.2f}× |")
print("\ncpu:", d["cpu_baseline"])
for name in ("bench_infer", "bench_cv", "bench_reference_arm"):
    p = os.path.join(ROOT, "profiles", f"{tag}_{name}.json")
    if os.path.exists(p):
        x = json.load(open(p))
        keep = {k: x[k] for k in ("value", "ms_per_step", "e2e", "clocks") if k in x}
        extra = {k: v for k, v in x.get("config", {}).items() if k in ("base32", "base64", "per_base")}
        print(f"\n{name}: {keep} {extra}")
        if "roofline" in x and x["roofline"]:
            print("   roofline:", {k: v for k, v in x["roofline"].items() if k in ("kernel", "achieved", "frac", "peak")})
        if "cpu_baseline" in x and x["cpu_baseline"]:
            print("   cpu:", x["cpu_baseline"])
