#!/usr/bin/env python
"""Runs ON THE GPU BOX: full-size parity tables (tests/parity_util.py) at the BASELINE.json shapes, written to
gpurun_out/<tag>_parity_*.txt; copy them to profiles/ after reading.   python tools/parity_fullsize.py [tag] [cases]"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import parity_util as pu  # noqa: E402

CASES = {
    "quick64": dict(batch=1, size=(64, 64, 64), base=64),                        # configs[0] shape on the GPU
    "cfg1_2x128": dict(batch=2, size=(128, 128, 128), base=64),                  # configs[1]
    "cfg3_window": dict(batch=1, size=(128, 128, 64), base=64),                  # configs[3]: one sliding window
    "cfg4_160_b32": dict(batch=1, size=(160, 160, 160), base=32, zero_fill=True),  # configs[4]
    "cfg4_160_b64": dict(batch=1, size=(160, 160, 160), base=64, zero_fill=True),
    "cube32": dict(batch=1, size=(32, 32, 32), base=64),
}

if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    cases = sys.argv[2].split(",") if len(sys.argv) > 2 else list(CASES)
    pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
    dev = torch.device("cuda:0")
    out = os.path.join(ROOT, "gpurun_out")
    summary = {}
    for name in cases:
        res = pu.train_step_parity(pkg, dev, **CASES[name])
        pu.write_report(res, os.path.join(out, f"{tag}_parity_{name}.txt"))
        summary[name] = pu.summarize(res)
        print(name, json.dumps(summary[name]), flush=True)
        torch.cuda.empty_cache()
    with open(os.path.join(out, f"{tag}_parity_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    # layer-by-layer check with identical inputs at the same shapes
    for name in cases:
        rows = pu.layerwise_check(pkg, dev, **CASES[name])
        pu.write_layerwise(rows, os.path.join(out, f"{tag}_layerwise_{name}.txt"), header=f"{name}: {CASES[name]}")
        worst = max((r for r in rows if not r[1].endswith("mismatches") and not r[1].endswith("(info)") and r[0] != "loss"), key=lambda r: r[2])
        print(name, "layerwise worst", worst, flush=True)
        torch.cuda.empty_cache()
