"""Not a pytest module: times the oracle's training step (the reference graph as stock torch ops, oracle/unet3d_oracle.py)
on the GPU through torch's own libraries — cuDNN 3-D convolutions, ATen BatchNorm / pooling / loss, torch.optim-style
Adam — at BASELINE configs[1] (2 x 5 x 128^3, base 64).  This is the "existing Blackwell kernels" bar of SURVEY 8(d); it
is reported in profiles/r1_notes.md next to the B200-native path and is never on the product path.
usage: python tests/library_baseline.py [out.json]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import importlib  # noqa: E402

import unet3d_oracle as oracle  # noqa: E402

pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = pkg.UNet3D(5, 1)                       # parameter container only (seed-identical init)
sd = {k: v.detach().clone().to(dev) for k, v in model.state_dict().items()}
del model
g = torch.Generator().manual_seed(1234)
x = torch.randn(2, 5, 128, 128, 128, generator=g).to(dev)
y = (torch.rand(2, 1, 128, 128, 128, generator=g) < 0.1).float().to(dev)
out = {}
for label, autocast, cl3d in (("fp32", False, False), ("bf16_autocast", True, False),
                              ("bf16_autocast_channels_last_3d", True, True)):
    try:
        xx = x.contiguous(memory_format=torch.channels_last_3d) if cl3d else x
        state = {}
        work = {k: v.clone() for k, v in sd.items()}
        if cl3d:
            work = {k: (v.contiguous(memory_format=torch.channels_last_3d) if v.dim() == 5 else v)
                    for k, v in work.items()}

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                return oracle.train_step(work, state, xx, y)

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 5
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[label] = {"ms_per_step": ms, "voxels_per_s": 2 * 128 ** 3 / (ms * 1e-3)}
    except Exception as e:  # report (e.g. an out-of-memory or unsupported layout), this is a side measurement
        out[label] = {"error": repr(e)[:300]}
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
