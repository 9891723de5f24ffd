"""Dev tool: per-op CUDA-event timeline of one training step (start ms, duration ms, stream, op) — shows which ops
of the main and the side stream actually overlap.  usage: python tools/timeline_step.py out.txt"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
ops = pkg.ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = pkg.UNet3D(5, 1).to(dev).train()
opt = pkg.FusedAdam(model, lr=1e-4, weight_decay=1e-5)
crit = pkg.BCEDiceLoss()
x = torch.randn(2, 5, 128, 128, 128, device=dev)
y = (torch.rand(2, 1, 128, 128, 128, device=dev) < 0.1).float()


def step():
    opt.zero_grad()
    loss = crit(model(x), y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
base = torch.cuda.Event(enable_timing=True)
base.record()
ops.timeline = []
step()
torch.cuda.synchronize()
tl, ops.timeline = ops.timeline, None
streams = {}
with open(sys.argv[1], "w") as f:
    for name, st, e0, e1 in tl:
        sid = streams.setdefault(st, len(streams))
        f.write(f"{base.elapsed_time(e0):9.3f} {e0.elapsed_time(e1):8.3f} s{sid} {name}\n")
print(len(tl), "ops")
