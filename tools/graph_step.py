"""Dev experiment: time the training step replayed from a CUDA graph against the eager launch sequence
(how much of the step is launch gaps?).  The captured Adam bias correction is frozen at the capture step: timing only."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
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
    return loss


def timed(fn, n=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
print(f"eager: {timed(step):.3f} ms/step")
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g):
        static_loss = step()
    g.replay()
    torch.cuda.synchronize()
    print(f"graph: {timed(g.replay):.3f} ms/step, loss {static_loss.item():.5f}")
except Exception as e:  # report, this is an experiment
    print("capture failed:", repr(e)[:2000])
print(f"eager again: {timed(step):.3f} ms/step")
