"""Dev tool: host-side enqueue time of one training step (no sync inside) vs its GPU time."""
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


for _ in range(3):
    step()
torch.cuda.synchronize()
for trial in range(3):
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host enqueue {1e3 * (t1 - t0) / 5:.2f} ms/step, total {1e3 * (t2 - t0) / 5:.2f} ms/step")
# with a sync every step (the e2e pattern)
for trial in range(3):
    t0 = time.perf_counter()
    for _ in range(5):
        step().item()
    t2 = time.perf_counter()
    print(f"item() every step: {1e3 * (t2 - t0) / 5:.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
