import importlib, sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
par = importlib.import_module("prostate-cancer-multimodal-segmentation_b200.parallel")
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = pkg.UNet3D(5, 1).to(dev).eval()
x = torch.rand(1, 5, 256, 256, 64, device=dev)
for wpl in (1, 3, 9):
    for _ in range(3):
        par.sliding_window_predict(model, x, (128, 128, 64), (64, 64, 64), windows_per_launch=wpl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        par.sliding_window_predict(model, x, (128, 128, 64), (64, 64, 64), windows_per_launch=wpl)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"windows_per_launch {wpl}: {ms:.2f} ms/volume, {256*256*64/ms/1e3:.1f} M voxels/s")
