"""Dev tool (GPU box): weight gradients of the full-resolution layers on wgrad_halo_kernel, event-timed, isolated."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
ops = pkg.ops
dev = torch.device("cuda:0")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (n, cin, cout, d, h, w) in [(2, 128, 64, 128, 128, 128), (2, 64, 128, 64, 64, 64), (2, 64, 64, 128, 128, 128),
                                (2, 128, 128, 64, 64, 64), (1, 128, 64, 160, 160, 160)]:
    x = ops.ActView(torch.randn(n, d, h, w, cin, device=dev).to(torch.bfloat16))
    dy = ops.ActView(torch.randn(n, d, h, w, cout, device=dev).to(torch.bfloat16))
    dw = torch.zeros(27, cout, cin, device=dev)
    t = timed(lambda: ops.conv3d_wgrad(x, dy, dw, cin, packed=True))
    fl = 2.0 * n * d * h * w * cin * cout * 27
    print(f"wgrad {cin}->{cout} @ {n}x{d}x{h}x{w}: {t:.4f} ms ({fl / t / 1e9:.0f} TFLOP/s)", flush=True)
