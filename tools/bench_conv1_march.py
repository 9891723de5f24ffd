"""Dev tool (GPU box): the first layer's forward at 2 x 5 x 128^3 -> 64 (and 1 x 5 x 160^3 -> 32 / 64): generic direct
kernels vs the depth-marching kernels (forward and weight gradient), event-timed, isolated."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
ops = pkg.ops
dev = torch.device("cuda:0")


def timed(fn, n=30):
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


for (n, d, h, w, cout) in [(2, 128, 128, 128, 64), (1, 160, 160, 160, 64), (1, 160, 160, 160, 32), (9, 64, 128, 128, 64)]:
    x = torch.randn(n, 5, d, h, w, device=dev)
    wt = torch.randn(cout, 5, 3, 3, 3, device=dev) * 0.1
    b = torch.randn(cout, device=dev) * 0.1
    w_rows = torch.empty(cout, 144, device=dev, dtype=torch.bfloat16)
    ops.pack_rows(wt, 144, w_rows)
    w_sl = torch.empty(3, cout, 64, device=dev, dtype=torch.bfloat16)
    ops.pack_conv1_slices(wt, w_sl)
    y = ops.ActView(torch.empty(n, d, h, w, cout, device=dev, dtype=torch.bfloat16))
    st_d = torch.empty(ops.conv1_direct_stat_rows(n, d, h, w, cout), cout, 2, device=dev)
    st_m = torch.empty(ops.conv1_march_stat_rows(n, d, h, w, cout), cout, 2, device=dev)
    scale, shift = torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev)
    vox = n * d * h * w
    for name, mode, args in (("train (bias + statistics)", ops.EPI_BIAS_STATS, (None, None)),
                             ("eval (folded BN + ReLU)", ops.EPI_AFFINE_RELU, (scale, shift))):
        t_d = timed(lambda: ops.conv1_direct_fprop(x, w_rows, b, y, st_d if mode == ops.EPI_BIAS_STATS else None, mode, *args))
        t_m = timed(lambda: ops.conv1_march_fprop(x, w_sl, b, y, st_m if mode == ops.EPI_BIAS_STATS else None, mode, *args))
        gb = vox * (20 + 2 * cout) / 1e9
        if mode == ops.EPI_BIAS_STATS:
            dw = torch.zeros(cout, 135, device=dev)
            t_dw = timed(lambda: ops.conv1_direct_wgrad(x, y, dw))
            t_mw = timed(lambda: ops.conv1_march_wgrad(x, y, dw))
            print(f"{n}x5x{d}x{h}x{w} -> {cout}, weight gradient: direct {t_dw:.4f} ms, march {t_mw:.4f} ms "
                  f"({2.0 * vox * cout * 135 / t_mw / 1e9:.0f} TFLOP/s)", flush=True)
        print(f"{n}x5x{d}x{h}x{w} -> {cout}, {name}: direct {t_d:.4f} ms, march {t_m:.4f} ms "
              f"({gb / t_m * 1e3:.0f} GB/s of algorithmic traffic, {2.0 * vox * cout * 135 / t_m / 1e9:.0f} TFLOP/s)", flush=True)
