"""Dev tool (GPU box): time the first-layer kernels (direct form vs materialised im2col) at 2x5x128^3, base 64."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
ops = pkg.ops
dev = torch.device("cuda:0")
n, d, h, w, cout = 2, 128, 128, 128, 64
x = torch.randn(n, 5, d, h, w, device=dev)
wt = torch.randn(cout, 5, 3, 3, 3, device=dev) * 0.1
b = torch.zeros(cout, device=dev)
w_rows = torch.empty(cout, 144, device=dev, dtype=torch.bfloat16)
ops.pack_rows(wt.contiguous(), 144, w_rows)
y = ops.ActView(ops.new_act(n, d, h, w, cout, dev))
dy = ops.ActView(torch.randn(n, d, h, w, cout, device=dev).to(torch.bfloat16))
rows = max(ops.conv3d_stat_rows(n, d, h, w, cout, 1), ops.conv1_direct_stat_rows(n, d, h, w, cout))
stats = torch.empty(rows, cout, 2, device=dev)
dw = torch.zeros(cout, 135, device=dev)
rv = ops.ActView(ops.new_act(n, d, h, w, 144, dev))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(name, fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name}: median {ts[len(ts) // 2]:.4f} ms  min {ts[0]:.4f}")


timeit("direct fprop (train epilogue)", lambda: ops.conv1_direct_fprop(x, w_rows, b, y, stats, ops.EPI_BIAS_STATS))
timeit("direct wgrad", lambda: ops.conv1_direct_wgrad(x, dy, dw))
timeit("im2col_input", lambda: ops.im2col_input(x, rv))
timeit("im2col fprop", lambda: ops.conv1_fprop(rv, w_rows, b, y, stats, ops.EPI_BIAS_STATS, k_real=135))
timeit("im2col wgrad", lambda: ops.conv1_wgrad(rv, dy, dw, 135))
scale, shift = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
timeit("direct fprop (eval epilogue: affine + relu, no statistics)",
       lambda: ops.conv1_direct_fprop(x, w_rows, None, y, None, ops.EPI_AFFINE_RELU, scale, shift))
