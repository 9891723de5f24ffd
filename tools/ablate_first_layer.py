"""Dev tool (GPU box, development library): which stage bounds the direct first-layer forward kernel at 2x5x128^3 ->
64 channels?  ablation 4 = epilogue hands the accumulator back without reading it, 5 = builders skip the input loads."""
import importlib, os, sys
os.environ["B200_DEV"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
ops = pkg.ops
lib = pkg.load_library()
dev = torch.device("cuda:0")
n, d, h, w, cout = 2, 128, 128, 128, 64
x = torch.randn(n, 5, d, h, w, device=dev)
wt = torch.randn(cout, 5, 3, 3, 3, device=dev) * 0.1
b = torch.zeros(cout, device=dev)
w_rows = torch.empty(cout, 144, device=dev, dtype=torch.bfloat16)
ops.pack_rows(wt.contiguous(), 144, w_rows)
y = ops.ActView(ops.new_act(n, d, h, w, cout, dev))
stats = torch.empty(ops.conv1_direct_stat_rows(n, d, h, w, cout), cout, 2, device=dev)
scale, shift = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
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
    return ts[len(ts) // 2]


for ab, what in ((0, "full kernel"), (4, "epilogue does nothing"), (5, "no input loads"), (2, "no MMAs")):
    lib.b200_dev_set_ablation(ab, 0, 0, 0)
    t = timeit(lambda: ops.conv1_direct_fprop(x, w_rows, b, y, stats, ops.EPI_BIAS_STATS))
    e = timeit(lambda: ops.conv1_direct_fprop(x, w_rows, None, y, None, ops.EPI_AFFINE_RELU, scale, shift))
    print(f"ablation {ab} ({what}): train epilogue {t:.4f} ms, eval epilogue {e:.4f} ms", flush=True)
lib.b200_dev_set_ablation(0, 0, 0, 0)
