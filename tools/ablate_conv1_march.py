"""Dev tool (GPU box, development library): which stage bounds the first layer's marching kernels at 2x5x128^3 -> 64?
ablation 1 = no output stores (forward) / dy slots armed without loads (weight gradient), 2 = builders skip the input
loads, 3 = no MMAs.  Results are garbage, timings are not."""
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
w_sl = torch.empty(3, cout, 64, device=dev, dtype=torch.bfloat16)
ops.pack_conv1_slices(wt, w_sl)
y = ops.ActView(torch.randn(n, d, h, w, cout, device=dev).to(torch.bfloat16))
stats = torch.empty(ops.conv1_march_stat_rows(n, d, h, w, cout), cout, 2, device=dev)
dw = torch.zeros(cout, 135, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for abl, what in ((0, "full kernel"), (1, "no output stores / no dy loads"), (2, "no input loads"), (3, "no MMAs")):
    lib.b200_dev_set_variant(4, abl)
    tf = timed(lambda: ops.conv1_march_fprop(x, w_sl, b, y, stats, ops.EPI_BIAS_STATS))
    tw = timed(lambda: ops.conv1_march_wgrad(x, y, dw))
    print(f"ablation {abl} ({what}): forward {tf:.4f} ms, weight gradient {tw:.4f} ms", flush=True)
lib.b200_dev_set_variant(4, 0)
