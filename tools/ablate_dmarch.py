"""Dev tool (GPU box, development library): the 64-column full-resolution layers (depth-marching kernel) at 2 x 128^3
with the epilogue reduced to the accumulator handshake (ablation 4), without loads (1) and without MMAs (2)."""
import importlib, os, sys
os.environ["B200_DEV"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
ops = pkg.ops
lib = pkg.load_library()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=7):
    for _ in range(2):
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


n, e = 2, 128
for cin, cout in ((64, 64), (128, 64)):
    x = ops.ActView(torch.randn(n, e, e, e, cin, device=dev).to(torch.bfloat16))
    y = ops.ActView(ops.new_act(n, e, e, e, cout, dev))
    wf = (torch.randn(27, cout, cin, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.zeros(cout, device=dev)
    stats = torch.empty(ops.conv3d_stat_rows(n, e, e, e, cout), cout, 2, device=dev)
    scale, shift = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    gf = 2.0 * n * e ** 3 * cin * cout * 27 / 1e9
    for ab, what in ((0, "full kernel"), (4, "epilogue does nothing"), (1, "no loads"), (2, "no MMAs")):
        lib.b200_dev_set_ablation(0, ab, 0, 0)
        t = timeit(lambda: ops.conv3d_fprop(x, wf, b, y, stats, ops.EPI_BIAS_STATS))
        ev = timeit(lambda: ops.conv3d_fprop(x, wf, None, y, None, ops.EPI_AFFINE_RELU, scale, shift))
        print(f"{cin}->{cout} ablation {ab} ({what}): train epilogue {t:.4f} ms ({gf / t:.0f} TFLOP/s), "
              f"eval epilogue {ev:.4f} ms", flush=True)
lib.b200_dev_set_ablation(0, 0, 0, 0)
