"""Dev tool (GPU box): the 64-column full-resolution convolutions on single-CTA MMAs with a multicast weight stream
(dmarch_pair_kernel) vs CTA-pair MMAs (dmarch2_kernel): results against each other and against F.conv3d on a small
shape, then event-timed at the step's shapes."""
import importlib, os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
ops = pkg.ops
from helpers import bf16_round, from_act, rel_l2, to_act, empty_act
dev = torch.device("cuda:0")


def run(x, wt, b, mode, pair_mma, stats_rows=None):
    n, cin, d, h, w = x.shape
    cout = wt.shape[0]
    ops.set_dmarch_pair_mma(pair_mma)
    xv = to_act(ops, x)
    wf = torch.empty(27, cout, cin, device=dev, dtype=torch.bfloat16)
    ops.pack_conv_weight(wt.contiguous(), cin, wf)
    y = empty_act(ops, n, cout, d, h, w, dev)
    rows = ops.conv3d_stat_rows(n, d, h, w, cout)
    stats = torch.zeros(rows, cout, 2, device=dev)
    ops.conv3d_fprop(xv, wf, b, y, stats, mode)
    torch.cuda.synchronize()
    return from_act(y), stats.double().sum(0)


for (n, cin, cout, d, h, w) in [(1, 64, 64, 8, 16, 8), (2, 64, 64, 9, 20, 12), (1, 128, 64, 21, 32, 16), (1, 32, 32, 16, 16, 16),
                                (3, 64, 64, 33, 16, 24)]:
    g = torch.Generator().manual_seed(5)
    x = bf16_round(torch.randn(n, cin, d, h, w, generator=g)).to(dev)
    wt = bf16_round(torch.randn(cout, cin, 3, 3, 3, generator=g) * (2.0 / (27 * cin)) ** 0.5).to(dev)
    b = (torch.randn(cout, generator=g) * 0.1).to(dev)
    ref = F.conv3d(x, wt, b, padding=1)
    y1, s1 = run(x, wt, b, ops.EPI_BIAS_STATS, False)
    y2, s2 = run(x, wt, b, ops.EPI_BIAS_STATS, True)
    print(f"{n}x{cin}->{cout}@{d}x{h}x{w}: old vs torch {rel_l2(y1, ref):.2e}, pair-mma vs torch {rel_l2(y2, ref):.2e}, "
          f"pair-mma vs old {rel_l2(y2, y1):.2e}, finite {bool(torch.isfinite(y2).all())}, "
          f"stats sum diff {float((s2[:, 0] - y2.double().sum((0, 2, 3, 4))).abs().max()):.3e}", flush=True)


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


for (n, cin, cout, d, h, w) in [(2, 64, 64, 128, 128, 128), (2, 128, 64, 128, 128, 128), (1, 64, 64, 160, 160, 160)]:
    xv = ops.ActView(torch.randn(n, d, h, w, cin, device=dev).to(torch.bfloat16))
    wf = (torch.randn(27, cout, cin, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.randn(cout, device=dev) * 0.1
    y = ops.ActView(torch.empty(n, d, h, w, cout, device=dev, dtype=torch.bfloat16))
    stats = torch.zeros(ops.conv3d_stat_rows(n, d, h, w, cout), cout, 2, device=dev)
    fl = 2.0 * n * d * h * w * cin * cout * 27
    res = {}
    for rnd in range(2):
        for pm in (False, True):
            ops.set_dmarch_pair_mma(pm)
            res[pm] = timed(lambda: ops.conv3d_fprop(xv, wf, b, y, stats, ops.EPI_BIAS_STATS))
        print(f"{n}x{cin}->{cout}@{d}x{h}x{w} round {rnd}: multicast {res[False]:.4f} ms ({fl / res[False] / 1e9:.0f} TFLOP/s), "
              f"pair-mma {res[True]:.4f} ms ({fl / res[True] / 1e9:.0f} TFLOP/s)", flush=True)
ops.set_dmarch_pair_mma(True)
