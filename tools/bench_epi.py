"""Dev tool (GPU box, development library): the staged epilogue (shared-memory tile + TMA store, statistics read back
from the tile) against the direct epilogue for 128-column tiles, forward and input-gradient, at the network's shapes.
b200_dev_set_variant(2, v): 0 = product choice (staged, one tile), 1 = direct epilogue, 2 = staged with two tiles."""
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


# (kind, n, extent, cin (K), cout (N columns))
CASES = [("fprop", 2, 64, 64, 128), ("fprop", 2, 64, 128, 128), ("fprop", 2, 64, 256, 128),
         ("dgrad", 2, 128, 64, 128), ("dgrad", 2, 64, 128, 128)]
for kind, n, e, k, ncol in CASES:
    xin = ops.ActView(torch.randn(n, e, e, e, k, device=dev).to(torch.bfloat16))
    out = ops.ActView(ops.new_act(n, e, e, e, ncol, dev))
    gf = 2.0 * n * e ** 3 * k * ncol * 27 / 1e9
    if kind == "fprop":
        wf = (torch.randn(27, ncol, k, device=dev) * 0.05).to(torch.bfloat16)
        b = torch.randn(ncol, device=dev) * 0.1
        stats = torch.empty(ops.conv3d_stat_rows(n, e, e, e, ncol), ncol, 2, device=dev)
        fn = lambda: ops.conv3d_fprop(xin, wf, b, out, stats, ops.EPI_BIAS_STATS)  # noqa: E731
    else:   # dx (ncol channels) from dy (k channels): weights packed [27][Cout = k][Cin = ncol]
        wf = (torch.randn(27, k, ncol, device=dev) * 0.05).to(torch.bfloat16)
        stats = None
        fn = lambda: ops.conv3d_dgrad(xin, wf, out)  # noqa: E731
    ref = None
    for v in (0, 1, 2):
        lib.b200_dev_set_variant(2, v)
        try:
            ms = timeit(fn)
        except Exception as ex:   # noqa: BLE001
            print(f"{kind} K={k} N={ncol} @{e}^3 variant {v}: {ex}")
            continue
        torch.cuda.synchronize()
        got = out.t.float().clone()
        st = stats.double().sum(0).clone() if stats is not None else None
        note = ""
        if ref is None:
            ref = (got, st)
        else:
            note = f" equal to variant 0: {torch.equal(got, ref[0])}"
            if st is not None:
                note += f", stats rel diff {((st - ref[1]).abs().max() / ref[1].abs().max()).item():.1e}"
        print(f"{kind} K={k} N={ncol} @{e}^3 variant {v}: {ms:.4f} ms ({gf / ms:.0f} TFLOP/s){note}", flush=True)
lib.b200_dev_set_variant(2, 0)
