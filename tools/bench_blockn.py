"""Dev tool (GPU box, development library): forced UMMA N (b200_dev_set_variant(3, n)) for the mid-level 3x3x3
convolutions — 256-column CTA-pair tiles leave the last wave of clusters partly idle at the 32^3 level."""
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


def timeit(fn, reps=9):
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


CASES = [("fprop", 2, 32, 128, 256), ("fprop", 2, 32, 256, 256), ("fprop", 2, 32, 512, 256),
         ("dgrad", 2, 32, 256, 256), ("dgrad", 2, 32, 256, 512),
         ("fprop", 2, 16, 256, 512), ("fprop", 2, 16, 512, 512), ("fprop", 2, 16, 1024, 512), ("dgrad", 2, 16, 512, 512)]
for kind, n, e, k, ncol in CASES:
    xin = ops.ActView(torch.randn(n, e, e, e, k, device=dev).to(torch.bfloat16))
    out = ops.ActView(ops.new_act(n, e, e, e, ncol, dev))
    gf = 2.0 * n * e ** 3 * k * ncol * 27 / 1e9
    if kind == "fprop":
        wf = (torch.randn(27, ncol, k, device=dev) * 0.05).to(torch.bfloat16)
        b = torch.randn(ncol, device=dev) * 0.1
        stats = torch.empty(2 * 148, ncol, 2, device=dev)
        fn = lambda: ops.conv3d_fprop(xin, wf, b, out, stats, ops.EPI_BIAS_STATS)  # noqa: E731
    else:
        wf = (torch.randn(27, k, ncol, device=dev) * 0.05).to(torch.bfloat16)
        fn = lambda: ops.conv3d_dgrad(xin, wf, out)  # noqa: E731
    ref = None
    for v in (0, 256, 128, 64):
        lib.b200_dev_set_variant(3, v)
        try:
            ms = timeit(fn)
        except Exception as ex:   # noqa: BLE001
            print(f"{kind} K={k} N={ncol} @{e}^3 block_n {v}: {ex}")
            continue
        torch.cuda.synchronize()
        got = out.t.float().clone()
        note = ""
        if ref is None:
            ref = got
        else:
            note = f" max |diff| vs default {(got - ref).abs().max().item():.2e}"
        print(f"{kind} K={k} N={ncol} @{e}^3 block_n {v or 'default'}: {ms:.4f} ms ({gf / ms:.0f} TFLOP/s){note}",
              flush=True)
lib.b200_dev_set_variant(3, 0)
