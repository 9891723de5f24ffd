"""Dev tool (GPU box, B200_DEV=1): time launcher variants of the transposed-convolution forward (csrc/api.cu,
b200_dev_set_variant) at the four decoder shapes of 2x5x128^3 base 64, checking every variant against torch."""
import importlib, os, sys
os.environ["B200_DEV"] = "1"
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
ops = pkg.ops
lib = pkg.load_library()
dev = torch.device("cuda:0")
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


SHAPES = [(2, 128, 64), (2, 256, 32), (2, 512, 16), (2, 1024, 8)]   # n, cin, extent of the input
VARIANTS = [(0, 0), (256, 1), (128, 1), (128, 2), (64, 2), (64, 1)]
for n, cin, e in SHAPES:
    cout = cin // 2
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn(n, cin, e, e, e, generator=g).to(torch.bfloat16).float().to(dev)
    wt = (torch.randn(cin, cout, 2, 2, 2, generator=g) * (1.0 / cin) ** 0.5).to(torch.bfloat16).float().to(dev)
    b = (torch.randn(cout, generator=g) * 0.1).to(dev)
    wf = torch.empty(8 * cout, cin, device=dev, dtype=torch.bfloat16)
    wd = torch.empty(8, cin, cout, device=dev, dtype=torch.bfloat16)
    b8 = torch.empty(8 * cout, device=dev)
    ops.pack_convt_weight(wt.contiguous(), b, wf, wd, b8)
    xa = ops.ActView(x.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16))
    cat = torch.zeros(n, 2 * e, 2 * e, 2 * e, 2 * cout, device=dev, dtype=torch.bfloat16)
    up = ops.ActView(cat, cout, cout)
    ref = F.conv_transpose3d(x, wt, b, stride=2) if e <= 32 else None
    for bn, cb in VARIANTS:
        lib.b200_dev_set_variant(0, bn)
        lib.b200_dev_set_variant(1, cb)
        try:
            ms = timeit(lambda: ops.convt2x_fwd(xa, wf, b8, up, (0, 0, 0)))
        except Exception as ex:   # noqa: BLE001
            print(f"convT {cin}->{cout} @{e}^3 block_n {bn} c_bufs {cb}: {ex}")
            continue
        err = ""
        if ref is not None:
            got = cat[..., cout:].permute(0, 4, 1, 2, 3).float()
            err = f" rel-L2 {((got - ref).norm() / ref.norm()).item():.2e}"
        byts = (x.numel() + 8 * x.numel() // 2) * 2
        print(f"convT {cin}->{cout} @{e}^3 block_n {bn or 'default'} c_bufs {cb or 'default'}: {ms:.4f} ms "
              f"({byts / ms / 1e6:.0f} GB/s){err}", flush=True)
