"""Dev tool (GPU box): the fused BatchNorm passes against the unfused chains at the full-resolution shape
(2 x 128^3 x 64 channels) and at 2 x 64^3 x 128, with achieved GB/s over the bytes each form has to move."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
ops = pkg.ops
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


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


for n, c, e in ((2, 64, 128), (2, 128, 64)):
    def act(*shape):
        return ops.ActView(torch.randn(*shape, device=dev).to(torch.bfloat16))
    y, dskip, dout, dy, a = (act(n, e, e, e, c) for _ in range(5))
    dpool, pooled = act(n, e // 2, e // 2, e // 2, c), act(n, e // 2, e // 2, e // 2, c)
    scale, shift, mean, rstd, gamma = (torch.rand(c, device=dev) + 0.5 for _ in range(5))
    partial = torch.empty(ops.bn_bwd_max_blocks(), c, 2, device=dev)
    coef = torch.empty(c, 2, device=dev)
    dgamma, dbeta, dbias = (torch.zeros(c, device=dev) for _ in range(3))
    T = y.t.numel() * 2 / 1e9   # GB of one tensor-sized stream
    bn = (y, scale, shift, mean, rstd, gamma, partial, coef, dgamma, dbeta, dy, dbias)
    rows = [
        ("bn_apply_relu", lambda: ops.bn_apply_relu(y, scale, shift, a), 2 * T),
        ("maxpool3d_fwd", lambda: ops.maxpool3d_fwd(a, pooled), 1.125 * T),
        ("bn_apply_relu_pool", lambda: ops.bn_apply_relu_pool(y, scale, shift, a, pooled), 2.125 * T),
        ("maxpool3d_bwd", lambda: ops.maxpool3d_bwd(a, dpool, dskip, dout), 3.125 * T),
        ("bn_bwd", lambda: ops.bn_bwd(dout, *bn), 5 * T),
    ]
    if c == 64:
        w = torch.randn(1, c, device=dev) * 0.1
        dl = torch.randn(n, 1, e, e, e, device=dev)
        dw, db = torch.zeros(1, c, device=dev), torch.zeros(1, device=dev)
        rows += [("head_bwd", lambda: ops.head_bwd(a, w, dl, dout, dw, db), 2 * T),
                 ("bn_bwd_head", lambda: ops.bn_bwd_head(dl, w, *bn, dw, db), 3 * T + 2 * dl.numel() * 4 / 1e9)]
    for name, fn, gb in rows:
        ms = timeit(fn)
        print(f"{n}x{e}^3x{c} {name:20s} {ms:7.4f} ms  {gb / ms * 1e3:6.0f} GB/s", flush=True)
