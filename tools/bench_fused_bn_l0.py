"""ncu target: one launch each of the BatchNorm-pass kernels at 2 x 128^3 x 64 (see bench_fused_bn.py)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
ops = pkg.ops
dev = torch.device("cuda:0")
n, c, e = 2, 64, 128
act = lambda *s: ops.ActView(torch.randn(*s, device=dev).to(torch.bfloat16))  # noqa: E731
y, dskip, dout, dy, a = (act(n, e, e, e, c) for _ in range(5))
dpool = act(n, e // 2, e // 2, e // 2, c)
scale, shift, mean, rstd, gamma = (torch.rand(c, device=dev) + 0.5 for _ in range(5))
partial = torch.empty(ops.bn_bwd_max_blocks(), c, 2, device=dev)
coef = torch.empty(c, 2, device=dev)
dgamma, dbeta, dbias = (torch.zeros(c, device=dev) for _ in range(3))
bn = (y, scale, shift, mean, rstd, gamma, partial, coef, dgamma, dbeta, dy, dbias)
w = torch.randn(1, c, device=dev) * 0.1
dl = torch.randn(n, 1, e, e, e, device=dev)
dw, db = torch.zeros(1, c, device=dev), torch.zeros(1, device=dev)
pooled = act(n, e // 2, e // 2, e // 2, c)
for _ in range(2):   # 9 launches per round: apply, apply+pool, pool bwd, bn_bwd (3), bn_bwd_head (3)
    ops.bn_apply_relu(y, scale, shift, a)
    ops.bn_apply_relu_pool(y, scale, shift, a, pooled)
    ops.maxpool3d_bwd(a, dpool, dskip, dout)
    ops.bn_bwd(dout, *bn)
    ops.bn_bwd_head(dl, w, *bn, dw, db)
torch.cuda.synchronize()
print("ok")
