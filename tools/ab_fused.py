"""Dev experiment (GPU box): the training step (2x5x128^3, base 64) replayed from a CUDA graph with the fused
BatchNorm passes (engine.FUSE_BN_PASSES) on and off, alternating in one process on one box."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
eng = importlib.import_module(pkg.__name__ + ".engine")
dev = torch.device("cuda:0")
base = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
model = pkg.UNet3D(5, 1, init_features=base).to(dev).train()
opt = pkg.FusedAdam(model, lr=1e-4, weight_decay=1e-5)
crit = pkg.BCEDiceLoss()
x = torch.randn(2, 5, 128, 128, 128, device=dev)
y = (torch.rand(2, 1, 128, 128, 128, device=dev) < 0.1).float()


def step():
    opt.zero_grad()
    loss = crit(model(x), y)
    loss.backward()
    opt.step()
    return loss


def timed(fn, n=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


graphs = {}
for fused in (True, False):
    eng.FUSE_BN_PASSES = fused
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    l0 = pkg.ops.launch_count
    opt.refresh_dynamic_scalars(advance=False)
    model.engine._pack_key = None
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss = step()
    graphs[fused] = (g, loss, pkg.ops.launch_count - l0)
for rnd in range(4):
    for fused in (True, False):
        g, loss, launches = graphs[fused]
        g.replay()
        print(f"round {rnd} fused={fused}: {timed(g.replay):.3f} ms/step ({launches} launches, loss {loss.item():.5f})",
              flush=True)

# per-op durations (events around every op, weight gradients NOT on the side stream so nothing overlaps)
model.engine.overlap_wgrad = False
for fused in (True, False):
    eng.FUSE_BN_PASSES = fused
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    pkg.ops.timeline = []
    step()
    torch.cuda.synchronize()
    tl, pkg.ops.timeline = pkg.ops.timeline, None
    agg = {}
    for name, st, e0, e1 in tl:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
    tot = sum(v[1] for v in agg.values())
    print(f"--- fused={fused}: sum of op durations {tot:.3f} ms")
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if not name.startswith("conv") and not name.startswith("convt"):
            print(f"   {name:22s} x{cnt:3d} {ms:7.3f} ms")
