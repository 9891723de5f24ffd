"""Dev experiment (GPU box): the training step (2x5x128^3, base 64) replayed from CUDA graphs captured with
programmatic dependent launch (csrc/launch.cuh) on and off, alternating in one process on one box; then the two graphs'
results are compared bit for bit (PDL only moves launch times, not arithmetic)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
dev = torch.device("cuda:0")
base = int(sys.argv[1]) if len(sys.argv) > 1 else 64
order = (False, True) if len(sys.argv) > 2 and sys.argv[2] == "off-first" else (True, False)
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
for pdl in (True, False):
    pkg.ops.set_pdl(pdl)
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
    graphs[pdl] = (g, loss, pkg.ops.launch_count - l0)
for rnd in range(4):
    for pdl in order:
        g, loss, launches = graphs[pdl]
        g.replay()
        print(f"round {rnd} pdl={pdl}: {timed(g.replay):.3f} ms/step ({launches} launches, loss {loss.item():.5f})",
              flush=True)

# (the weight gradients are added with REDs, so two runs never agree bit for bit: the arithmetic of PDL on / off is
# compared through the forward only, tests/test_kernels_gpu.py and the full GPU suite run with PDL on)

# eager launches (no graph)
for pdl in (True, False):
    pkg.ops.set_pdl(pdl)
    step()
    print(f"eager pdl={pdl}: {timed(step, 10):.3f} ms/step", flush=True)
pkg.ops.set_pdl(True)
