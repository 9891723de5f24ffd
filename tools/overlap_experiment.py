"""Dev tool (GPU box): how much of the backward's HBM-bound work hides under the weight-gradient GEMMs?
Times the graph-replayed step with the side stream on / off and prints a per-phase summary of the event timeline."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
ops = pkg.ops
dev = torch.device("cuda:0")
x = torch.randn(2, 5, 128, 128, 128, device=dev)
y = (torch.rand(2, 1, 128, 128, 128, device=dev) < 0.1).float()
GEMM = ("conv3d_fprop", "conv3d_dgrad", "conv3d_wgrad", "conv1_direct_fprop", "conv1_direct_wgrad", "convt2x_fwd",
        "convt2x_dgrad", "convt2x_wgrad", "conv1_fprop", "conv1_wgrad")
for overlap in (True, False):
    torch.manual_seed(0)
    model = pkg.UNet3D(5, 1).to(dev).train()
    model.engine.overlap_wgrad = overlap
    opt = pkg.FusedAdam(model, lr=1e-4, weight_decay=1e-5)
    crit = pkg.BCEDiceLoss()
    g = pkg.GraphedTrainStep(model, crit, opt)
    for _ in range(5):
        g(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g(x, y)
    e1.record()
    torch.cuda.synchronize()
    print(f"overlap_wgrad={overlap}: graph replay {e0.elapsed_time(e1) / 10:.3f} ms/step (replays {g.replays})")
    # eager timeline
    def step():
        opt.zero_grad(); loss = crit(model(x), y); loss.backward(); opt.step()
    step(); torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True); base.record()
    ops.timeline = []
    step(); torch.cuda.synchronize()
    tl, ops.timeline = ops.timeline, None
    rows = [(base.elapsed_time(a), a.elapsed_time(b), st, name) for name, st, a, b in tl]
    main = rows[0][2]
    t_end = max(r[0] + r[1] for r in rows)
    first_bwd = next(r[0] for r in rows if r[3] == "loss_bwd")
    gemm_f = sum(r[1] for r in rows if r[3] in GEMM and r[0] < first_bwd)
    bw_f = sum(r[1] for r in rows if r[3] not in GEMM and r[0] < first_bwd)
    gemm_b = sum(r[1] for r in rows if r[3] in GEMM and r[0] >= first_bwd)
    bw_b = sum(r[1] for r in rows if r[3] not in GEMM and r[0] >= first_bwd)
    side = sum(r[1] for r in rows if r[2] != main)
    print(f"   eager step {t_end - rows[0][0]:.2f} ms: forward {first_bwd - rows[0][0]:.2f} (GEMM {gemm_f:.2f} + other {bw_f:.2f}); "
          f"backward+adam {t_end - first_bwd:.2f} (GEMM {gemm_b:.2f} of which side stream {side:.2f}, other {bw_b:.2f})")
    del g, model, opt
    torch.cuda.empty_cache()
