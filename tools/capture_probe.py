"""Dev probe (GPU box): which call patterns around GraphedTrainStep break CUDA-graph capture.  Variant D keeps the previous eager step's loss (with its autograd graph) alive across the capturing call, K discards it, L detaches it; A/B/C/E run eval forwards first, F-I toggle launch switches.  usage: python tools/capture_probe.py <variant>"""
import importlib, os, sys, traceback
import torch
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
ops = pkg.ops
dev = torch.device("cuda:0")
variant = sys.argv[1]
if variant == "F":
    ops.set_pdl = lambda on: True   # the capture keeps programmatic launches
shape = (2, 5, 64, 64, 64) if variant == "G" else (1, 5, 32, 32, 48)
torch.manual_seed(0)
model = pkg.UNet3D(5, 1, init_features=64).to(dev)
x = torch.randn(*shape, device=dev)
y = (torch.rand(shape[0], 1, *shape[2:], device=dev) < 0.2).float()
if variant in ("A", "B", "E"):
    model.eval()
    with torch.no_grad():
        a = model(x)
if variant == "C":
    other = pkg.UNet3D(5, 1, init_features=64).to(dev).eval()
    with torch.no_grad():
        a = other(x)
    del other
if variant == "E":
    del a
model.train()
if variant == "H":
    model.engine.overlap_wgrad = False
if variant == "I":
    ops.set_pdl(False)
opt = pkg.FusedAdam(model, lr=1e-4)
stepper = pkg.GraphedTrainStep(model, pkg.BCEDiceLoss(), opt)
orig = stepper._capture
def cap(key, xx, yy):
    try:
        orig(key, xx, yy)
    except Exception:
        print('capture raised', type(sys.exc_info()[1]).__name__)
        raise
stepper._capture = cap
for i in range(4):
    if variant == "B" and i == 2:
        torch.cuda.synchronize()
    if variant == "K":
        stepper(x, y)
    elif variant == "L":
        loss = stepper(x, y).detach()
    else:
        loss = stepper(x, y)
torch.cuda.synchronize()
print(variant, "replays", stepper.replays, "disabled", (stepper.disabled or "")[:80].replace("\n", " "), flush=True)
