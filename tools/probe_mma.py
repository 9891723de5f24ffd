"""Dev tool: tcgen05.mma cycles vs N (SS mode, operands in shared memory).  python tools/probe_mma.py"""
import importlib, os, sys
os.environ["B200_DEV"] = "1"   # the probes live in the development library only (build.build(dev=True))
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
lib = pkg.load_library()
dev = torch.device("cuda:0")
for blocks in (1, 148):
    for n in (32, 64, 96, 128, 192, 256):
        for stages in (1, 4):
            if 1024 + stages * (16384 + n * 128) + 64 > 227 * 1024:
                continue
            out = torch.zeros(blocks, dtype=torch.int64, device=dev)
            iters = 2000
            rc = lib.b200_probe_mma(n, iters, stages, out.data_ptr(), blocks, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            cyc = out.float().mean().item() / (iters * 4)
            ideal = 128 * n / 256
            print(f"blocks {blocks:3d} N {n:3d} stages {stages}: {cyc:7.1f} cycles/MMA (math floor {ideal:5.1f}) "
                  f"-> {100 * ideal / cyc:5.1f}% of tensor peak, rc={rc}")

print("--- M and majorness sweep (148 CTAs)")
for m in (64, 128):
    for mn in (0, 1):
        for n in (64, 128, 192, 256):
            out = torch.zeros(148, dtype=torch.int64, device=dev)
            iters = 2000
            rc = lib.b200_probe_mma2(m, n, mn, iters, 2, out.data_ptr(), 148, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            cyc = out.float().mean().item() / (iters * 4)
            print(f"M {m:3d} N {n:3d} {'MN-major' if mn else 'K-major '}: {cyc:7.1f} cycles/MMA; useful-M flops/cycle "
                  f"{2 * m * n * 16 / cyc:8.0f} (peak 8192), rc={rc}")
