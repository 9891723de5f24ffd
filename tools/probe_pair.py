"""Dev tool: functional check and cycle count of the CTA-pair (cta_group::2) MMA primitives.  python tools/probe_pair.py"""
import importlib, os, sys
os.environ["B200_DEV"] = "1"   # the probes live in the development library only (build.build(dev=True))
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("prostate-cancer-multimodal-segmentation_b200")
lib = pkg.load_library()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for n in (64, 128, 192, 256):
    for k in (64, 256):
        a = torch.randn(256, k, generator=g).to(torch.bfloat16).to(dev)
        b = torch.randn(n, k, generator=g).to(torch.bfloat16).to(dev)
        d = torch.zeros(1, 256, n, device=dev)
        cyc = torch.zeros(1, dtype=torch.int64, device=dev)
        rc = lib.b200_probe_pair(a.data_ptr(), b.data_ptr(), n, k, 1, d.data_ptr(), cyc.data_ptr(), 1,
                                 torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = a.float() @ b.float().t()
        err = ((d[0] - ref).norm() / ref.norm()).item()
        print(f"N {n:3d} K {k:3d}: rc={rc} rel-L2 vs torch {err:.2e}  rows0-127 {((d[0,:128]-ref[:128]).norm()/ref[:128].norm()).item():.1e}"
              f" rows128-255 {((d[0,128:]-ref[128:]).norm()/ref[128:].norm()).item():.1e}")
print("--- cycles per M=256 MMA (74 pairs, K = 256 resident, 500 repeats)")
for n in (64, 128, 192, 256):
    k, iters, pairs = 256, 500, 74
    a = torch.randn(256, k, generator=g).to(torch.bfloat16).to(dev)
    b = torch.randn(n, k, generator=g).to(torch.bfloat16).to(dev)
    d = torch.zeros(pairs, 256, n, device=dev)
    cyc = torch.zeros(pairs, dtype=torch.int64, device=dev)
    rc = lib.b200_probe_pair(a.data_ptr(), b.data_ptr(), n, k, iters, d.data_ptr(), cyc.data_ptr(), pairs,
                             torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    per = cyc.float().mean().item() / (iters * 16)
    print(f"N {n:3d}: rc={rc} {per:7.1f} cycles per 256xNx16 MMA (math floor {n / 2:5.1f} per SM) -> "
          f"{100 * (n / 2) / per:5.1f}% of the pair's tensor peak")
