"""SASS opcode histogram of the product library (run in the build container: needs cuobjdump, no GPU).
usage: python tools/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "prostate-cancer-multimodal-segmentation_b200", "libb200unet3d.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "RED", "ATOMG")
print("# cuobjdump -sass libb200unet3d.so (sm_100a): instructions per kernel; tensor-core / TMEM / TMA opcodes by prefix")
print("# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store,")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA would be the legacy mma.sync path (absent)")
print(f"{'kernel':46s} {'instrs':>7s} " + " ".join(f"{k:>12s}" for k in KEY))
for k, c in hist.items():
    tot = sum(c.values())
    cols = []
    for key in KEY:
        if key == "UTCHMMA":
            n = sum(v for op, v in c.items() if op.startswith("UTCHMMA") and ".2CTA" not in op)
        elif key == "UTCHMMA.2CTA":
            n = sum(v for op, v in c.items() if op.startswith("UTCHMMA") and ".2CTA" in op)
        else:
            n = sum(v for op, v in c.items() if op.startswith(key))
        cols.append(n)
    name = k if len(k) <= 46 else k[:43] + "..."
    print(f"{name:46s} {tot:7d} " + " ".join(f"{n:12d}" for n in cols))
