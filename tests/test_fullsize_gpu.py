"""Parity at the real BASELINE.json shapes (configs[1] 2x5x128^3, configs[3] one 128x128x64 window, configs[4] 5x160^3
zero_fill at base 32 and 64) plus an odd-extent / two-class case, on the GPU against the oracle run on the same GPU.

Two kinds of evidence (tests/parity_util.py):

1. test_layerwise_*: every op of one training step — 18 convolutions, 18 BatchNorm+ReLU, 4 poolings, 4 transposed
   convolutions, the 1x1x1 head and the loss, forward AND backward — is checked in isolation on the engine's own
   inputs against the torch fp32 op the reference calls there.  This is where north_star's per-layer bound (2e-2
   relative L2) applies; measured: 1.7e-3 for every bf16 output (= the bf16 rounding of the output itself), <= 4e-4 for
   every fp32 output (weight gradients), poolings bit-exact.  The assertion is 4e-3, five times tighter than required.

2. test_end_to_end_*: the whole step against the oracle's fp32 form (logits <= 2e-2, loss <= 1e-3: north_star) and,
   per layer and per parameter, against its bf16-storage form.  End to end a 23-layer BatchNorm'd network with
   bf16-rounded tensors amplifies differences of ONE fp32 rounding to tens of percent in the encoder gradients (flipped
   bf16 roundings flip ReLU masks downstream): the storage form moves that far from itself when its input and weights
   are perturbed by 2^-21 relative (column `storage~_vs_storage`), and it and torch's own bf16 autocast sit that far
   from fp32.  So the end-to-end bound is: the engine is as close to the storage form as the storage form is to its
   perturbed self (x1.5: the yardstick is one random draw), and as close to fp32 as the storage form is (x1.5: the yardstick is one random draw), with the 2e-2 floor.
"""
import os

import pytest
import torch

import parity_util as pu
from conftest import ROOT

pytestmark = pytest.mark.gpu
OUT = os.path.join(ROOT, "gpurun_out")
TOL_SPEC = 2e-2     # north_star: per layer, relative L2
TOL_LAYER = 4e-3    # what is asserted for isolated layers (measured 1.7e-3 / 4e-4)
SLACK = 1.5         # on the bf16-storage yardsticks (each is a single random draw)

CASES = {
    "cfg1_2x128": dict(batch=2, size=(128, 128, 128), base=64),
    "cfg3_window": dict(batch=1, size=(128, 128, 64), base=64),
    "cfg4_160_b32": dict(batch=1, size=(160, 160, 160), base=32, zero_fill=True),
    "cfg4_160_b64": dict(batch=1, size=(160, 160, 160), base=64, zero_fill=True),
    "odd_2class": dict(batch=1, size=(20, 36, 18), base=64, n_classes=2),
}


@pytest.mark.parametrize("case", list(CASES))
def test_layerwise_identical_inputs(pkg, cuda_dev, case):
    rows = pu.layerwise_check(pkg, cuda_dev, **CASES[case])
    pu.write_layerwise(rows, os.path.join(OUT, f"layerwise_{case}.txt"), header=f"{case}: {CASES[case]}")
    assert len(rows) > 150
    bad = []
    for op, what, v in rows:
        if what.endswith("mismatches"):
            if v != 0:
                bad.append((op, what, v))
        elif op == "loss":
            if v > 1e-5:
                bad.append((op, what, v))
        elif "(degenerate" in what or what.endswith("(info)"):
            continue   # BatchNorm backward over < 16 samples per channel (bottom level of the odd-extent case)
        elif not (v <= TOL_LAYER):
            bad.append((op, what, v))
    assert not bad, f"{case}: ops beyond {TOL_LAYER} (spec {TOL_SPEC}) on identical inputs: {bad}"
    torch.cuda.empty_cache()


@pytest.mark.parametrize("case", ["cfg1_2x128", "cfg3_window", "cfg4_160_b32", "cfg4_160_b64"])
def test_end_to_end_step(pkg, cuda_dev, case):
    res = pu.train_step_parity(pkg, cuda_dev, **CASES[case])
    pu.write_report(res, os.path.join(OUT, f"parity_{case}.txt"))
    assert abs(res["loss"]["ours"] - res["loss"]["fp32"]) < 1e-3
    assert res["logits"]["ours_vs_fp32"] < TOL_SPEC
    assert res["logits"]["ours_vs_storage"] < max(TOL_SPEC, SLACK * res["logits"]["storage~_vs_storage"])
    assert res["mask_mismatch_fp32_sure"] <= 1e-4 * res["config"]["batch"] * torch.tensor(res["config"]["size"]).prod().item()
    assert max(res["bn_buffers"].values()) < 1e-2
    assert res["dead_bias_ratio_max"] < 5e-2
    bad = {}
    for name, e in res["acts"]["storage"].items():
        lim = max(TOL_SPEC, SLACK * res["acts"]["storage~"][name])
        if not e <= lim:
            bad["act " + name] = (e, lim)
    for name, row in res["grads"].items():
        lim_s = max(TOL_SPEC, SLACK * row["storage~_vs_storage"])
        lim_f = max(TOL_SPEC, SLACK * row["storage_vs_fp32"])
        if not row["ours_vs_storage"] <= lim_s:
            bad["grad/storage " + name] = (row["ours_vs_storage"], lim_s)
        if not row["ours_vs_fp32"] <= lim_f:
            bad["grad/fp32 " + name] = (row["ours_vs_fp32"], lim_f)
    assert not bad, f"{case}: beyond the bf16-storage yardstick: {bad}"
    # the layers next to the loss cross few bf16 tensors: there the plain 2e-2 holds end to end
    for name in ("outc.weight", "outc.bias", "up4.conv.conv.4.weight", "up4.conv.conv.4.bias", "up4.conv.conv.3.weight"):
        assert res["grads"][name]["ours_vs_fp32"] < TOL_SPEC, (name, res["grads"][name])
    torch.cuda.empty_cache()
