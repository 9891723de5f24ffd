"""GPU parity tests of the "next" rows: resample-to-target, per-modality normalisation, hard Dice / IoU
counts on the device, and the host-side staging helpers (H2D prefetcher, one-step-late scalar read-back).
Oracle: oracle/unet3d_oracle.py (resample3d_itk is a restatement of ITK's published algorithm — parity unpinned,
SimpleITK is not installable offline)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unet3d_oracle as oracle  # noqa: E402

pytestmark = pytest.mark.gpu

RESAMPLE_CASES = [
    # in (D,H,W) -> out (D,H,W)
    ((6, 5, 7), (6, 5, 7)),          # identity
    ((8, 8, 8), (16, 16, 16)),       # 2x up: last output index per axis falls outside the buffer -> 0
    ((20, 24, 18), (16, 16, 16)),    # down, non-integer ratios
    ((5, 9, 13), (12, 7, 13)),       # mixed up / down / same, odd sizes
    ((3, 3, 3), (16, 16, 16)),       # > 2x up: several trailing indices outside
    ((1, 4, 4), (2, 2, 2)),          # single slice
]


@pytest.mark.parametrize("case", RESAMPLE_CASES)
@pytest.mark.parametrize("nearest", [False, True])
def test_resample3d_matches_itk_restatement(ops, cuda_dev, case, nearest):
    src, dst = case
    g = torch.Generator().manual_seed(hash((src, dst)) % 1000)
    x = torch.randn(2, 3, *src, generator=g)
    if nearest:
        x = (x > 0.5).float() * torch.randint(0, 3, x.shape, generator=g).float()   # label values 0, 1, 2
    out = ops.resample3d(x.to(cuda_dev), dst, nearest=nearest, binarize=nearest).cpu().numpy()
    ref = oracle.resample3d_itk(x.numpy(), dst, nearest=nearest, binarize=nearest)
    assert out.shape == ref.shape == (2, 3) + dst
    if nearest:
        assert np.array_equal(out, ref)          # index arithmetic: bit-exact
    else:
        # same fp32 lerp order as the restatement; allow one rounding of the fused multiply-adds
        assert np.allclose(out, ref, rtol=1e-5, atol=1e-6)
    if src == dst:
        assert np.array_equal(out, (x.numpy() > 0).astype(np.float32) if nearest else x.numpy())


def test_resample_case_contract(pkg, cuda_dev):
    g = torch.Generator().manual_seed(3)
    img = torch.rand(5, 20, 24, 18, generator=g).to(cuda_dev)
    lab = (torch.rand(1, 10, 12, 9, generator=g) > 0.7).float().to(cuda_dev) * 2.0
    i2, l2 = pkg.data.resample_case(img, lab, (16, 16, 16))
    assert i2.shape == (5, 16, 16, 16) and l2.shape == (1, 16, 16, 16)
    assert set(torch.unique(l2).tolist()) <= {0.0, 1.0}
    i3, l3 = pkg.data.resample_case(i2, l2, (16, 16, 16))
    assert i3 is i2 and torch.equal(l3, l2)


def test_minmax_normalize_matches_reference_formula(pkg, ops, cuda_dev):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(5, 9, 17, 33, generator=g) * 37.0 + 11.0
    x[3] = 4.25                                   # constant modality -> zeros
    x[4] = -x[4].abs()                            # all negative
    ref = oracle.minmax_normalize(x.numpy())
    out = pkg.predict.normalize_modalities_(x.to(cuda_dev).clone()).cpu().numpy()
    assert np.array_equal(out[3], np.zeros_like(out[3]))
    assert np.allclose(out, ref, rtol=1e-6, atol=1e-7)
    assert out.min() == 0.0 and out.max() == 1.0
    host = pkg.predict.normalize_modalities(x.numpy())
    assert np.allclose(out, host, rtol=1e-6, atol=1e-7)


def test_seg_counts_exact_and_metrics(pkg, ops, cuda_dev):
    g = torch.Generator().manual_seed(9)
    n, shape = 3, (1, 20, 31, 17)
    score = torch.rand(n, *shape, generator=g)
    label = (torch.rand(n, *shape, generator=g) > 0.6).float()
    score[1] = 0.0                                # empty prediction
    label[2] = 0.0                                # empty target
    counts = ops.seg_counts(score.to(cuda_dev), label.to(cuda_dev), 0.5).cpu()
    p, t = (score > 0.5), (label > 0.5)
    ref = torch.stack([(p & t).flatten(1).sum(1), p.flatten(1).sum(1), t.flatten(1).sum(1)], 1)
    assert torch.equal(counts, ref)
    for i in range(n):
        d_ref, j_ref = oracle.hard_dice_iou(p[i].float().numpy(), t[i].float().numpy())
        assert abs(pkg.validate.calculate_dice_score(p[i].float().to(cuda_dev), t[i].to(cuda_dev)) - d_ref) < 1e-12
        assert abs(pkg.validate.calculate_iou(p[i].float().to(cuda_dev), t[i].to(cuda_dev)) - j_ref) < 1e-12
    with pytest.raises(ValueError):
        ops.seg_counts(score.to(cuda_dev), label[:2].to(cuda_dev))


def test_seg_counts_large_exceeds_float24(ops, cuda_dev):
    # 2^25 positive voxels: a float32 sum of ones would stall at 2^24, the int64 counts do not
    n = 1 << 25
    ones = torch.ones(1, n, device=cuda_dev)
    counts = ops.seg_counts(ones, ones, 0.5)
    assert counts.tolist() == [[n, n, n]]


def test_device_prefetcher_and_async_scalars(pkg, cuda_dev):
    g = torch.Generator().manual_seed(11)
    batches = [{"image": torch.randn(2, 5, 4, 4, 4, generator=g), "label": torch.rand(2, 1, 4, 4, 4, generator=g),
                "case_id": [f"c{i}a", f"c{i}b"]} for i in range(5)]
    batches[2]["image"] = batches[2]["image"].pin_memory()
    reader = pkg.data.AsyncScalarReader()
    seen = []
    for i, b in enumerate(pkg.data.DevicePrefetcher(batches, cuda_dev)):
        assert b["image"].is_cuda and b["label"].is_cuda and b["case_id"] == batches[i]["case_id"]
        seen.append((b["image"].cpu().clone(), b["label"].cpu().clone()))
        reader.push((b["image"].sum() + b["label"].sum()))
    vals = reader.finish()
    assert len(seen) == len(vals) == 5
    for i, (im, lb) in enumerate(seen):
        assert torch.equal(im, batches[i]["image"]) and torch.equal(lb, batches[i]["label"])
        assert abs(vals[i] - (batches[i]["image"].sum() + batches[i]["label"].sum()).item()) < 1e-3
    assert list(pkg.data.DevicePrefetcher([], cuda_dev)) == []
    with pytest.raises(ValueError):
        pkg.data.DevicePrefetcher(batches, "cpu")


def test_validate_loop_on_device(pkg, cuda_dev):
    torch.manual_seed(0)
    model = pkg.UNet3D(5, 1, init_features=16).to(cuda_dev)
    loader = pkg.data.get_dataloader(batch_size=2, target_size=(16, 16, 16), n_cases=4, is_training=False)
    rows = pkg.validate.validate(model, loader, cuda_dev)
    assert [r["case_id"] for r in rows] == [f"case_{i:04d}" for i in range(4)]
    # independent recomputation with torch on the host
    model.eval()
    k = 0
    for batch in loader:
        mask = (model.predict(batch["image"].to(cuda_dev)) > 0.5).float().cpu()
        for i in range(mask.shape[0]):
            d_ref, j_ref = oracle.hard_dice_iou(mask[i].numpy(), batch["label"][i].numpy())
            assert abs(rows[k]["dice"] - d_ref) < 1e-9 and abs(rows[k]["iou"] - j_ref) < 1e-9
            k += 1
