"""CPU tests of the host-side logic: data contract, fold splits, window schedule, gradient-bucket planning, and the
data-parallel gradient exchange run for real over a world-size-2 gloo group (no GPU compute)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_pkg

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unet3d_oracle as oracle  # noqa: E402


def test_synthetic_batch_contract_and_zero_fill(pkg):
    loader = pkg.data.get_dataloader(batch_size=2, target_size=(16, 16, 16), n_cases=6, missing_strategy="zero_fill",
                                     is_training=False)
    batch = next(iter(loader))
    assert batch["image"].shape == (2, 5, 16, 16, 16) and batch["image"].dtype == torch.float32
    assert batch["label"].shape == (2, 1, 16, 16, 16)
    assert set(batch["label"].unique().tolist()) <= {0.0, 1.0}
    assert len(batch["case_id"]) == 2
    ds = pkg.data.SyntheticProstateDataset(40, (8, 8, 8), "zero_fill", missing_prob=0.5)
    zero_channels = 0
    for i in range(len(ds)):
        img = ds[i]["image"]
        assert img[0].abs().sum() > 0  # ADC always present
        zero_channels += sum(int(img[m].abs().sum() == 0) for m in range(1, 5))
    assert zero_channels > 20  # whole channels are zero, not partially masked
    skip = pkg.data.SyntheticProstateDataset(40, (8, 8, 8), "skip", missing_prob=0.5)
    assert 0 < len(skip) < 40
    dup = pkg.data.SyntheticProstateDataset(40, (8, 8, 8), "duplicate", missing_prob=0.5)
    assert all(dup[i]["image"].abs().sum((1, 2, 3)).min() > 0 for i in range(10))
    with pytest.raises(ValueError):
        pkg.data.SyntheticProstateDataset(4, (8, 8, 8), "interpolate")


def test_kfold_splits_are_json_safe_partitions(pkg):
    import json
    splits = pkg.data.get_kfold_splits(10, 5)
    assert len(splits) == 5
    seen = []
    for tr, va in splits:
        assert sorted(tr + va) == list(range(10)) and len(va) == 2
        seen += va
    assert sorted(seen) == list(range(10))
    json.dumps(splits)


def test_window_schedule_matches_oracle(pkg):
    par = pkg.parallel
    for extent, window, stride in [(256, 128, 64), (64, 64, 64), (100, 64, 64), (40, 64, 64), (130, 128, 64)]:
        assert par.window_origins(extent, window, stride) == oracle.window_origins(extent, window, stride)
    sched, win = par.window_schedule((8, 5, 256, 256, 64), (128, 128, 64), (64, 64, 64))
    assert win == (128, 128, 64) and len(sched) == 8 * 9   # BASELINE cfg #4: 9 windows per volume
    per_rank = [len(sched[r::8]) for r in range(8)]
    assert per_rank == [9] * 8


def test_bucket_plan_and_batch_sharding(pkg):
    par = pkg.parallel
    ends = [10, 30, 35, 100, 180, 181, 400]
    buckets = par.plan_buckets(400, ends, 50)
    assert buckets[0][0] == 0 and buckets[-1][1] == 400
    assert all(a[1] == b[0] for a, b in zip(buckets, buckets[1:]))
    assert all(hi in ends for _, hi in buckets)
    assert all(hi - lo >= 50 for lo, hi in buckets[:-1])
    assert [par.shard_batch(16, r, 8) for r in range(8)] == [slice(2 * r, 2 * r + 2) for r in range(8)]
    assert [par.shard_batch(5, r, 2) for r in range(2)] == [slice(0, 3), slice(3, 5)]
    # the real model's plan: ~25 MB buckets over the 361 MB flat gradient
    m = pkg.UNet3D(5, 1)
    eng = m.engine
    slots, off = {}, 0
    for _, p in eng.ordered_params():
        slots[id(p)] = (off, p.numel())
        off += (p.numel() + 63) // 64 * 64
    plan = par.plan_buckets(off, sorted({o + n for o, n in slots.values()}), int(25 * 2 ** 20 / 4))
    assert 8 <= len(plan) <= 16 and plan[-1][1] == off


class _FakeEngine:
    def __init__(self, rank):
        n = 1000
        self.flat_grad = torch.arange(n, dtype=torch.float32) * (rank + 1)
        self._slots = {i: (i * 100, 100) for i in range(10)}


def _dp_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_pkg()
    eng = _FakeEngine(rank)
    sync = pkg.parallel.GradSync(eng, bucket_mb=250 * 4 / 2 ** 20)   # 250-element buckets
    for hi in (100, 200, 300, 600, 1000):     # backward reporting completed prefixes, as Engine.backward does
        sync.ready(hi)
    sync.finish()
    expect = torch.arange(1000, dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok = torch.equal(eng.flat_grad, expect) and sync.launched == 4
    # a second step reuses the plan
    eng.flat_grad = torch.ones(1000) * (rank + 1)
    sync.ready(1000)
    sync.finish()
    ok = ok and torch.equal(eng.flat_grad, torch.full((1000,), float(sum(r + 1 for r in range(world)))))
    out[rank] = ok
    dist.destroy_process_group()


def test_gradient_sync_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(world, 29611, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_oracle_dp_semantics_small():
    """two identical shards: the averaged gradient equals the single-shard gradient (rank-local BN and loss)"""
    torch.manual_seed(0)
    pkg = load_pkg()
    m = pkg.UNet3D(5, 1, init_features=16)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 16, 16, 16, generator=g)
    y = (torch.rand(2, 1, 16, 16, 16, generator=g) > 0.8).float()
    sd_a = {k: v.clone() for k, v in sd.items()}
    sd_b = {k: v.clone() for k, v in sd.items()}
    l1, g1, _ = oracle.train_step(sd_a, {}, x, y)
    l2, g2 = oracle.dp_train_step(sd_b, {}, [x, x], [y, y])
    assert abs(l1.item() - l2.item()) < 1e-6
    for k in g1:
        assert torch.allclose(g1[k], g2[k], rtol=1e-4, atol=1e-7), k
    for k in sd_a:
        assert torch.allclose(sd_a[k].float(), sd_b[k].float(), rtol=1e-4, atol=1e-6), k


def test_physical_weight_layout_views(pkg):
    """conv weights live as [27][Cout][Cin] (packed tap order kd, kw, kh); Parameter.data is the permuted torch-shaped
    view, so state_dict / load_state_dict / optimizers are unaffected"""
    import importlib
    eng = importlib.import_module(pkg.__name__ + ".engine")
    w = torch.randn(32, 16, 3, 3, 3)
    p = torch.nn.Parameter(w.clone())
    flat = torch.zeros(8 + w.numel())
    view = eng._slot_view(flat, 8, p)
    assert view.shape == w.shape and view.stride() == eng._phys_strides(32, 16)
    view.copy_(w)
    phys = flat[8:].view(27, 32, 16)
    native = [(t // 9) * 9 + (t % 3) * 3 + (t // 3) % 3 for t in range(27)]
    assert torch.equal(phys, w.reshape(32, 16, 27).permute(2, 0, 1)[native])
    assert torch.equal(view, w) and torch.equal(view.contiguous(), w)
    # non-conv3 parameters keep their natural layout; the 5-modality first conv is not a packed conv
    b = torch.nn.Parameter(torch.randn(7))
    assert eng._slot_view(flat, 0, b).is_contiguous()
    assert not eng._is_conv3(torch.nn.Parameter(torch.randn(64, 5, 3, 3, 3)))
    assert not eng._is_conv3(torch.nn.Parameter(torch.randn(1, 64, 1, 1, 1)))
    # a state_dict made of such views round-trips through torch.save / load_state_dict
    import io
    buf = io.BytesIO()
    torch.save({"w": view}, buf)
    buf.seek(0)
    assert torch.equal(torch.load(buf)["w"], w)


def test_sharded_loader_and_window_ownership(pkg):
    par = pkg.parallel
    batch = {"image": torch.arange(5).view(5, 1).float(), "label": torch.zeros(5, 1), "case_id": list("abcde"),
             "meta": "kept"}
    parts = [par.shard_dict_batch(batch, r, 2) for r in range(2)]
    assert parts[0]["case_id"] == ["a", "b", "c"] and parts[1]["case_id"] == ["d", "e"]
    assert torch.equal(torch.cat([p["image"] for p in parts]), batch["image"]) and parts[0]["meta"] == "kept"
    assert par.shard_dict_batch(batch, 0, 1) is batch
    assert par.shard_dict_batch({"image": torch.zeros(1, 1), "label": torch.zeros(1, 1)}, 0, 2) is None  # fewer samples than ranks
    loader = [batch, {"image": torch.zeros(1, 1), "label": torch.zeros(1, 1), "case_id": ["z"]}]
    assert [len(b["case_id"]) for b in par.ShardedLoader(loader, 1, 2)] == [2]   # the short batch is skipped on every rank
    # data-parallel ranks iterate identical global batches: the shuffle order comes from a seeded generator
    a = [b["case_id"] for b in pkg.data.get_dataloader(batch_size=2, target_size=(8, 8, 8), n_cases=6)]
    b = [b["case_id"] for b in pkg.data.get_dataloader(batch_size=2, target_size=(8, 8, 8), n_cases=6)]
    assert a == b
    # window cover = per-axis counts whose product is the number of windows over a voxel
    shape, window, stride = (2, 5, 48, 40, 32), (32, 32, 32), (16, 16, 16)
    cov = par.window_cover(shape, window, stride)
    sched, win = par.window_schedule(shape, window, stride)
    cnt = torch.zeros(48, 40, 32)
    for v, d0, h0, w0 in sched:
        if v == 0:
            cnt[d0:d0 + win[0], h0:h0 + win[1], w0:w0 + win[2]] += 1
    prod = cov[:48].view(-1, 1, 1) * cov[48:88].view(1, -1, 1) * cov[88:].view(1, 1, -1)
    assert torch.equal(prod.float(), cnt)
    # BASELINE configs[3]: 8 volumes on 8 ranks -> one whole volume per rank, nothing to exchange
    own = par.volume_owners((8, 5, 256, 256, 64), (128, 128, 64), (64, 64, 64), 8)
    assert own == {v: (v, [v]) for v in range(8)}
    assert par.owned_volumes((8, 5, 256, 256, 64), (128, 128, 64), (64, 64, 64), 3, 8) == [3]
    # a single volume on 2 ranks is split and owned by rank 0
    assert par.volume_owners((1, 5, 256, 256, 64), (128, 128, 64), (64, 64, 64), 2) == {0: (0, [0, 1])}
    assert par.dist_info() == (0, 1) and par.all_mean(2.5, "cpu") == 2.5


def test_load_multimodal_images_case_directory(pkg, tmp_path):
    """script/predict.py:8-82 on a case directory of .npy volumes: order, min-max, zero_fill / duplicate / skip"""
    import numpy as np
    rng = np.random.default_rng(0)
    vols = {}
    for m in pkg.predict.PREDICT_MODALITIES:
        os.makedirs(tmp_path / m)
        if m != "T2 fs":
            vols[m] = (rng.normal(size=(4, 5, 6)) * 3 + 1).astype(np.float32)
            np.save(tmp_path / m / "vol.npy", vols[m])
    img, names = pkg.load_multimodal_images(str(tmp_path))
    assert names == ["ADC", "DWI", "gaoqing-T2", "T2 fs", "T2 not fs"] and img.shape == (5, 4, 5, 6)
    assert img.dtype == np.float32 and img[3].max() == 0.0
    for i, m in enumerate(names):
        if m in vols:
            v = vols[m]
            assert np.allclose(img[i], oracle.minmax_normalize(v[None])[0], atol=1e-6)
    dup, _ = pkg.load_multimodal_images(str(tmp_path), "duplicate")
    assert np.array_equal(dup[3], dup[0])
    with pytest.raises(FileNotFoundError):
        pkg.load_multimodal_images(str(tmp_path), "skip")
    os.rmdir(tmp_path / "T2 fs")
    with pytest.raises(FileNotFoundError):
        pkg.load_multimodal_images(str(tmp_path))


def test_build_is_content_addressed(pkg):
    from importlib import import_module
    b = import_module(pkg.__name__ + ".build")
    assert not b.is_stale() and b.build() == b.LIB_PATH          # built by the fixture's load; a no-op now
    assert open(b.LIB_PATH + ".hash").read().strip() == b.source_hash()
    assert b.source_hash() != b.source_hash(dev=True)
