"""Data-parallel training over NCCL: N ranks must reproduce the single-process emulation of the same N shards — same
replicated weights, rank-local BatchNorm statistics and loss, mean of the per-shard gradients, one fused Adam step
(oracle/unet3d_oracle.py:dp_train_step states the same semantics, checked in tests/test_dp_onegpu.py).
With two GPUs the group has two ranks; on a one-GPU box the same code paths run as a ONE-rank NCCL group (bucketed
ncclAllReduce launches issued during backward, their capture into the CUDA graph, the teardown order) — NCCL cannot put
two ranks on one device; the two-rank semantics on one GPU are covered over gloo in tests/test_dp_onegpu.py."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_pkg

pytestmark = pytest.mark.gpu


def _world():
    return 2 if torch.cuda.device_count() >= 2 else 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pkg = load_pkg()
    par = pkg.parallel
    torch.manual_seed(100 + rank)                       # different init per rank: the broadcast must fix it
    model = pkg.UNet3D(5, 1, init_features=16).to(dev).train()
    opt = pkg.FusedAdam(model, lr=1e-3, weight_decay=1e-5)
    sync = par.make_data_parallel(model, opt, bucket_mb=0.25)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2 * world, 5, 32, 32, 32, generator=g)
    y = (torch.rand(2 * world, 1, 32, 32, 32, generator=g) > 0.85).float()
    sl = par.shard_batch(x.shape[0], rank, world)
    crit = pkg.BCEDiceLoss()
    w0 = model.engine.flat_param.clone()

    opt.zero_grad()
    loss = crit(model(x[sl].to(dev)), y[sl].to(dev))
    loss.backward()
    grads = model.engine.flat_grad.clone() / world      # the all-reduce leaves the SUM; Adam folds 1/world
    opt.step()
    torch.cuda.synchronize()
    w1 = model.engine.flat_param.clone()

    ok = True
    # every rank holds the same weights before and after
    for t in (w0, w1):
        ref = t.clone()
        dist.broadcast(ref, 0)
        ok = ok and torch.equal(ref, t)
    ok = ok and sync.launched >= 2                       # several buckets, issued during backward
    # single-process emulation on this rank: all shards in turn from the same initial weights
    emu = pkg.UNet3D(5, 1, init_features=16).to(dev).train()
    emu.engine.prepare(dev)
    emu.engine.flat_param.copy_(w0)
    emu.engine.external_epoch += 1
    emu_opt = pkg.FusedAdam(emu, lr=1e-3, weight_decay=1e-5)
    emu_opt.grad_scale = 1.0 / world
    emu_opt.zero_grad()
    for r in range(world):
        s = par.shard_batch(x.shape[0], r, world)
        crit(emu(x[s].to(dev)), y[s].to(dev)).backward()   # accumulates into the flat gradient
    emu_g = emu.engine.flat_grad.clone() / world
    emu_opt.step()
    torch.cuda.synchronize()
    rel = ((grads - emu_g).norm() / emu_g.norm()).item()
    relw = ((w1 - emu.engine.flat_param).norm() / (w1 - w0).norm()).item()
    out[rank] = (bool(ok), rel, relw)
    dist.destroy_process_group()


def test_nccl_dp_matches_emulation():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    world = _world()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29633, out), nprocs=world, join=True)
    for r in range(world):
        ok, rel, relw = out[r]
        assert ok, f"rank {r}: replicas diverged or too few buckets"
        assert rel < 1e-3, f"rank {r}: all-reduced gradient vs emulation rel-L2 {rel}"
        assert relw < 5e-2, f"rank {r}: Adam update vs emulation rel-L2 {relw}"


def _worker_graph(rank, world, port, out):
    """data-parallel ranks replaying the step from a CUDA graph that also holds the NCCL bucket all-reduces"""
    import gc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pkg = load_pkg()
    par = pkg.parallel
    torch.manual_seed(100 + rank)
    model = pkg.UNet3D(5, 1, init_features=16).to(dev).train()
    opt = pkg.FusedAdam(model, lr=1e-3, weight_decay=1e-5)
    par.make_data_parallel(model, opt, bucket_mb=0.25)
    stepper = pkg.GraphedTrainStep(model, pkg.BCEDiceLoss(), opt, capture_collectives=True)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2 * world, 5, 32, 32, 32, generator=g)
    y = (torch.rand(2 * world, 1, 32, 32, 32, generator=g) > 0.85).float()
    sl = par.shard_batch(x.shape[0], rank, world)        # every rank trains on its own shard
    xs, ys = x[sl].to(dev), y[sl].to(dev)
    ok, losses = True, []
    for i in range(5):
        losses.append(stepper(xs, ys).item())
        torch.cuda.synchronize()
        for t in (model.engine.flat_param, model.engine.flat_grad):   # identical on every rank only if the recorded
            ref = t.clone()                                            # all-reduces really ran in the replay
            dist.broadcast(ref, 0)
            ok = ok and torch.equal(ref, t)
    out[rank] = (bool(ok), stepper.replays, stepper.disabled, losses)
    del stepper
    gc.collect()
    torch.cuda.synchronize()
    dist.destroy_process_group()


def test_nccl_dp_graph_replay_keeps_replicas_identical():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    world = _world()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_graph, args=(world, 29634, out), nprocs=world, join=True)
    for r in range(world):
        ok, replays, disabled, losses = out[r]
        assert disabled is None and replays == 3, (replays, disabled)
        assert ok, f"rank {r}: replicas or all-reduced gradients differ between ranks after a replayed step"
        assert losses[-1] < losses[0]

