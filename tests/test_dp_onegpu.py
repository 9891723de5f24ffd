"""Multi-rank paths on ONE GPU: two processes share cuda:0 and exchange through a gloo group on CUDA tensors, so the
data-parallel step, the sharded sliding-window inference and `run.py train` under a process group are exercised (and
checked against the oracle) on a single-GPU box.  The ranks' kernels never wait on one another on the device (the
exchange is host-mediated), so sharing the GPU is safe.  tests/test_dp_gpu.py runs the same paths over NCCL when two
GPUs are present.

Oracle: oracle.dp_train_step — every rank's shard forward/backward from the same weights with rank-local
BatchNorm statistics and loss, mean of the parameter gradients, one Adam step, BatchNorm buffers of rank 0.  Bounds as
in tests/test_fullsize_gpu.py: per parameter max(2e-2, 1.5 x the bf16-storage oracle's own sensitivity to a one-fp32-
rounding perturbation — the larger of two draws)."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_pkg

pytestmark = pytest.mark.gpu
PORT = 29711


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK="0",
                      WORLD_SIZE=str(world), B200_DIST_BACKEND="gloo")
    torch.cuda.set_device(0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------------------ data-parallel step
def _dp_worker(rank, world, port, out):
    _init(rank, world, port)
    import unet3d_oracle as oracle
    import parity_util as pu
    pkg = load_pkg()
    par = pkg.parallel
    r, w, dev = par.init_distributed()
    assert (r, w) == (rank, world) and not par.backend_is_nccl()
    torch.manual_seed(100 + rank)                       # different init per rank: the broadcast must fix it
    model = pkg.UNet3D(5, 1, init_features=16).to(dev).train()
    opt = pkg.FusedAdam(model, lr=1e-3, weight_decay=1e-5)
    sync = par.make_data_parallel(model, opt, bucket_mb=0.25)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}   # rank 0's, after the broadcast
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2 * world, 5, 48, 48, 48, generator=g).to(dev)
    y = (torch.rand(2 * world, 1, 48, 48, 48, generator=g) > 0.85).float().to(dev)
    sl = par.shard_batch(x.shape[0], rank, world)
    crit = pkg.BCEDiceLoss()
    opt.zero_grad()
    loss = crit(model(x[sl]), y[sl])
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().clone() / world for n, p in model.named_parameters()}   # buffer holds the SUM
    opt.step()
    torch.cuda.synchronize()
    res = {"buckets": sync.launched}
    # replicas identical before and after
    same = True
    for t in (model.engine.flat_param, model.engine.flat_grad):
        ref = t.clone()
        dist.broadcast(ref, 0)
        same = same and bool(torch.equal(ref, t))
    res["replicas_identical"] = same
    mean_loss = par.all_mean(loss.item(), dev)
    if rank == 0:
        shards = [par.shard_batch(x.shape[0], q, world) for q in range(world)]
        xs, ys = [x[s] for s in shards], [y[s] for s in shards]

        def run(sd, store, perturb=None):
            sd = {k: v.clone() for k, v in sd.items()}
            xs_ = xs
            if perturb is not None:
                gen = torch.Generator(device=dev).manual_seed(perturb)
                names = oracle.param_names(sd)
                sd = {k: (pu._perturbed(v, gen) if k in names else v) for k, v in sd.items()}
                xs_ = [pu._perturbed(t, gen) for t in xs]
            st = {}
            l, gr = oracle.dp_train_step(sd, st, xs_, ys, lr=1e-3, weight_decay=1e-5, store=store)
            return l, gr, sd

        l32, g32, sd32 = run(sd0, None)
        ls, gs, sds = run(sd0, oracle.store_bf16)
        _, gp, _ = run(sd0, oracle.store_bf16, perturb=5)
        _, gp2, _ = run(sd0, oracle.store_bf16, perturb=6)
        res["loss_err"] = abs(mean_loss - l32.item())
        bad = {}
        for n, gg in grads.items():
            if pu.is_dead_bias(n):
                continue
            sens = max(_rel(gp[n], gs[n]), _rel(gp2[n], gs[n]))
            e_s, e_f = _rel(gg, gs[n]), _rel(gg, g32[n])
            if e_s > max(2e-2, 1.5 * sens) or e_f > max(2e-2, 1.5 * _rel(gs[n], g32[n])):
                bad[n] = (e_s, sens, e_f)
        res["bad_grads"] = bad
        res["near_loss"] = {n: _rel(grads[n], g32[n]) for n in ("outc.weight", "outc.bias", "up4.conv.conv.4.weight")}
        # Adam: the first step moves every weight by ~lr*sign(g); BatchNorm buffers follow rank 0
        msd = model.state_dict()
        agree = []
        for n, p in model.named_parameters():
            if pu.is_dead_bias(n):
                continue
            upd, ref = p.detach() - sd0[n], sd32[n] - sd0[n]
            agree.append((torch.sign(upd) == torch.sign(ref)).float().mean().item())
        res["adam_sign_agreement_min"] = min(agree)
        res["bn_buffers"] = max(_rel(msd[k], sd32[k]) for k in msd if "running" in k)
        res["nbt"] = int(msd["inc.conv.1.num_batches_tracked"].item())
    out[rank] = res
    par.shutdown_distributed()


def test_two_rank_dp_step_on_one_gpu_vs_oracle():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = mp.Manager().dict()
    mp.spawn(_dp_worker, args=(2, PORT, out), nprocs=2, join=True)
    for r in range(2):
        assert out[r]["replicas_identical"], f"rank {r}: replicas or all-reduced gradients differ"
        assert out[r]["buckets"] >= 2
    r0 = out[0]
    assert r0["loss_err"] < 1e-3
    assert not r0["bad_grads"], r0["bad_grads"]
    assert all(v < 2e-2 for v in r0["near_loss"].values()), r0["near_loss"]
    assert r0["adam_sign_agreement_min"] > 0.75
    assert r0["bn_buffers"] < 2e-2 and r0["nbt"] == 1


# ------------------------------------------------------------------------------------------------ sliding windows
def _window_worker(rank, world, port, out):
    _init(rank, world, port)
    import unet3d_oracle as oracle
    pkg = load_pkg()
    par = pkg.parallel
    _, _, dev = par.init_distributed()
    torch.manual_seed(1)
    model = pkg.UNet3D(5, 1, init_features=16).to(dev)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
    g = torch.Generator().manual_seed(3)
    x = torch.rand(3, 5, 48, 40, 32, generator=g).to(dev)
    window, stride = (32, 32, 32), (16, 16, 16)
    owners = par.volume_owners(tuple(x.shape), window, stride, world)
    mine = par.owned_volumes(tuple(x.shape), window, stride, rank, world)
    logits, mask = par.sliding_window_logits(model, x, window, stride, rank=rank, world=world, want=("logits", "mask"))
    full = par.sliding_window_logits(model, x, window, stride)            # single-process result on this rank
    res = {"owners": owners, "mine": mine}
    res["owned_match"] = all(torch.allclose(logits[v], full[v], rtol=1e-5, atol=1e-5) for v in mine)
    res["mask_match"] = all(torch.equal(mask[v], (torch.sigmoid(logits[v]) > 0.5).float()) for v in mine)
    everywhere = par.sliding_window_logits(model, x, window, stride, rank=rank, world=world, exchange="all")
    res["all_match"] = bool(torch.allclose(everywhere, full, rtol=1e-5, atol=1e-5))
    if rank == 0:
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        ref = oracle.sliding_window_logits(x, sd, window, stride)
        res["vs_oracle"] = _rel(full, ref)
    out[rank] = res
    par.shutdown_distributed()


def test_two_rank_sliding_window_on_one_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    out = mp.Manager().dict()
    mp.spawn(_window_worker, args=(2, PORT + 1, out), nprocs=2, join=True)
    owners = out[0]["owners"]
    assert any(len(rs) > 1 for _, rs in owners.values()), "the case must split a volume over both ranks"
    assert sorted(out[0]["mine"] + out[1]["mine"]) == [0, 1, 2]
    for r in range(2):
        assert out[r]["owned_match"] and out[r]["mask_match"] and out[r]["all_match"], (r, dict(out[r]))
    assert out[0]["vs_oracle"] < 2e-2


# ------------------------------------------------------------------------------------------------ run.py train
def _cli_worker(rank, world, port, save):
    _init(rank, world, port)
    pkg = load_pkg()
    cli = importlib.import_module(pkg.__name__ + ".cli")
    cli.main(["train", "--epochs", "1", "--batch_size", "2", "--learning_rate", "0.001", "--save_dir", save,
              "--init_features", "16", "--size", "32", "32", "32", "--n_cases", "2", "--seed", "11", "--loss", "bce_dice"])


def test_cli_train_two_ranks_checkpoint_vs_oracle_dp_step(pkg, cuda_dev, tmp_path):
    """`run.py train` under a 2-rank process group: one epoch = one global batch of 2 = one sample per rank; the
    checkpoint rank 0 writes must be the oracle's data-parallel step from the same seeded initialisation"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import unet3d_oracle as oracle
    save = str(tmp_path / "ck")
    mp.spawn(_cli_worker, args=(2, PORT + 2, save), nprocs=2, join=True)
    files = sorted(os.listdir(save))
    assert "latest_checkpoint.pth" in files and "best_model_epoch_1.pth" in files, files
    got = torch.load(os.path.join(save, "best_model_epoch_1.pth"), map_location=cuda_dev)
    torch.manual_seed(11)
    sd0 = {k: v.detach().clone().to(cuda_dev) for k, v in pkg.UNet3D(5, 1, init_features=16).state_dict().items()}
    loader = pkg.data.get_dataloader(None, batch_size=2, missing_strategy="zero_fill", target_size=(32, 32, 32),
                                     is_training=True, data_type="BPH", n_cases=2, seed=1234)
    batch = next(iter(loader))
    xs = [batch["image"][i:i + 1].to(cuda_dev) for i in range(2)]
    ys = [batch["label"][i:i + 1].to(cuda_dev) for i in range(2)]
    sd = {k: v.clone() for k, v in sd0.items()}
    oracle.dp_train_step(sd, {}, xs, ys, lr=1e-3, weight_decay=1e-5, loss="bce_dice")
    agree = []
    for k in oracle.param_names(sd0):
        if k.endswith(".bias") and (".conv.0." in k or ".conv.3." in k):
            continue
        upd, ref = got[k] - sd0[k], sd[k] - sd0[k]
        assert 0.5e-3 < upd.abs().mean().item() < 1.5e-3, k      # the first Adam step moves every weight by ~lr
        agree.append((torch.sign(upd) == torch.sign(ref)).float().mean().item())
    assert min(agree) > 0.75 and sum(agree) / len(agree) > 0.9, (min(agree), sum(agree) / len(agree))
    for k in got:
        if "running" in k:
            assert _rel(got[k], sd[k]) < 5e-2, k               # BatchNorm buffers of rank 0's shard
    assert int(got["inc.conv.1.num_batches_tracked"].item()) == 1
