"""GPU tests of the reference-facing entry points: BaseTrainer / CrossValidationTrainer loops, checkpoint formats,
ModelPredictor, validation metrics, sliding-window inference against the oracle's restatement, and the CLI."""
import importlib
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from helpers import rel_l2

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unet3d_oracle as oracle  # noqa: E402

pytestmark = pytest.mark.gpu


def _cfg(tmp_path, **kw):
    cfg = {"data_dir": None, "num_epochs": 3, "batch_size": 2, "learning_rate": 1e-3, "device": "cuda:0",
           "save_dir": str(tmp_path), "data_type": "BPH", "handle_missing_modalities": "zero_fill",
           "validation": True, "init_features": 16, "target_size": (16, 16, 16), "n_cases": 6, "loss": "bce_dice",
           "seed": 0}
    cfg.update(kw)
    return cfg


def test_base_trainer_trains_and_checkpoints(pkg, cuda_dev, tmp_path):
    tr = pkg.BaseTrainer(_cfg(tmp_path))
    for attr in ("model", "criterion", "optimizer", "scheduler", "train_loader", "val_loader", "device", "config"):
        assert hasattr(tr, attr)
    first = tr.train_epoch()
    best = tr.train()
    # the training loss falls over the three further epochs (`best` is the monitored validation loss)
    assert tr.history[-1]["train_loss"] < first and best == min(h["val_loss"] for h in tr.history)
    files = os.listdir(tmp_path)
    assert "latest_checkpoint.pth" in files and any(f.startswith("best_model_epoch_") for f in files)
    ckpt = torch.load(os.path.join(tmp_path, "latest_checkpoint.pth"), weights_only=False)
    assert set(ckpt) >= {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "loss", "config"}
    assert len(ckpt["model_state_dict"]) == 136 - 0 or len(ckpt["model_state_dict"]) > 100
    # resume
    tr2 = pkg.BaseTrainer(_cfg(tmp_path))
    assert tr2.load_checkpoint(os.path.join(tmp_path, "latest_checkpoint.pth")) >= 1
    assert torch.equal(tr2.model.outc.weight.cpu(), ckpt["model_state_dict"]["outc.weight"])
    # both checkpoint containers load into the predictor (script/predict.py:139-145)
    best_file = [f for f in files if f.startswith("best_model_epoch_")][0]
    for f in ("latest_checkpoint.pth", best_file):
        pred = pkg.ModelPredictor(os.path.join(tmp_path, f), "cuda:0", init_features=16)
        vol = pkg.data.SyntheticProstateDataset(1, (16, 16, 16))[0]["image"].numpy()
        out = pred.predict(pkg.preprocess_image(vol))
        assert out.shape == (16, 16, 16) and out.dtype == np.float32 and 0.0 <= out.min() and out.max() <= 1.0
    mask = pred.save_prediction(out, os.path.join(tmp_path, "pred.npy"))
    assert mask.dtype == np.uint8 and os.path.exists(os.path.join(tmp_path, "pred.npy"))
    assert pkg.Trainer is pkg.BaseTrainer and issubclass(pkg.BPHTrainer, pkg.BaseTrainer)


def test_torch_optimizer_and_grad_clip_variants(pkg, cuda_dev, tmp_path):
    tr = pkg.BPHTrainer(_cfg(tmp_path, optimizer="torch", clip_grad_norm=1.0, validation=False, num_epochs=1))
    assert tr.config["data_type"] == "BPH" and tr.val_loader is None
    a = tr.train_epoch()
    tr2 = pkg.BaseTrainer(_cfg(tmp_path, clip_grad_norm=1.0, validation=False, num_epochs=1))
    b = tr2.train_epoch()
    assert np.isfinite(a) and np.isfinite(b)
    assert tr2.optimizer.grad_scale <= 1.0


def test_cross_validation_trainer(pkg, cuda_dev, tmp_path):
    cv = pkg.CrossValidationTrainer(_cfg(tmp_path, n_splits=2, num_epochs=2, n_cases=4))
    res = cv.train()
    assert len(res) == 2 and all(np.isfinite(r["best_val_loss"]) for r in res)
    data = json.load(open(os.path.join(tmp_path, "cv_results.json")))
    assert data["n_splits"] == 2 and len(data["folds"]) == 2
    assert os.path.exists(os.path.join(tmp_path, "best_model_fold_0.pth"))
    x = torch.zeros(1, 1, 4, 4, 4)
    assert cv._fix_labels(x, torch.zeros(1, 4, 4, 4)).shape == x.shape
    assert cv._fix_labels(x, torch.zeros(1, 1, 8, 8, 8)).shape == x.shape


def test_cross_validation_fold_parallel_schedule(pkg, cuda_dev, tmp_path):
    """fold-parallel mode: rank r of `world` trains folds r, r + world, ... (replicas only); emulated rank by rank"""
    done = {}
    for rank in range(2):
        cv = pkg.CrossValidationTrainer(_cfg(tmp_path, n_splits=3, num_epochs=1, n_cases=6, fold_parallel=True))
        res = cv.train(rank=rank, world=2)
        done[rank] = sorted(r["fold"] for r in res)
    assert done == {0: [0, 2], 1: [1]}
    assert all(os.path.exists(os.path.join(tmp_path, f"best_model_fold_{k}.pth")) for k in range(3))
    # without the flag the explicit rank is ignored: every fold runs, as in the reference
    cv = pkg.CrossValidationTrainer(_cfg(tmp_path, n_splits=2, num_epochs=1, n_cases=4))
    assert sorted(r["fold"] for r in cv.train(rank=1, world=2)) == [0, 1]


def test_validation_metrics(pkg, cuda_dev):
    val = pkg.validate
    a = torch.zeros(4, 4, 4); a[:2] = 1
    b = torch.zeros(4, 4, 4); b[1:3] = 1
    assert abs(val.calculate_dice_score(a, b) - 0.5) < 1e-6
    assert abs(val.calculate_iou(a, b) - 1 / 3) < 1e-6
    assert abs(val.calculate_dice_score(torch.zeros(2, 2, 2), torch.zeros(2, 2, 2)) - 1.0) < 1e-6


def test_sliding_window_inference_vs_oracle(pkg, cuda_dev):
    torch.manual_seed(1)
    model = pkg.UNet3D(5, 1, init_features=16).to(cuda_dev)
    # eval-mode statistics that are not the identity
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x = torch.rand(2, 5, 48, 40, 32, device=cuda_dev)
    window, stride = (32, 32, 32), (16, 16, 16)
    ref = oracle.sliding_window_logits(x, sd, window, stride)
    got = pkg.parallel.sliding_window_logits(model, x, window, stride)
    assert got.shape == ref.shape == (2, 1, 48, 40, 32)
    assert rel_l2(got, ref) < 2e-2
    # sharding the window list over ranks changes nothing but the order of additions
    parts = [pkg.parallel.sliding_window_logits(model, x, window, stride, rank=r, world=3, reduce=False)
             for r in range(3)]
    merged = sum(p[0] for p in parts)
    pkg.ops.window_finalize(merged, parts[0][1])   # divide by the per-voxel window count
    assert torch.allclose(merged, got, rtol=1e-5, atol=1e-5)
    # contiguous blocks of the schedule: volume 0 is shared by ranks 0 and 1 and owned by rank 0
    assert pkg.parallel.volume_owners(tuple(x.shape), window, stride, 3) == {0: (0, [0, 1]), 1: (1, [1, 2])}
    sched, _ = pkg.parallel.window_schedule(x.shape, window, stride)
    assert len(sched) == 2 * 2 * 2 * 1   # origins (0,16) x (0,8) x (0,) per volume
    probs, mask = pkg.parallel.sliding_window_predict(model, x, window, stride)
    assert torch.equal(mask, (probs > 0.5).float())
    sure = ref.abs() > 0.05 * ref.abs().mean()
    assert torch.equal(mask[sure], (ref > 0).float()[sure])
    # whole-volume prediction (reference semantics) is the one-window special case
    whole = pkg.parallel.sliding_window_logits(model, x, (64, 64, 64), (64, 64, 64))
    assert torch.allclose(whole, model(x), atol=1e-5)
    del parts


def test_cli_train_validate_predict(pkg, cuda_dev, tmp_path):
    cli = importlib.import_module(pkg.__name__ + ".cli")
    save = str(tmp_path / "ck")
    common = ["--init_features", "16", "--size", "16", "16", "16", "--n_cases", "4"]
    cli.main(["train", "--epochs", "1", "--save_dir", save] + common)
    best = [f for f in os.listdir(save) if f.startswith("best_model_epoch_")]
    assert best
    model_path = os.path.join(save, best[0])
    s = cli.main(["validate", "--model_path", model_path, "--output_dir", str(tmp_path / "val")] + common)
    assert 0.0 <= s["mean_dice"] <= 1.0
    assert os.path.exists(tmp_path / "val" / "validation_results.json")
    out = cli.main(["predict", "--model_path", model_path, "--output_dir", str(tmp_path / "pred")] + common)
    assert out.shape == (16, 16, 16)
    rep = cli.main(["check"])
    assert rep["abi_version"] == importlib.import_module(pkg.__name__ + "._lib").ABI_VERSION and rep["sm_count"] > 0
    # failures are reported, not raised (run.py:339-344)
    assert cli.main(["predict", "--model_path", "/nonexistent.pth"] + common) is None
