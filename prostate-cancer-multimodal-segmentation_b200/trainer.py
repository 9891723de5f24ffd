"""Host-side training loops with the reference's entry points and config keys (utils/trainer.py:23-345,
train_bph_optimized.py:34-475): BaseTrainer, BPHTrainer, CrossValidationTrainer.  The loops stay Python; every step's
arithmetic (forward, loss, backward, optimizer) runs in the B200 kernels.

Differences from the reference, all on purpose:
  * the as-shipped constructor cannot run on current torch (ReduceLROnPlateau(verbose=...), get_dataloader(mode=...),
    SURVEY.md 3.5); the same config keys are honoured and mapped to working calls;
  * data comes from `config['train_loader']` / `config['val_loader']` when given, else from the synthetic source in
    data.py (NIfTI IO is out of scope);
  * the optimizer is the fused Adam (same update rule); `config['optimizer'] = 'torch'` selects torch.optim.Adam;
  * optional `config['clip_grad_norm']` (train_bph.py:166) is folded into the fused optimizer's gradient scale;
  * under `python -m torch.distributed.run` (one process per GPU) the trainers are data-parallel: identical replicas
    (rank 0's weights and BatchNorm buffers are broadcast), every global batch of `config['batch_size']` samples is
    split contiguously over the ranks, BatchNorm statistics and the loss stay rank-local, the flat gradient is
    all-reduced in buckets while backward still runs (parallel.GradSync) and 1/world is folded into the fused Adam.
    Epoch losses are averaged over ranks (so scheduler and early stopping agree everywhere); rank 0 alone writes
    checkpoints, after every rank has taken rank 0's BatchNorm buffers.
"""
import json
import os

import torch
from torch import optim

from . import data as _data
from . import ops
from . import parallel as _par
from .losses import BCEDiceLoss, DiceLoss
from .optim import FusedAdam
from .unet3d import UNet3D


class BaseTrainer:
    max_patience = 20  # utils/trainer.py:306

    def __init__(self, config):
        self.config = config
        self.rank, self.world = _par.dist_info()
        if self.world > 1 and "device" not in config:
            self.device = torch.device("cuda", torch.cuda.current_device())   # set by parallel.init_distributed
        else:
            self.device = torch.device(config.get("device", "cuda"))
        self.model = self._create_model()
        self.criterion = self._create_criterion()
        self.optimizer = self._create_optimizer()
        self.scheduler = self._create_scheduler()
        self.grad_sync = None
        if self.world > 1:
            if not isinstance(self.optimizer, FusedAdam):
                raise ValueError("data-parallel training needs the fused optimizer (config['optimizer'] = 'fused'): "
                                 "the 1/world gradient scale is folded into its update")
            self.grad_sync = _par.make_data_parallel(self.model, self.optimizer,
                                                     bucket_mb=config.get("bucket_mb", 25.0))
        self.train_loader = self._create_dataloader("train")
        self.val_loader = self._create_dataloader("test") if config.get("validation", False) else None
        if self.rank == 0:
            os.makedirs(config["save_dir"], exist_ok=True)
        self.history = []
        self._graphed = None

    def close(self):
        """drop the recorded step (it holds NCCL kernels: must go before the process group does)"""
        self._graphed = None
        import gc
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    # ---- factories (names as in the reference)
    def _create_model(self):
        if self.config.get("seed") is not None:   # reproducible initialisation (the reference never seeds)
            torch.manual_seed(int(self.config["seed"]))
        return UNet3D(n_modalities=5, n_classes=1,
                      init_features=self.config.get("init_features", 64)).to(self.device)

    def _create_criterion(self):
        return BCEDiceLoss() if self.config.get("loss", "dice") == "bce_dice" else DiceLoss()

    def _create_optimizer(self):
        lr = self.config["learning_rate"]
        if self.config.get("optimizer", "fused") == "torch":
            return optim.Adam(self.model.parameters(), lr=lr, weight_decay=1e-5)
        return FusedAdam(self.model, lr=lr, weight_decay=1e-5)

    def _create_scheduler(self):
        return optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, mode="min", patience=10, factor=0.5)

    def _create_dataloader(self, mode):
        key = "train_loader" if mode == "train" else "val_loader"
        if self.config.get(key) is not None:
            loader = self.config[key]
        else:
            loader = self._synthetic_loader(mode)
        # data-parallel: the loader yields GLOBAL batches, every rank takes its contiguous share
        return _par.ShardedLoader(loader, self.rank, self.world) if self.world > 1 else loader

    def _synthetic_loader(self, mode):
        return _data.get_dataloader(
            self.config.get("data_dir"), batch_size=self.config["batch_size"],
            missing_strategy=self.config.get("handle_missing_modalities", "zero_fill"),
            target_size=tuple(self.config.get("target_size", (128, 128, 128))), is_training=(mode == "train"),
            data_type=self.config.get("data_type", "BPH"), n_cases=self.config.get("n_cases", 8),
            seed=1234 if mode == "train" else 4321)

    # ---- one step / epoch
    def _step(self, images, labels):
        clip = self.config.get("clip_grad_norm")
        if isinstance(self.optimizer, FusedAdam) and not clip and self.config.get("cuda_graph", True):
            # the whole step replayed from a CUDA graph after two eager steps (graph.py); same arithmetic
            if getattr(self, "_graphed", None) is None:
                from .graph import GraphedTrainStep
                # data-parallel ranks record the NCCL bucket all-reduces into the graph as well (gloo cannot be
                # captured: those ranks launch eagerly)
                self._graphed = GraphedTrainStep(self.model, self.criterion, self.optimizer,
                                                 capture_collectives=self.world > 1 and _par.backend_is_nccl())
            return self._graphed(images, labels)
        self.optimizer.zero_grad()
        outputs = self.model(images)
        loss = self.criterion(outputs, labels)
        loss.backward()
        clip = self.config.get("clip_grad_norm")
        if clip and isinstance(self.optimizer, FusedAdam):
            eng = self.model.engine
            acc = torch.zeros(2, device=self.device)
            ops.sumsq(eng.flat_grad, acc)
            norm = float(acc[0].sqrt())
            self.optimizer.grad_scale = min(1.0, clip / (norm + 1e-6))
        elif clip:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), clip)
        self.optimizer.step()
        return loss

    def train_epoch(self):
        self.model.train()
        total, n = 0.0, 0
        losses = _data.AsyncScalarReader()  # per-step loss read back one step late: no GPU drain between steps
        for batch in _data.DevicePrefetcher(self.train_loader, self.device):  # next batch's H2D under this step
            losses.push(self._step(batch["image"], batch["label"]))
            n += 1
        total = sum(losses.finish())
        return _par.all_mean(total / max(n, 1), self.device)

    def validate_epoch(self):
        if self.val_loader is None:
            return None
        _par.sync_buffers(self.model)   # every rank evaluates with rank 0's running statistics
        self.model.eval()
        total, n = 0.0, 0
        with torch.no_grad():
            for batch in _data.DevicePrefetcher(self.val_loader, self.device):
                total += self.criterion(self.model(batch["image"]), batch["label"]).item()
                n += 1
        return _par.all_mean(total / max(n, 1), self.device)

    def save_checkpoint(self, epoch, loss, is_best=False):
        """latest_checkpoint.pth (dict) and best_model_epoch_{E}.pth (raw state_dict): utils/trainer.py:255-278.
        Data-parallel: every rank calls this (the buffer broadcast is a collective), rank 0 writes."""
        _par.sync_buffers(self.model)
        if self.rank != 0:
            return
        sd = {k: v.detach().clone().cpu() for k, v in self.model.state_dict().items()}
        ckpt = {"epoch": epoch, "model_state_dict": sd, "optimizer_state_dict": self.optimizer.state_dict(),
                "scheduler_state_dict": self.scheduler.state_dict(), "loss": loss,
                "config": {k: v for k, v in self.config.items() if not k.endswith("_loader")}}
        torch.save(ckpt, os.path.join(self.config["save_dir"], "latest_checkpoint.pth"))
        if is_best:
            path = os.path.join(self.config["save_dir"], f"best_model_epoch_{epoch}.pth")
            torch.save(sd, path)
            print(f"best model saved to {path}")

    def load_checkpoint(self, path):
        """resume (the reference documents resume, README.md:275, but never implemented it)"""
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        self.model.load_state_dict(ckpt["model_state_dict"] if "model_state_dict" in ckpt else ckpt)
        if "optimizer_state_dict" in ckpt:
            self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        if "scheduler_state_dict" in ckpt:
            self.scheduler.load_state_dict(ckpt["scheduler_state_dict"])
        return ckpt.get("epoch", 0)

    def train(self):
        cfg = self.config
        say = print if self.rank == 0 else (lambda *a, **k: None)
        say(f"training {cfg.get('data_type', 'BPH')}: epochs {cfg['num_epochs']}, batch {cfg['batch_size']}, "
            f"lr {cfg['learning_rate']}, device {self.device}, ranks {self.world}, save_dir {cfg['save_dir']}")
        best, patience = float("inf"), 0
        for epoch in range(cfg["num_epochs"]):
            train_loss = self.train_epoch()
            val_loss = self.validate_epoch()
            monitored = val_loss if val_loss is not None else train_loss
            say(f"epoch {epoch + 1}/{cfg['num_epochs']}: train {train_loss:.4f}"
                + (f", val {val_loss:.4f}" if val_loss is not None else ""))
            self.history.append({"epoch": epoch + 1, "train_loss": train_loss, "val_loss": val_loss})
            self.scheduler.step(monitored)
            if monitored < best:
                best, patience = monitored, 0
                self.save_checkpoint(epoch + 1, monitored, is_best=True)
            else:
                patience += 1
            if patience >= self.max_patience:
                say(f"early stop: {self.max_patience} epochs without improvement")
                break
        say(f"done, best loss {best:.4f}")
        return best


Trainer = BaseTrainer  # run.py:30 imports this name


class BPHTrainer(BaseTrainer):
    """train_bph_optimized.py:34-75: BaseTrainer pinned to the BPH cohort"""

    def __init__(self, config):
        config = dict(config)
        config["data_type"] = "BPH"
        super().__init__(config)


class CrossValidationTrainer:
    """5-fold cross-validation with the AMP-style loop of train_bph_optimized.py:78-475 (per-fold fresh model,
    optimizer and scheduler; early stop after 15 stale epochs; best_model_fold_{k}.pth; cv_results.json)."""
    max_patience = 15

    def __init__(self, config):
        self.config = config
        self.rank, self.world = _par.dist_info()
        # config['fold_parallel']: folds are dealt to the ranks (replicas only, no gradient exchange); otherwise
        # every fold is trained data-parallel over all ranks, folds one after another (train_bph_optimized.py:428-429)
        self.dp = self.world > 1 and not config.get("fold_parallel", False)
        if self.world > 1 and "device" not in config:
            self.device = torch.device("cuda", torch.cuda.current_device())
        else:
            self.device = torch.device(config.get("device", "cuda"))
        self.n_splits = config.get("n_splits", 5)
        self.n_cases = config.get("n_cases", 10)
        self.splits = _data.get_kfold_splits(self.n_cases, self.n_splits)
        self.fold_results = []
        os.makedirs(config["save_dir"], exist_ok=True)

    def _create_model(self):
        return UNet3D(5, 1, init_features=self.config.get("init_features", 64)).to(self.device)

    def _loader(self, indices, training):
        loader = _data.get_dataloader(self.config.get("data_dir"), batch_size=self.config["batch_size"],
                                      missing_strategy=self.config.get("handle_missing_modalities", "zero_fill"),
                                      target_size=tuple(self.config.get("target_size", (128, 128, 128))),
                                      is_training=training, data_type=self.config.get("data_type", "BPH"),
                                      indices=indices, n_cases=self.n_cases)
        return _par.ShardedLoader(loader, self.rank, self.world) if self.dp else loader

    @staticmethod
    def _fix_labels(outputs, labels):
        # train_bph_optimized.py:273-291: add the channel axis / nearest-resize labels to the output grid
        if labels.dim() == 4:
            labels = labels.unsqueeze(1)
        if labels.shape[2:] != outputs.shape[2:]:
            labels = torch.nn.functional.interpolate(labels, size=outputs.shape[2:], mode="nearest")
        return labels

    def train_fold(self, fold_idx, train_idx, val_idx):
        model = self._create_model()
        opt = FusedAdam(model, lr=self.config["learning_rate"], weight_decay=1e-5)
        if self.dp:
            _par.make_data_parallel(model, opt, bucket_mb=self.config.get("bucket_mb", 25.0))
        sched = optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", patience=10, factor=0.5)
        crit = DiceLoss()
        train_loader, val_loader = self._loader(train_idx, True), self._loader(val_idx, False)
        scaler = torch.amp.GradScaler("cuda", enabled=False)  # bf16 kernels need no loss scaling; API kept
        best, patience, hist = float("inf"), 0, []
        for epoch in range(self.config["num_epochs"]):
            model.train()
            tl, n = 0.0, 0
            for batch in _data.DevicePrefetcher(train_loader, self.device):
                images, labels = batch["image"], batch["label"]
                opt.zero_grad()
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    outputs = model(images)
                    loss = crit(outputs, self._fix_labels(outputs, labels))
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
                tl += loss.item(); n += 1
            if self.dp:
                _par.sync_buffers(model)
            model.eval()
            vl, m = 0.0, 0
            with torch.no_grad():
                for batch in _data.DevicePrefetcher(val_loader, self.device):
                    images, labels = batch["image"], batch["label"]
                    outputs = model(images)
                    vl += crit(outputs, self._fix_labels(outputs, labels)).item(); m += 1
            tl, vl = tl / max(n, 1), vl / max(m, 1)
            if self.dp:
                tl, vl = _par.all_mean(tl, self.device), _par.all_mean(vl, self.device)
            hist.append({"epoch": epoch + 1, "train_loss": tl, "val_loss": vl})
            sched.step(vl)
            if vl < best:
                best, patience = vl, 0
                self.save_best_model(model, fold_idx, epoch + 1, vl)
            else:
                patience += 1
            if patience >= self.max_patience:
                break
        result = {"fold": fold_idx, "best_val_loss": best, "history": hist, "train_idx": list(train_idx),
                  "val_idx": list(val_idx)}
        self.fold_results.append(result)
        return result

    def save_best_model(self, model, fold_idx, epoch, loss):
        if self.dp:
            _par.sync_buffers(model)
            if self.rank != 0:
                return
        sd = {k: v.detach().clone().cpu() for k, v in model.state_dict().items()}
        torch.save({"epoch": epoch, "model_state_dict": sd, "fold_idx": fold_idx, "loss": loss,
                    "config": {k: v for k, v in self.config.items() if not k.endswith("_loader")}},
                   os.path.join(self.config["save_dir"], f"best_model_fold_{fold_idx}.pth"))

    def train(self, rank=None, world=None):
        """Folds are independent models (train_bph_optimized.py:428-429 runs them one after another).  With
        `config['fold_parallel']` and an initialised process group (or explicit rank / world) every rank trains folds
        rank, rank + world, ... on its own GPU — replicas only, no data-path collective — and the per-fold results are
        gathered on every rank before `cv_results.json` is written by rank 0."""
        import torch.distributed as dist
        parallel = bool(self.config.get("fold_parallel", False))
        if parallel and rank is None and dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
        if not parallel or rank is None:
            rank, world = 0, 1
        for fold_idx, (tr, va) in enumerate(self.splits):
            if fold_idx % world == rank:
                self.train_fold(fold_idx, tr, va)
        if self.dp:   # every rank trained every fold together and holds the same (rank-averaged) results
            rank = self.rank
        elif world > 1 and dist.is_available() and dist.is_initialized():
            parts = [None] * world
            dist.all_gather_object(parts, self.fold_results)
            self.fold_results = sorted((r for part in parts for r in part), key=lambda r: r["fold"])
        if rank == 0:
            self.save_results()
            self.print_summary()
        return self.fold_results

    def save_results(self):
        with open(os.path.join(self.config["save_dir"], "cv_results.json"), "w") as f:
            json.dump({"n_splits": self.n_splits, "folds": self.fold_results}, f, indent=1)

    def print_summary(self):
        vals = [r["best_val_loss"] for r in self.fold_results]
        if vals:
            mean = sum(vals) / len(vals)
            print(f"cross-validation: {len(vals)} folds, best val loss mean {mean:.4f} "
                  f"(min {min(vals):.4f}, max {max(vals):.4f})")
