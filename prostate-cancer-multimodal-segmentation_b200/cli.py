"""Command line with the reference's sub-commands and flags (run.py:242-344): check | train | validate | predict.
The reference bodies import names that do not exist; here each sub-command is wired to working code."""
import argparse
import json
import os
import sys
import traceback

import numpy as np
import torch


def _add_common(p):
    p.add_argument("--data_type", choices=["BPH", "PCA"], default="BPH")
    p.add_argument("--missing_strategy", choices=["zero_fill", "skip", "duplicate"], default="zero_fill")
    p.add_argument("--device", default="cuda")
    p.add_argument("--init_features", type=int, default=64)
    p.add_argument("--size", type=int, nargs=3, default=[128, 128, 128], help="synthetic volume extent")
    p.add_argument("--n_cases", type=int, default=8)


def build_parser():
    ap = argparse.ArgumentParser(prog="run.py", description="prostate multimodal MRI segmentation (B200 hot path)")
    sub = ap.add_subparsers(dest="command", required=True)
    sub.add_parser("check", help="environment / build / device report")
    t = sub.add_parser("train")
    _add_common(t)
    t.add_argument("--epochs", type=int, default=10)
    t.add_argument("--batch_size", type=int, default=2)
    t.add_argument("--learning_rate", type=float, default=1e-4)
    t.add_argument("--optimized", action="store_true")
    t.add_argument("--cross_validation", action="store_true")
    t.add_argument("--save_dir", default="checkpoints")
    t.add_argument("--seed", type=int, default=None, help="seed of the weight initialisation")
    t.add_argument("--loss", choices=["dice", "bce_dice"], default="dice")
    t.add_argument("--fold_parallel", action="store_true",
                   help="multi-GPU cross-validation: deal the folds to the ranks instead of training every fold "
                        "data-parallel over all ranks")
    v = sub.add_parser("validate")
    _add_common(v)
    v.add_argument("--model_path", required=True)
    v.add_argument("--output_dir", default="validation_results")
    p = sub.add_parser("predict")
    _add_common(p)
    p.add_argument("--model_path", required=True)
    p.add_argument("--input_dir", default=None,
                   help="case directory with one <modality>.npy per modality (ADC, DWI, T2 fs, T2 not fs, gaoqing-T2), "
                        "or one .npy file of shape (5,D,H,W); a synthetic volume if omitted")
    p.add_argument("--output_dir", default="predictions")
    p.add_argument("--window", type=int, nargs=3, default=None, help="sliding-window extent (D H W)")
    p.add_argument("--stride", type=int, nargs=3, default=None)
    return ap


def cmd_check(args):
    from . import build as _build, load_library, lib_path
    report = {"torch": torch.__version__, "cuda_available": torch.cuda.is_available(), "library": lib_path(),
              "library_built": os.path.exists(lib_path())}
    try:
        lib = load_library()
        report["abi_version"] = lib.b200_abi_version()
        report["sm_count"] = lib.b200_sm_count()
    except Exception as e:  # report, do not raise: this is the diagnostic command
        report["library_error"] = str(e)
    if torch.cuda.is_available():
        report["device"] = torch.cuda.get_device_name(0)
        report["capability"] = list(torch.cuda.get_device_capability(0))
    print(json.dumps(report, indent=1))
    return report


def _config(args):
    return {"data_dir": None, "num_epochs": args.epochs, "batch_size": args.batch_size,
            "learning_rate": args.learning_rate, "device": args.device, "save_dir": args.save_dir,
            "data_type": args.data_type, "handle_missing_modalities": args.missing_strategy, "validation": True,
            "init_features": args.init_features, "target_size": tuple(args.size), "n_cases": args.n_cases,
            "n_splits": 5, "seed": getattr(args, "seed", None), "loss": getattr(args, "loss", "dice")}


def _distributed():
    """(rank, world, device-or-None): under `python -m torch.distributed.run` (WORLD_SIZE > 1) this process is one
    data-parallel rank bound to GPU LOCAL_RANK; otherwise a plain single-process run"""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return 0, 1, None
    from . import parallel
    return parallel.init_distributed()


def cmd_train(args):
    from . import parallel
    from .trainer import BaseTrainer, BPHTrainer, CrossValidationTrainer
    rank, world, dev = _distributed()
    cfg = _config(args)
    if dev is not None:
        cfg["device"] = str(dev)
        cfg["fold_parallel"] = bool(getattr(args, "fold_parallel", False))
    trainer = None
    try:
        if args.cross_validation:
            return CrossValidationTrainer(cfg).train()
        trainer = (BPHTrainer if args.optimized else BaseTrainer)(cfg)
        return trainer.train()
    finally:
        if trainer is not None:
            trainer.close()
        if world > 1:
            parallel.shutdown_distributed()


def cmd_validate(args):
    from . import data, validate as val
    from .predict import ModelPredictor
    pred = ModelPredictor(args.model_path, args.device, args.init_features)
    loader = data.get_dataloader(batch_size=1, missing_strategy=args.missing_strategy, target_size=tuple(args.size),
                                 is_training=False, data_type=args.data_type, n_cases=args.n_cases, seed=4321)
    rows = val.validate(pred.model, loader, pred.device)
    os.makedirs(args.output_dir, exist_ok=True)
    summary = {"mean_dice": float(np.mean([r["dice"] for r in rows])), "mean_iou": float(np.mean([r["iou"] for r in rows])),
               "cases": rows}
    with open(os.path.join(args.output_dir, "validation_results.json"), "w") as f:
        json.dump(summary, f, indent=1)
    print(f"validated {len(rows)} cases: Dice {summary['mean_dice']:.4f}, IoU {summary['mean_iou']:.4f}")
    return summary


def cmd_predict(args):
    from . import data, parallel
    from .predict import ModelPredictor, load_multimodal_images, normalize_modalities, preprocess_image
    rank, world, dev = _distributed()
    try:
        pred = ModelPredictor(args.model_path, str(dev) if dev is not None else args.device, args.init_features)
        if args.input_dir and os.path.isdir(args.input_dir):
            # a case directory with one .npy volume per modality (script/predict.py:8-82 reads .nii.gz there)
            image, _ = load_multimodal_images(args.input_dir, handle_missing=args.missing_strategy)
        elif args.input_dir:
            image = normalize_modalities(np.load(args.input_dir))
        else:
            image = normalize_modalities(data.SyntheticProstateDataset(1, tuple(args.size))[0]["image"].numpy())
        x = preprocess_image(image)
        # sliding windows are sharded over the ranks; whole-volume prediction is computed by every rank
        out = pred.predict(x, window=tuple(args.window) if args.window else None,
                           stride=tuple(args.stride) if args.stride else None, rank=rank, world=world)
        if rank == 0:
            os.makedirs(args.output_dir, exist_ok=True)
            path = os.path.join(args.output_dir, "prediction.npy")
            mask = pred.save_prediction(out, path)
            print(f"prediction {out.shape} saved to {path}; foreground voxels {int(mask.sum())}")
        return out
    finally:
        if world > 1:
            parallel.shutdown_distributed()


def main(argv=None):
    args = build_parser().parse_args(argv)
    try:
        return {"check": cmd_check, "train": cmd_train, "validate": cmd_validate, "predict": cmd_predict}[args.command](args)
    except Exception:  # run.py:339-344: print the traceback, do not propagate
        traceback.print_exc()
        return None


if __name__ == "__main__":
    main()
    sys.exit(0)
