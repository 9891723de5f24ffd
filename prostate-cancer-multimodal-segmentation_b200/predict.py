"""Inference entry points with the reference's names (script/predict.py:84-197): preprocess_image, ModelPredictor
(.predict -> (D,H,W) float probabilities, .save_prediction -> thresholded uint8 volume).  Whole-volume prediction is
the reference semantics; `window=` switches to sliding-window inference (BASELINE cfg #4), optionally sharded over
the ranks of a process group."""
import os

import numpy as np
import torch

from . import parallel
from .unet3d import UNet3D


PREDICT_MODALITIES = ["ADC", "DWI", "gaoqing-T2", "T2 fs", "T2 not fs"]   # channel order of script/predict.py:24


def _read_volume(path):
    if path.endswith(".npy"):
        return np.load(path)
    import SimpleITK as sitk  # noqa: N813  (NIfTI needs SimpleITK, which is not part of this build)
    return sitk.GetArrayFromImage(sitk.ReadImage(path))


def load_multimodal_images(case_dir, handle_missing="zero_fill"):
    """(image (5, D, H, W) float32 in [0, 1], modality names) of one case directory — script/predict.py:8-82 with the
    same layout, order, missing-modality rules and per-modality min-max normalisation.  Each modality is a
    sub-directory `case_dir/<modality>/` holding one volume file; `.npy` volumes are read natively (`.nii` only when
    SimpleITK is importable: NIfTI IO is outside this build).  A missing sub-directory raises FileNotFoundError; an
    empty one is zero-filled (shape of the first modality found, or 64^3 before any), duplicated from the first
    modality found (`duplicate`), or an error (`skip`), exactly as the reference does."""
    try:
        import SimpleITK  # noqa: F401, N813
        exts = (".npy", ".nii")
    except ImportError:
        exts = (".npy",)
    images, reference = [], None
    for modality in PREDICT_MODALITIES:
        mdir = os.path.join(case_dir, modality)
        if not os.path.exists(mdir):
            raise FileNotFoundError(f"模态目录不存在: {mdir}")
        files = sorted(f for f in os.listdir(mdir) if f.endswith(exts))
        if not files:
            if handle_missing == "zero_fill":
                img = np.zeros_like(reference, dtype=np.float32) if reference is not None else \
                    np.zeros((64, 64, 64), dtype=np.float32)
                print(f"警告: 模态 {modality} 缺失，使用零填充")
            elif handle_missing == "duplicate" and reference is not None:
                img = reference.copy()
                print(f"警告: 模态 {modality} 缺失，使用参考模态填充")
            else:
                raise FileNotFoundError(f"在 {mdir} 中未找到.nii文件")
        else:
            if len(files) > 1:
                print(f"警告: 在 {mdir} 中找到多个.nii文件，将使用第一个文件")
            img = _read_volume(os.path.join(mdir, files[0]))
            if reference is None:
                reference = img
        img = img.astype(np.float32)
        lo, hi = img.min(), img.max()
        img = (img - lo) / (hi - lo) if hi - lo != 0 else np.zeros_like(img, dtype=np.float32)
        images.append(img)
    return np.stack(images, axis=0), list(PREDICT_MODALITIES)


def preprocess_image(image):
    """(5, D, H, W) numpy -> (1, 5, D, H, W) float tensor (script/predict.py:84-101)"""
    return torch.from_numpy(np.ascontiguousarray(image)).float().unsqueeze(0)


def normalize_modalities(image):
    """per-modality min-max to [0, 1] as load_multimodal_images does after reading each file (script/predict.py:69-75)"""
    out = np.zeros_like(image, dtype=np.float32)
    for m in range(image.shape[0]):
        lo, hi = float(image[m].min()), float(image[m].max())
        out[m] = (image[m] - lo) / (hi - lo) if hi > lo else 0.0
    return out


def normalize_modalities_(image):
    """same, in place on a CUDA tensor (M, D, H, W): one min/max reduction + one apply kernel for all modalities"""
    from . import ops
    return ops.minmax_normalize_(image)


class ModelPredictor:
    def __init__(self, model_path, device="cuda", init_features=64):
        self.device = torch.device(device)
        self.init_features = init_features
        self.model = self._load_model(model_path)

    def _load_model(self, model_path):
        model = UNet3D(n_modalities=5, n_classes=1, init_features=self.init_features).to(self.device)
        ckpt = torch.load(model_path, map_location="cpu", weights_only=False)
        # both checkpoint containers of the reference (script/predict.py:139-145)
        model.load_state_dict(ckpt["model_state_dict"] if "model_state_dict" in ckpt else ckpt)
        model.eval()
        return model

    def predict(self, image_tensor, window=None, stride=None, rank=0, world=1):
        with torch.no_grad():
            x = image_tensor.to(self.device)
            if window is None:
                out = self.model.predict(x)
            else:
                out, _ = parallel.sliding_window_predict(self.model, x, window, stride or window, rank=rank,
                                                         world=world)
            return out.squeeze(0).squeeze(0).cpu().numpy()

    def save_prediction(self, prediction, output_path, reference_image_path=None):
        """threshold 0.5 -> uint8 (script/predict.py:185).  NIfTI writing needs SimpleITK, which is not part of this
        build: .npy is written unless SimpleITK is importable."""
        binary = (prediction > 0.5).astype(np.uint8)
        try:
            import SimpleITK as sitk  # noqa: N813
        except ImportError:
            np.save(output_path if output_path.endswith(".npy") else output_path + ".npy", binary)
            return binary
        img = sitk.GetImageFromArray(binary)
        if reference_image_path and os.path.exists(reference_image_path):
            img.CopyInformation(sitk.ReadImage(reference_image_path))
        sitk.WriteImage(img, output_path)
        return binary
