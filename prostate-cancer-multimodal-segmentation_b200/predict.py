"""Inference entry points with the reference's names (script/predict.py:84-197): preprocess_image, ModelPredictor
(.predict -> (D,H,W) float probabilities, .save_prediction -> thresholded uint8 volume).  Whole-volume prediction is
the reference semantics; `window=` switches to sliding-window inference (BASELINE cfg #4), optionally sharded over
the ranks of a process group."""
import os

import numpy as np
import torch

from . import parallel
from .unet3d import UNet3D


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
