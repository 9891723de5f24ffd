"""B200-native (sm_100a) hot path of the 5-modality prostate-MRI 3D U-Net: drop-in for the reference's
``models/unet3d.py`` (UNet3D), ``utils/losses.py`` (DiceLoss, BCEDiceLoss) and its trainer / predict entry points.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic on the path runs in the
hand-written CUDA kernels behind the C ABI of ``include/b200_unet3d.h`` (``libb200unet3d.so``, built in-tree).
"""
from ._lib import B200Error, load as load_library, lib_path  # noqa: F401
from . import ops  # noqa: F401
from .unet3d import UNet3D, DoubleConv3D, Down3D, Up3D  # noqa: F401
from .losses import DiceLoss, BCEDiceLoss  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .graph import GraphedTrainStep  # noqa: F401
from . import data, parallel, predict, trainer, validate  # noqa: F401
from .trainer import BaseTrainer, BPHTrainer, CrossValidationTrainer, Trainer  # noqa: F401
from .predict import ModelPredictor, load_multimodal_images, preprocess_image  # noqa: F401

__all__ = ["B200Error", "load_library", "lib_path", "ops", "UNet3D", "DoubleConv3D", "Down3D", "Up3D", "DiceLoss",
           "BCEDiceLoss", "FusedAdam", "GraphedTrainStep", "BaseTrainer", "BPHTrainer", "CrossValidationTrainer", "Trainer",
           "ModelPredictor", "load_multimodal_images", "preprocess_image", "data", "parallel", "predict", "trainer", "validate"]
