"""Drop-in for the reference's ``models/unet3d.py``: same classes, constructor signatures, attribute names, parameter
and buffer names (136 state_dict entries at base 64), same initialisation RNG order — but ``UNet3D.forward`` runs the
B200 engine (tcgen05 implicit-GEMM convolutions, fused BatchNorm/ReLU, in-place skip concat) instead of torch.nn ops.

The torch.nn layer objects below are *parameter containers*: they give identical ``state_dict`` keys, identical
default/Kaiming initialisation (models/unet3d.py:227-245) and keep ``isinstance`` checks in user code working.  Inside a
``UNet3D`` the engine schedules the whole network; called on their own, ``DoubleConv3D`` / ``Down3D`` / ``Up3D`` run the
same kernels block by block (blocks.py).
"""
import torch
import torch.nn as nn

from . import ops
from ._lib import B200Error
from .engine import Engine


class DoubleConv3D(nn.Module):
    """(Conv3d 3x3x3 pad 1 -> BatchNorm3d -> ReLU) x 2 — reference models/unet3d.py:5-55"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv3d(in_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm3d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv3d(out_channels, out_channels, kernel_size=3, padding=1),
            nn.BatchNorm3d(out_channels),
            nn.ReLU(inplace=True),
        )

    def forward(self, x):
        """stand-alone call (inside a UNet3D the network engine schedules the block): same B200 kernels"""
        from .blocks import run_block
        return run_block(self, "double", x)


class Down3D(nn.Module):
    """MaxPool3d(2) then DoubleConv3D — reference models/unet3d.py:58-96"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool3d(2), DoubleConv3D(in_channels, out_channels))

    def forward(self, x):
        from .blocks import run_block
        return run_block(self, "down", x)


class Up3D(nn.Module):
    """ConvTranspose3d(k2,s2) -> pad to the skip -> cat([skip, up]) -> DoubleConv3D — reference models/unet3d.py:99-158"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.up = nn.ConvTranspose3d(in_channels, in_channels // 2, kernel_size=2, stride=2)
        self.conv = DoubleConv3D(in_channels, out_channels)

    def forward(self, x1, x2):
        """x1: the coarser feature map (upsampled 2x), x2: the skip connection (models/unet3d.py:134-158)"""
        from .blocks import run_block
        return run_block(self, "up", x1, x2)


class _UNetFunction(torch.autograd.Function):
    """one autograd node for the whole network: forward and backward are engine schedules"""

    @staticmethod
    def forward(ctx, model, x, *params):
        logits, _, tape = model._engine.forward(x, training=model.training)
        ctx.model = model
        ctx.tape = tape
        ctx.saved_nothing = any(st is None for st in tape.dcs.values())   # eval-mode forward: BatchNorm folded
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        tape, ctx.tape = ctx.tape, None
        if tape is None:
            raise RuntimeError("UNet3D (B200): backward called twice on the same forward; activations were freed")
        if ctx.saved_nothing:
            raise B200Error("UNet3D (B200): backward through an eval-mode forward is not supported (BatchNorm is folded "
                            "into the convolution epilogue and nothing is saved); call model.train() first, or run "
                            "the forward under torch.no_grad()")
        ctx.model._engine.backward(tape, dlogits)
        # parameter gradients were accumulated straight into Parameter.grad (views of the flat gradient buffer);
        # the input gets no gradient (the reference never asks for one: images do not require grad)
        return (None, None) + (None,) * (len(ctx.needs_input_grad) - 2)


class UNet3D(nn.Module):
    """3D U-Net for 5-modality prostate MRI — reference models/unet3d.py:160-344.

    ``init_features`` (default 64, the value hard-coded at models/unet3d.py:190) is the only extension.
    """

    def __init__(self, n_modalities=5, n_classes=2, init_features=64):
        super().__init__()
        if init_features % 16 != 0:
            raise ValueError("init_features must be a multiple of 16 (tcgen05 K/N granularity)")
        self.n_modalities = n_modalities
        self.n_classes = n_classes
        self.init_features = init_features
        f = init_features
        self.inc = DoubleConv3D(n_modalities, f)
        self.down1 = Down3D(f, f * 2)
        self.down2 = Down3D(f * 2, f * 4)
        self.down3 = Down3D(f * 4, f * 8)
        self.down4 = Down3D(f * 8, f * 16)
        self.up1 = Up3D(f * 16, f * 8)
        self.up2 = Up3D(f * 8, f * 4)
        self.up3 = Up3D(f * 4, f * 2)
        self.up4 = Up3D(f * 2, f)
        self.outc = nn.Conv3d(f, n_classes, kernel_size=1)
        self._init_weights()
        object.__setattr__(self, "_engine", Engine(self))

    def _init_weights(self):
        # same traversal and initialisers as the reference (ConvTranspose3d keeps torch's default init)
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    @property
    def engine(self) -> Engine:
        return self._engine

    def forward(self, x):
        """logits (N, n_classes, D, H, W) fp32 for x (N, n_modalities, D, H, W)"""
        eng = self._engine
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if not all(p.requires_grad for p in self.parameters()):
                # the engine's backward produces every parameter gradient in one schedule; a silently updated
                # "frozen" layer would be worse than refusing
                raise B200Error("UNet3D (B200): freezing a subset of the parameters (requires_grad=False) is not "
                                "supported; train all parameters or run under torch.no_grad()")
            eng.prepare(x.device)  # parameters must be in their final (flat) storage before autograd sees them
            return _UNetFunction.apply(self, x, *[p for _, p in eng.ordered_params()])
        logits, _, _ = eng.forward(x, training=self.training)
        return logits

    def predict(self, x):
        """eval + no_grad + sigmoid, as reference models/unet3d.py:298-318 (sigmoid fused into the head kernel)"""
        self.eval()
        with torch.no_grad():
            _, probs, _ = self._engine.forward(x, training=False, want_probs=True)
            return probs

    def inference(self, x, threshold=0.5):
        """binary mask (float 0/1), reference models/unet3d.py:320-344"""
        probs = self.predict(x)
        return (probs > threshold).float()
