"""Stand-alone execution of the reference's helper modules — DoubleConv3D, Down3D, Up3D (models/unet3d.py:42-55,
84-96, 134-158) called on their own, outside a UNet3D — on the same B200 kernels the network engine schedules.

Tensors cross this boundary as the reference's: fp32 (N, C, D, H, W) in and out, gradients for inputs and parameters
through torch autograd.  Inside, activations are NDHWC bf16 and the block runs as engine._DoubleConv / the transposed
convolution + in-place concat of the network path.  Weights stay ordinary torch parameters (any optimizer); their
bf16 operands are re-packed when a parameter changed.  There is no CPU path.
"""
import torch

from . import ops
from ._lib import B200Error
from .engine import _ConvTPack, _DoubleConv, _Scratch, _Side
from .ops import ActView, new_act


class _Runner:
    """packed operands of one block, rebuilt when the device changes and re-packed when a parameter changes"""

    def __init__(self, module, kind):
        self.module, self.kind = module, kind
        self.device = None
        self.key = None

    def prepare(self, device):
        if device.type != "cuda":
            raise B200Error(f"{type(self.module).__name__} (B200) runs on CUDA tensors only: there is no CPU path")
        m = self.module
        if self.device != device:
            seq = {"double": lambda: m.conv, "down": lambda: m.maxpool_conv[1].conv, "up": lambda: m.conv.conv}[self.kind]()
            self.dc = _DoubleConv(seq, device)
            self.tp = _ConvTPack(m.up, device) if self.kind == "up" else None
            self.scratch = _Scratch(device, self.dc.cout)
            self.device, self.key = device, None
        key = tuple((p.data_ptr(), p._version) for p in m.parameters())
        if key != self.key:
            self.dc.pack()
            if self.tp is not None:
                self.tp.pack()
            self.key = key


def _pack_ncdhw(x, view):
    x = x.detach()
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.float().contiguous()
    ops.pack_input(x, view)


class _BlockFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, n_inputs, *args):
        inputs, params = args[:n_inputs], args[n_inputs:]
        m, kind = runner.module, runner.kind
        x = inputs[0]
        dev = x.device
        with torch.cuda.device(dev):
            runner.prepare(dev)
            dc, tp = runner.dc, runner.tp
            training = m.training
            n = x.shape[0]
            saved = {"kind": kind, "shapes": [tuple(t.shape) for t in inputs]}
            if kind == "double":
                xd = x.detach()
                if xd.dtype != torch.float32 or not xd.is_contiguous():
                    xd = xd.float().contiguous()
                xin = dc.p1.make_input(xd)
            elif kind == "down":
                _, c, d, h, w = x.shape
                if dc.p1.im2col or c % 8:
                    raise B200Error(f"Down3D (B200): {c} input channels; stand-alone pooling blocks need a multiple "
                                    "of 16 channels (every Down3D of the network has >= 64)")
                xa = ActView(new_act(n, d, h, w, dc.p1.cin_pad, dev))
                _pack_ncdhw(x, xa)
                xin = ActView(new_act(n, d // 2, h // 2, w // 2, dc.p1.cin_pad, dev))
                ops.maxpool3d_fwd(xa, xin)
                saved["pool_in"] = xa
            else:  # up: ConvTranspose3d(x1) -> zero pad to the skip's extent -> cat([skip, up]) -> DoubleConv3D
                x1, x2 = inputs
                _, c1, d1, h1, w1 = x1.shape
                _, c2, d2, h2, w2 = x2.shape
                if c1 != tp.cin or c2 + tp.cout != dc.p1.cin or c2 % 8 or tp.cout % 8:
                    raise B200Error(f"Up3D: channels of x1 {c1} / skip {c2} do not match the module "
                                    f"(in {tp.cin}, up {tp.cout}, conv in {dc.p1.cin}; multiples of 8)")
                dd, dh, dw = d2 - 2 * d1, h2 - 2 * h1, w2 - 2 * w1
                if min(dd, dh, dw) < 0:
                    raise B200Error("Up3D: the skip tensor is smaller than the upsampled one")
                pad = (dd // 2, dh // 2, dw // 2)
                x1a = ActView(new_act(n, d1, h1, w1, c1, dev))
                _pack_ncdhw(x1, x1a)
                cat = new_act(n, d2, h2, w2, c2 + tp.cout, dev)
                _pack_ncdhw(x2, ActView(cat, 0, c2))
                upper = ActView(cat, c2, tp.cout)
                if dd or dh or dw:
                    ops.fill_zero(upper)
                ops.convt2x_fwd(x1a, tp.wf, tp.bias8, upper, pad)
                xin = ActView(cat)
                saved.update(x1a=x1a, pad=pad, c2=c2)
            nb, d, h, w, _ = xin.shape
            out = ActView(new_act(nb, d, h, w, dc.cout, dev))
            saved["st"] = dc.forward(xin, out, training)
            saved["xin"] = xin
            ctx.runner, ctx.saved, ctx.n_inputs = runner, saved, n_inputs
            ctx.param_list = params
            return out.to_ncdhw()

    @staticmethod
    def backward(ctx, dout):
        runner, sv = ctx.runner, ctx.saved
        if sv is None:
            raise RuntimeError("B200 block: backward called twice on the same forward; activations were freed")
        ctx.saved = None
        if sv["st"] is None:
            raise B200Error("B200 block: backward through an eval-mode forward is not supported (BatchNorm is folded "
                            "into the convolution epilogue); call .train() first or run under torch.no_grad()")
        m, kind = runner.module, runner.kind
        dc, tp = runner.dc, runner.tp
        dev = dout.device
        with torch.cuda.device(dev):
            grads = {}

            def g(p):
                if id(p) not in grads:
                    grads[id(p)] = torch.zeros(p.shape, device=dev, dtype=torch.float32)
                return grads[id(p)]

            xin = sv["xin"]
            n, d, h, w, cin_pad = xin.shape
            dov = ActView(new_act(n, d, h, w, dc.cout, dev))
            _pack_ncdhw(dout, dov)
            need_dx = any(ctx.needs_input_grad[2:2 + ctx.n_inputs])
            if need_dx and dc.p1.im2col:
                raise B200Error("input gradient of a thin-input (im2col'd) convolution is not available: detach the "
                                "input of this DoubleConv3D (the network input never needs a gradient)")
            dxin = ActView(new_act(n, d, h, w, cin_pad, dev)) if need_dx else None
            side = _Side(True, dev)
            dc.backward(sv["st"], dov, dxin, g, runner.scratch, side)
            dins = [None] * ctx.n_inputs
            if kind == "double":
                if need_dx:
                    dins[0] = dxin.to_ncdhw()[:, :sv["shapes"][0][1]]
            elif kind == "down":
                if need_dx:
                    xa = sv["pool_in"]
                    dxa = ActView(new_act(*xa.shape, dev))
                    ops.maxpool3d_bwd(xa, dxin, None, dxa)
                    dins[0] = dxa.to_ncdhw()[:, :sv["shapes"][0][1]]
            else:
                x1a, pad, c2 = sv["x1a"], sv["pad"], sv["c2"]
                # parameters of the transposed convolution always get their gradients: they need d(cat)
                if dxin is None:
                    raise B200Error("Up3D backward needs the gradient of the concatenated tensor")
                dupper = ActView(dxin.t, c2, tp.cout)
                _, d1, h1, w1, _ = x1a.shape
                if (2 * d1, 2 * h1, 2 * w1) == (d, h, w):
                    ops.channel_sum(dupper, g(m.up.bias))
                else:
                    ops.channel_sum_box(dupper, pad, (2 * d1, 2 * h1, 2 * w1), g(m.up.bias))
                side.run(lambda: ops.convt2x_wgrad(x1a, dupper, pad, g(m.up.weight)), keep=(x1a, dupper))
                if ctx.needs_input_grad[2]:
                    dx1 = ActView(new_act(*x1a.shape, dev))
                    ops.convt2x_dgrad(dupper, pad, tp.wd, dx1)
                    dins[0] = dx1.to_ncdhw()
                if ctx.needs_input_grad[3]:
                    dins[1] = ActView(dxin.t, 0, c2).to_ncdhw()
            side.join()
            pgrads = tuple(grads.get(id(p)) if p.requires_grad else None for p in ctx.param_list)
        return (None, None) + tuple(dins) + pgrads


def run_block(module, kind, *inputs):
    """forward of a stand-alone DoubleConv3D ('double'), Down3D ('down') or Up3D ('up') module"""
    runner = module.__dict__.get("_b200_runner")
    if runner is None:
        runner = _Runner(module, kind)
        object.__setattr__(module, "_b200_runner", runner)
    for t in inputs:
        if t.dim() != 5:
            raise ValueError(f"expected 5-D inputs (N, C, D, H, W), got shape {tuple(t.shape)}")
        if not t.is_cuda:
            raise B200Error(f"{type(module).__name__} (B200) runs on CUDA tensors only: there is no CPU path")
    params = list(module.parameters())
    if kind == "up" and torch.is_grad_enabled() and any(p.requires_grad for p in params):
        # the concat gradient is needed for the transposed convolution's own parameters
        inputs = tuple(t if t.requires_grad else t.detach().requires_grad_(True) for t in inputs)
    return _BlockFunction.apply(runner, len(inputs), *inputs, *params)
