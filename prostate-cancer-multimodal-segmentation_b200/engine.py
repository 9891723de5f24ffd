"""Execution engine of the B200 U-Net: a static schedule of C-ABI kernel launches for the forward, the backward and
the weight packing of one ``UNet3D`` (reference graph: models/unet3d.py:247-296; autograd of it: utils/trainer.py:191).

Data layout in HBM
  * activations: NDHWC bf16.  Every encoder level k owns one *concat buffer* (N, Dk, Hk, Wk, 2*Ck): the encoder's
    DoubleConv writes the skip into channels [0, Ck), the decoder's transposed conv writes the upsampled map into
    [Ck, 2*Ck) — ``F.pad`` + ``torch.cat`` (models/unet3d.py:143-156) never run as copies.
  * parameters: one flat fp32 buffer (all 82 tensors, reverse-forward order, nn.Parameter.data are views into it) and
    one flat fp32 gradient buffer of the same layout (Parameter.grad are views) so that the optimizer is one launch
    and data-parallel buckets are contiguous ranges that complete in order during backward.
  * packed bf16 weight shadows per conv: [27][Cout][Cin] (fprop, K-major B operand) and [27][Cin][Cout] (dgrad).
"""
import math

import torch

from . import ops
from ._lib import B200Error
from .ops import ActView, new_act

BN_EPS_DEFAULT = 1e-5


def _pad16(c: int) -> int:
    return (c + 15) // 16 * 16


def _is_conv3(p) -> bool:
    """3x3x3 conv weight that the engine stores physically as [27][Cout][Cin] (packed tap order kd, kw, kh)"""
    return p.dim() == 5 and tuple(p.shape[2:]) == (3, 3, 3) and p.shape[1] % 16 == 0


def _phys_strides(cout: int, cin: int):
    cc = cout * cin
    return (cin, 1, 9 * cc, cc, 3 * cc)  # logical (cout, cin, kd, kh, kw) over physical (kd, kw, kh, cout, cin)


def _slot_view(flat: torch.Tensor, off: int, p) -> torch.Tensor:
    """logical-shape view of flat[off : off + numel] for parameter p (permuted for physically packed conv weights)"""
    n = p.numel()
    if _is_conv3(p):
        cout, cin = p.shape[0], p.shape[1]
        return flat[off:off + n].view(3, 3, 3, cout, cin).permute(3, 4, 0, 2, 1)
    return flat[off:off + n].view(p.shape)


# A/B switch for tools/ab_fused.py only: False runs the unfused chains (bn_apply_relu + maxpool3d_fwd, head_bwd +
# bn_bwd) the fused BatchNorm passes replace; same results, more HBM traffic.
FUSE_BN_PASSES = True
# forward of the 5-modality first layer: depth-marching kernel (csrc/conv1_march.cu) instead of the generic direct one
USE_CONV1_MARCH = True

_SPLITK_WS = {}   # device -> list of scratch tensors, the last one is the largest


def _splitk_workspace(device, n, d, h, w, out_cols):
    """The fp32 scratch of the split-K form of a deep-level conv / dgrad (None when the shape does not split).  One buffer
    per device, grown on demand; all split-K launches run on the compute stream, so consecutive layers share it.  A
    buffer that has been handed out is never freed: a captured CUDA graph (graph.GraphedTrainStep) replays with the
    pointer it recorded."""
    nbytes = ops.conv3d_workspace_bytes(n, d, h, w, out_cols)
    if nbytes == 0:
        return None
    bufs = _SPLITK_WS.setdefault(device, [])
    if not bufs or bufs[-1].numel() * 4 < nbytes:
        if torch.cuda.is_current_stream_capturing():
            raise B200Error("split-K scratch must exist before CUDA-graph capture: run one eager step of this shape first")
        bufs.append(torch.empty((nbytes + 3) // 4, device=device, dtype=torch.float32))
    return bufs[-1]


class _ConvPack:
    """bf16 operand of one Conv3d 3x3x3.

    * Engine-managed weights live physically as fp32 [27][Cout][Cin]; their bf16 operand is a slice of the engine's
      bf16 shadow of the flat parameter buffer (written by the fused Adam, or by one cast kernel) — no packing kernel,
      and the weight gradient accumulates into the same contiguous layout (coalesced).
    * Stand-alone (torch-layout) weights are packed by b200_pack_conv_weight into a private buffer.
    * Thin inputs (Cin not a multiple of 16, i.e. the 5-modality first layer) use the im2col form: the input is
      expanded to rows of pad16(27*Cin) columns and the conv runs as a 1-tap GEMM on the flattened weights."""

    def __init__(self, conv, device):
        self.conv = conv
        self.device = device
        self.cout, self.cin = conv.out_channels, conv.in_channels
        self.im2col = self.cin % 16 != 0 and 27 * self.cin <= 512
        # the 5-modality first layer reads the fp32 network input directly: its im2col rows are built in shared
        # memory inside the GEMM kernels (ops.conv1_direct_*), never in HBM
        self.direct = self.im2col and ops.conv1_direct_supported(self.cin, self.cout, 4)
        # ... and its forward marches along depth when the layer is narrow enough (csrc/conv1_march.cu)
        self.march = USE_CONV1_MARCH and self.direct and ops.conv1_march_supported(self.cin, self.cout)
        self._slices = (torch.empty(3, self.cout, 64, device=device, dtype=torch.bfloat16) if self.march else None)
        self.shadow = None   # set by the engine: bf16 [27][Cout][Cin] view of its parameter shadow
        self._own = None
        # input gradients of 64 / 32 channels run on the CTA-pair depth-marching kernel, which reads the weights K-major:
        # a transposed copy [27][Cin][Cout], refreshed with the other packs (three small layers of the network)
        self._wt = (torch.empty(27, self.cin, self.cout, device=device, dtype=torch.bfloat16)
                    if (not self.im2col and self.cin in (32, 64) and self.cin == _pad16(self.cin)) else None)
        if self.im2col:
            self.k_real = 27 * self.cin
            self.cin_pad = _pad16(self.k_real)  # width of the im2col rows
            self._own = torch.empty(self.cout, self.cin_pad, device=device, dtype=torch.bfloat16)
        else:
            self.k_real = self.cin
            self.cin_pad = _pad16(self.cin)

    def _phys(self, t) -> bool:
        return (not self.im2col) and self.shadow is not None and t.stride() == _phys_strides(self.cout, self.cin)

    @property
    def wf(self):
        if self._phys(self.conv.weight.data):
            return self.shadow
        if self._own is None:
            self._own = torch.empty(27, self.cout, self.cin_pad, device=self.device, dtype=torch.bfloat16)
        return self._own

    def pack(self):
        if self.im2col:
            ops.pack_rows(self.conv.weight.data, self.cin_pad, self._own)
            if self.march:
                ops.pack_conv1_slices(self.conv.weight.data.contiguous(), self._slices)
        elif not self._phys(self.conv.weight.data):
            ops.pack_conv_weight(self.conv.weight.data.contiguous(), self.cin_pad, self.wf)
        if self._wt is not None:
            ops.transpose_taps(self.wf, self._wt)

    def make_input(self, x: torch.Tensor):
        """fp32 (N,C,D,H,W) -> the operand this conv reads: channel-padded NDHWC bf16, im2col rows, or (direct first
        layer) the fp32 tensor itself"""
        if self.direct and ops.conv1_direct_supported(self.cin, self.cout, x.shape[-1]):
            return RawInput(x)
        n, _, d, h, w = x.shape
        v = ActView(new_act(n, d, h, w, self.cin_pad, x.device))
        (ops.im2col_input if self.im2col else ops.pack_input)(x, v)
        return v

    def fprop(self, xin, bias, y, stats, mode, scale=None, shift=None, workspace=None):
        if isinstance(xin, RawInput) and self.march:
            ops.conv1_march_fprop(xin.t, self._slices, bias, y, stats, mode, scale, shift)
        elif isinstance(xin, RawInput):
            ops.conv1_direct_fprop(xin.t, self.wf, bias, y, stats, mode, scale, shift)
        elif self.im2col:
            ops.conv1_fprop(xin, self.wf, bias, y, stats, mode, scale, shift, k_real=self.k_real)
        else:
            ops.conv3d_fprop(xin, self.wf, bias, y, stats, mode, scale, shift, k_real=self.k_real,
                             workspace=workspace)

    def wgrad(self, xin, dy, dw):
        if isinstance(xin, RawInput) and self.march:
            ops.conv1_march_wgrad(xin.t, dy, dw.view(self.cout, -1))
        elif isinstance(xin, RawInput):
            ops.conv1_direct_wgrad(xin.t, dy, dw.view(self.cout, -1))
        elif self.im2col:
            ops.conv1_wgrad(xin, dy, dw.view(self.cout, -1), self.k_real)
        elif dw.stride() == _phys_strides(self.cout, self.cin):
            ops.conv3d_wgrad(xin, dy, dw, self.cin, packed=True)
        elif dw.is_contiguous():
            ops.conv3d_wgrad(xin, dy, dw, self.cin, packed=False)
        else:
            raise B200Error("conv weight gradient buffer has an unsupported memory layout")

    def dgrad(self, dy, dx, workspace=None):
        if self.im2col:
            raise B200Error("input gradient of an im2col'd (thin-input) convolution is not available")
        n, d, h, w, _ = dy.shape
        if self._wt is not None and ops.conv3d_dgrad_kmajor_supported(n, d, h, w, dx.c):
            ops.conv3d_dgrad_kmajor(dy, self._wt, dx)
        else:
            ops.conv3d_dgrad(dy, self.wf, dx, workspace=workspace)


class RawInput:
    """the fp32 (N, C, D, H, W) network input standing where an ActView would (direct first layer)"""
    __slots__ = ("t",)

    def __init__(self, t: torch.Tensor):
        self.t = t

    @property
    def shape(self):
        n, c, d, h, w = self.t.shape
        return (n, d, h, w, c)


class _ConvTPack:
    def __init__(self, up, device):
        self.up = up
        self.cin, self.cout = up.in_channels, up.out_channels
        self.wf = torch.empty(8 * self.cout, self.cin, device=device, dtype=torch.bfloat16)
        self.wd = torch.empty(8, self.cin, self.cout, device=device, dtype=torch.bfloat16)
        self.bias8 = torch.empty(8 * self.cout, device=device, dtype=torch.float32)

    def pack(self):
        ops.pack_convt_weight(self.up.weight.data, self.up.bias.data, self.wf, self.wd, self.bias8)


class _DCState:
    """what one DoubleConv3D saves for its backward"""
    __slots__ = ("xin", "y1", "a1", "y2", "bn1", "bn2")


class _DoubleConv:
    """Conv3d -> BatchNorm3d -> ReLU, twice (models/unet3d.py:27-40)"""

    def __init__(self, seq, device):
        self.conv1, self.bn1, self.conv2, self.bn2 = seq[0], seq[1], seq[3], seq[4]
        self.p1 = _ConvPack(self.conv1, device)
        self.p2 = _ConvPack(self.conv2, device)
        self.cout = self.conv1.out_channels

    def pack(self):
        self.p1.pack()
        self.p2.pack()

    # ---- forward
    def _conv_bn_relu_train(self, xin: ActView, pack, bn, out: ActView, pool_out=None):
        n, d, h, w, _ = xin.shape
        dev = xin.t.device
        cout = pack.cout
        if n * d * h * w <= 1:  # same contract as torch.nn.functional.batch_norm in training mode
            raise ValueError(f"Expected more than 1 value per channel when training, got input size "
                             f"{(n, cout, d, h, w)}")
        y = ActView(new_act(n, d, h, w, cout, dev))
        ws = None if pack.im2col else _splitk_workspace(dev, n, d, h, w, cout)
        rows = ((ops.conv1_march_stat_rows if pack.march else ops.conv1_direct_stat_rows)(n, d, h, w, cout)
                if isinstance(xin, RawInput)
                else ops.conv3d_stat_rows(n, d, h, w, cout, 1 if pack.im2col else 27, with_workspace=ws is not None))
        stats = torch.empty(rows, cout, 2, device=dev, dtype=torch.float32)
        pack.fprop(xin, pack.conv.bias.data, y, stats, ops.EPI_BIAS_STATS, workspace=ws)
        vec = torch.empty(4, cout, device=dev, dtype=torch.float32)  # mean, rstd, scale, shift
        momentum = bn.momentum
        nbt = None
        if bn.track_running_stats:
            if momentum is None:  # cumulative moving average, as torch: needs the count on the host
                bn.num_batches_tracked.add_(1)
                momentum = 1.0 / float(bn.num_batches_tracked.item())
            else:
                nbt = bn.num_batches_tracked  # incremented inside the finalize launch
        ops.bn_finalize(stats, rows, n * d * h * w, cout, bn.weight.data, bn.bias.data, bn.eps,
                        momentum if momentum is not None else 0.0,
                        bn.running_mean if bn.track_running_stats else None,
                        bn.running_var if bn.track_running_stats else None, vec[0], vec[1], vec[2], vec[3],
                        num_batches_tracked=nbt)
        if pool_out is not None and FUSE_BN_PASSES:   # encoder block: the next level's MaxPool3d(2) in the same pass
            ops.bn_apply_relu_pool(y, vec[2], vec[3], out, pool_out)
        else:
            ops.bn_apply_relu(y, vec[2], vec[3], out)
            if pool_out is not None:
                ops.maxpool3d_fwd(out, pool_out)
        return y, vec

    def _conv_bn_relu_eval(self, xin: ActView, pack, bn, out: ActView):
        dev = xin.t.device
        vec = torch.empty(2, pack.cout, device=dev, dtype=torch.float32)
        ops.bn_fold_eval(bn.weight.data, bn.bias.data, bn.running_mean, bn.running_var, pack.conv.bias.data, bn.eps,
                         vec[0], vec[1])
        n, d, h, w, _ = out.shape
        ws = None if pack.im2col else _splitk_workspace(dev, n, d, h, w, pack.cout)
        pack.fprop(xin, None, out, None, ops.EPI_AFFINE_RELU, vec[0], vec[1], workspace=ws)

    def forward(self, xin: ActView, out: ActView, training: bool, pool_out=None):
        """pool_out: view that receives MaxPool3d(2)(out) (models/unet3d.py:80, the next level's input)"""
        n, d, h, w, _ = xin.shape
        dev = xin.t.device
        a1 = ActView(new_act(n, d, h, w, self.cout, dev))
        if not (training or not self.bn1.track_running_stats):
            self._conv_bn_relu_eval(xin, self.p1, self.bn1, a1)
            self._conv_bn_relu_eval(a1, self.p2, self.bn2, out)
            if pool_out is not None:
                ops.maxpool3d_fwd(out, pool_out)
            return None
        st = _DCState()
        st.xin, st.a1 = xin, a1
        st.y1, st.bn1 = self._conv_bn_relu_train(xin, self.p1, self.bn1, a1)
        st.y2, st.bn2 = self._conv_bn_relu_train(a1, self.p2, self.bn2, out, pool_out)
        return st

    # ---- backward
    def backward(self, st: _DCState, dout: ActView, dxin, grads, scratch, side, taps=None, name=""):
        """dout: gradient w.r.t. the block output; dxin: view to receive the input gradient (None: not needed).
        Weight gradients are launched through `side` (runs them on the side stream, see Engine._Side).
        taps (parity tooling): dict that receives every intermediate gradient view of the block."""
        n, d, h, w, _ = st.y2.shape
        dev = st.y2.t.device
        g = grads
        # second conv
        dy2 = ActView(new_act(n, d, h, w, self.cout, dev))
        bn2 = (st.y2, st.bn2[2], st.bn2[3], st.bn2[0], st.bn2[1], self.bn2.weight.data, scratch.partial, scratch.coef,
               g(self.bn2.weight), g(self.bn2.bias), dy2, g(self.conv2.bias))
        if isinstance(dout, HeadGrad) and not FUSE_BN_PASSES:
            t = ActView(new_act(n, d, h, w, self.cout, dev))
            ops.head_bwd(dout.act, dout.w, dout.dlogits, t, dout.dw, dout.db)
            dout = t
        if isinstance(dout, HeadGrad):   # dout = dlogits . w_head, never written; head dw / db on the way
            ops.bn_bwd_head(dout.dlogits, dout.w, *bn2, dout.dw, dout.db)
        else:
            ops.bn_bwd(dout, *bn2)
        # dgrad first: it is on the critical chain (the next BatchNorm backward needs da1) and must win the SMs; the
        # weight gradient then runs beside that BatchNorm backward
        da1 = ActView(new_act(n, d, h, w, self.cout, dev))
        self.p2.dgrad(dy2, da1, workspace=_splitk_workspace(dev, n, d, h, w, da1.c))
        side.run(lambda: self.p2.wgrad(st.a1, dy2, g(self.conv2.weight)), keep=(st.a1, dy2))
        st.y2 = None
        # first conv
        dy1 = ActView(new_act(n, d, h, w, self.cout, dev))
        ops.bn_bwd(da1, st.y1, st.bn1[2], st.bn1[3], st.bn1[0], st.bn1[1], self.bn1.weight.data, scratch.partial,
                   scratch.coef, g(self.bn1.weight), g(self.bn1.bias), dy1, g(self.conv1.bias))
        if dxin is not None:
            self.p1.dgrad(dy1, dxin, workspace=_splitk_workspace(dev, n, d, h, w, dxin.c))
        side.run(lambda: self.p1.wgrad(st.xin, dy1, g(self.conv1.weight)), keep=(st.xin, dy1))
        if taps is not None:
            taps[name] = {"dout": dout if isinstance(dout, ActView) else dout.materialized, "dy2": dy2, "da1": da1,
                          "dy1": dy1, "dx": dxin}


class HeadGrad:
    """Gradient w.r.t. the last block's output that is not materialised: dlogits (N, ncls, D, H, W) . w (ncls, C); dw /
    db are the head's gradient buffers (accumulated by the BatchNorm-backward reduce pass)."""
    __slots__ = ("dlogits", "w", "dw", "db", "act", "materialized")

    def __init__(self, dlogits, w, dw, db, act=None, materialized=None):
        self.dlogits, self.w, self.dw, self.db, self.act, self.materialized = dlogits, w, dw, db, act, materialized


class _Side:
    """Weight-gradient GEMMs run on a second stream: they hang off the backward chain (nothing downstream reads them
    before the optimizer), so they overlap with the HBM-bound BatchNorm-backward / pool-backward kernels of the next
    layer, which fit beside a persistent GEMM CTA on every SM.  Buffers they read are kept alive until join()."""

    def __init__(self, enabled=True, device=None):
        self.enabled = enabled
        self.stream = torch.cuda.Stream(device) if enabled else None
        self.keep = []

    def run(self, fn, keep=()):
        if not self.enabled:
            fn()
            return
        main = torch.cuda.current_stream()
        self.stream.wait_event(main.record_event())
        with torch.cuda.stream(self.stream):
            fn()
        self.keep.extend(keep)

    def sync_point(self, fn):
        """run fn (data-parallel bucket launch) where every gradient produced so far on either stream is visible"""
        if not self.enabled:
            fn()
            return
        main = torch.cuda.current_stream()
        self.stream.wait_event(main.record_event())
        with torch.cuda.stream(self.stream):
            fn()

    def join(self):
        if self.enabled:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.keep = []


class _Scratch:
    def __init__(self, device, cmax):
        self.partial = torch.empty(ops.bn_bwd_max_blocks(), cmax, 2, device=device, dtype=torch.float32)
        self.coef = torch.empty(cmax, 2, device=device, dtype=torch.float32)


class _Tape:
    """activations kept between forward and backward of one step"""
    __slots__ = ("dims", "pads", "cats", "dcs", "pooled", "dec_in", "last", "x_shape")


class Engine:
    def __init__(self, model):
        self.model = model
        self.device = None
        self._pack_key = None
        self.external_epoch = 0  # bumped by anything that writes parameter memory behind torch's back (FusedAdam)
        self.grad_sync = None    # optional data-parallel hook: object with .ready(hi) and .finish()
        self.overlap_wgrad = True  # weight gradients on a side stream (see _Side)
        self.flat_param = None
        self.flat_grad = None
        self.flat_bf16 = None    # bf16 shadow of flat_param: the conv kernels' weight operand
        self._shadow_key = None
        self._slots = None       # id(param) -> (offset, numel)
        self.keep_tape = False   # parity tooling: keep the last forward's tape in .last_tape (see layer_outputs)
        self.last_tape = None
        self.grad_taps = None    # parity tooling: dict that receives every intermediate gradient of the next backward

    # ------------------------------------------------------------------ parameters
    def ordered_params(self):
        """(name, parameter) in reverse-forward order: the order in which backward completes their gradients"""
        m = self.model
        out = [("outc.weight", m.outc.weight), ("outc.bias", m.outc.bias)]

        def dc(prefix, seq):
            return [(f"{prefix}.4.weight", seq[4].weight), (f"{prefix}.4.bias", seq[4].bias),
                    (f"{prefix}.3.weight", seq[3].weight), (f"{prefix}.3.bias", seq[3].bias),
                    (f"{prefix}.1.weight", seq[1].weight), (f"{prefix}.1.bias", seq[1].bias),
                    (f"{prefix}.0.weight", seq[0].weight), (f"{prefix}.0.bias", seq[0].bias)]

        for j in (4, 3, 2, 1):
            up = getattr(m, f"up{j}")
            out += dc(f"up{j}.conv.conv", up.conv.conv)
            out += [(f"up{j}.up.weight", up.up.weight), (f"up{j}.up.bias", up.up.bias)]
        for k in (4, 3, 2, 1):
            out += dc(f"down{k}.maxpool_conv.1.conv", getattr(m, f"down{k}").maxpool_conv[1].conv)
        out += dc("inc.conv", m.inc.conv)
        return out

    def _is_flat(self) -> bool:
        if self.flat_param is None:
            return False
        base = self.flat_param.data_ptr()
        for _, p in self.ordered_params():
            off, n = self._slots[id(p)]
            d = p.data
            if d.data_ptr() != base + 4 * off or d.numel() != n or d.dtype != torch.float32:
                return False
            if _is_conv3(p) and d.stride() != _phys_strides(p.shape[0], p.shape[1]):
                return False
        return True

    def flatten(self, device):
        """Re-home every parameter into one flat fp32 buffer (values preserved), allocate the flat gradient and the
        bf16 operand shadow.  3x3x3 conv weights are stored physically as [27][Cout][Cin]; `Parameter.data` is the
        permuted (Cout, Cin, 3, 3, 3) view, so state_dict / load_state_dict / any optimizer see torch's shape."""
        params = self.ordered_params()
        slots, off = {}, 0
        for _, p in params:
            slots[id(p)] = (off, p.numel())
            off += (p.numel() + 63) // 64 * 64   # 128-byte aligned bf16 shadow slices (TMA rows must not straddle lines)
        flat = torch.zeros(off, device=device, dtype=torch.float32)
        for _, p in params:
            o, _n = slots[id(p)]
            view = _slot_view(flat, o, p)
            view.copy_(p.data)
            p.data = view
            p.grad = None
        self._slots = slots
        self.flat_param = flat
        self.flat_grad = torch.zeros(off, device=device, dtype=torch.float32)
        self.flat_bf16 = torch.empty(off, device=device, dtype=torch.bfloat16)
        self._pack_key = None
        self._shadow_key = None
        for dc in [self.inc] + self.downs + [d for _, d in self.ups]:
            for pk in (dc.p1, dc.p2):
                if not pk.im2col:
                    o, n = slots[id(pk.conv.weight)]
                    pk.shadow = self.flat_bf16[o:o + n].view(27, pk.cout, pk.cin)

    def grad_view(self, p):
        o, _n = self._slots[id(p)]
        return _slot_view(self.flat_grad, o, p)

    def current_key(self):
        return (self.external_epoch, self.flat_param.data_ptr(), tuple(p._version for _, p in self.ordered_params()))

    def _build(self, device):
        m = self.model
        self.device = device
        self.inc = _DoubleConv(m.inc.conv, device)
        self.downs = [_DoubleConv(getattr(m, f"down{k}").maxpool_conv[1].conv, device) for k in (1, 2, 3, 4)]
        self.ups = [(_ConvTPack(getattr(m, f"up{j}").up, device), _DoubleConv(getattr(m, f"up{j}").conv.conv, device))
                    for j in (1, 2, 3, 4)]
        cmax = max([self.inc.cout] + [d.cout for d in self.downs])
        self.scratch = _Scratch(device, cmax)
        self._pack_key = None

    def prepare(self, device):
        if device.type != "cuda":
            raise B200Error("UNet3D (B200) runs on CUDA tensors only: there is no CPU path. "
                            "Move the model and its input to a cuda device.")
        with torch.cuda.device(device):
            self._prepare(device)

    def _prepare(self, device):
        if self.device != device:
            self._build(device)
        if not self._is_flat():
            self.flatten(device)
        key = self.current_key()
        if key != self._pack_key:
            if key != self._shadow_key:  # the fused Adam refreshes the shadow itself
                ops.cast_bf16(self.flat_param, self.flat_bf16)
                self._shadow_key = key
            self.inc.pack()
            for d in self.downs:
                d.pack()
            for t, d in self.ups:
                t.pack()
                d.pack()
            self._pack_key = key

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, training: bool, want_probs: bool = False):
        """logits (+ probabilities) and the tape of one forward; runs on x's device whatever the current device is"""
        if x.is_cuda:
            with torch.cuda.device(x.device):
                return self._forward(x, training, want_probs)
        return self._forward(x, training, want_probs)

    def _forward(self, x: torch.Tensor, training: bool, want_probs: bool = False):
        m = self.model
        if x.dim() != 5:
            raise ValueError(f"expected a 5-D input (N, C, D, H, W), got shape {tuple(x.shape)}")
        n, c, D, H, W = x.shape
        if c != m.n_modalities:
            raise RuntimeError(f"expected input with {m.n_modalities} channels, got {c} (shape {tuple(x.shape)})")
        if min(D, H, W) < 16:
            raise RuntimeError(f"every spatial extent must be >= 16 (four 2x poolings), got {(D, H, W)}")
        dev = x.device
        self.prepare(dev)
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        f = m.init_features
        ch = [f, 2 * f, 4 * f, 8 * f, 16 * f]
        dims = [(D, H, W)]
        for _ in range(4):
            d, h, w = dims[-1]
            dims.append((d // 2, h // 2, w // 2))

        tape = _Tape()
        tape.dims, tape.x_shape = dims, tuple(x.shape)
        x0 = self.inc.p1.make_input(x)
        # encoder: skip of level k lives in the lower half of cats[k]
        cats = [new_act(n, *dims[k], 2 * ch[k], dev) for k in range(4)]
        dcs = {}
        # (the MaxPool3d(2) input of level k + 1 is produced by level k's last BatchNorm + ReLU pass)
        pooled = [ActView(new_act(n, *dims[k], ch[k - 1], dev)) for k in range(1, 5)]
        dcs["inc"] = self.inc.forward(x0, ActView(cats[0], 0, ch[0]), training, pool_out=pooled[0])
        bottom = None
        for k in range(1, 5):
            if k < 4:
                out = ActView(cats[k], 0, ch[k])
            else:
                bottom = out = ActView(new_act(n, *dims[4], ch[4], dev))
            dcs[f"down{k}"] = self.downs[k - 1].forward(pooled[k - 1], out, training,
                                                        pool_out=pooled[k] if k < 4 else None)
        # decoder
        cur = bottom
        dec_in, pads = [], []
        for j in range(1, 5):
            k = 4 - j
            tp, dc = self.ups[j - 1]
            (d2, h2, w2), (d1, h1, w1) = dims[k], cur.shape[1:4]
            dd, dh, dw = d2 - 2 * d1, h2 - 2 * h1, w2 - 2 * w1
            pad = (dd // 2, dh // 2, dw // 2)
            upper = ActView(cats[k], ch[k], ch[k])
            if dd or dh or dw:
                ops.fill_zero(upper)  # F.pad border (models/unet3d.py:149-151)
            ops.convt2x_fwd(cur, tp.wf, tp.bias8, upper, pad)
            dec_in.append(cur)
            pads.append(pad)
            out = ActView(new_act(n, *dims[k], ch[k], dev))
            dcs[f"up{j}"] = dc.forward(ActView(cats[k]), out, training)
            cur = out
        logits = torch.empty(n, m.n_classes, D, H, W, device=dev, dtype=torch.float32)
        probs = torch.empty_like(logits) if want_probs else None
        ops.head_fwd(cur, m.outc.weight.data.view(m.n_classes, -1), m.outc.bias.data, logits, probs)
        tape.cats, tape.dcs, tape.pooled, tape.dec_in, tape.pads, tape.last = cats, dcs, pooled, dec_in, pads, cur
        if self.keep_tape:
            self.last_tape = tape
        return logits, probs, tape

    def layer_outputs(self, tape: _Tape):
        """(reference module path, ActView) of every tensor a training-mode forward keeps: raw conv outputs at
        '<block>.0' / '<block>.3', post-ReLU activations at '<block>.2' / '<block>.5', transposed-conv outputs at
        'upJ.up' — the keys the oracle's `taps` use (parity tests compare layer by layer).  Call before backward
        (backward releases the tape as it goes)."""
        f = self.model.init_features
        ch = [f, 2 * f, 4 * f, 8 * f, 16 * f]
        outs = {"inc": ActView(tape.cats[0], 0, ch[0])}
        for k in (1, 2, 3):
            outs[f"down{k}"] = ActView(tape.cats[k], 0, ch[k])
        outs["down4"] = tape.dec_in[0]
        for j in (1, 2, 3):
            outs[f"up{j}"] = tape.dec_in[j]
        outs["up4"] = tape.last
        prefix = {"inc": "inc.conv"}
        for k in (1, 2, 3, 4):
            prefix[f"down{k}"] = f"down{k}.maxpool_conv.1.conv"
        for j in (1, 2, 3, 4):
            prefix[f"up{j}"] = f"up{j}.conv.conv"
        for name in ["inc", "down1", "down2", "down3", "down4", "up1", "up2", "up3", "up4"]:
            st = tape.dcs[name]
            if name.startswith("up"):
                k = 4 - int(name[2])
                yield f"{name}.up", ActView(tape.cats[k], ch[k], ch[k])
            if st is None:
                continue
            if name != "inc":   # (the first block reads im2col rows of the network input)
                yield f"{prefix[name]}.in", st.xin
            yield f"{prefix[name]}.0", st.y1
            yield f"{prefix[name]}.2", st.a1
            yield f"{prefix[name]}.3", st.y2
            yield f"{prefix[name]}.5", outs[name]

    # ------------------------------------------------------------------ backward
    def _begin_grads(self):
        """Parameter.grad protocol: grads accumulate into views of the flat gradient buffer.
        Returns g(param) -> fp32 tensor that kernels accumulate into, and a list of foreign grads to fold back."""
        params = self.ordered_params()
        ours, none_count = {}, 0
        for _, p in params:
            if p.grad is None:
                none_count += 1
        foreign = []
        if none_count == len(params):
            self.flat_grad.zero_()
        for _, p in params:
            gv = self.grad_view(p)
            if p.grad is None:
                if none_count != len(params):
                    gv.zero_()
                p.grad = gv
                ours[id(p)] = gv
            elif p.grad.data_ptr() == gv.data_ptr() and p.grad.dtype == torch.float32:
                ours[id(p)] = p.grad  # accumulate in place (zero_grad(set_to_none=False) or gradient accumulation)
            else:
                tmp = torch.zeros_like(gv)
                ours[id(p)] = tmp
                foreign.append((p, tmp))
        return (lambda p: ours[id(p)]), foreign

    def backward(self, tape: _Tape, dlogits: torch.Tensor):
        with torch.cuda.device(dlogits.device):
            return self._backward(tape, dlogits)

    def _backward(self, tape: _Tape, dlogits: torch.Tensor):
        m = self.model
        dev = dlogits.device
        n = tape.x_shape[0]
        dims = tape.dims
        f = m.init_features
        ch = [f, 2 * f, 4 * f, 8 * f, 16 * f]
        g, foreign = self._begin_grads()
        sync = self.grad_sync
        gt = self.grad_taps
        side = _Side(self.overlap_wgrad, dev)
        if dlogits.dtype != torch.float32 or not dlogits.is_contiguous():
            dlogits = dlogits.float().contiguous()

        def mark(p):
            if sync is not None:
                o, cnt = self._slots[id(p)]
                side.sync_point(lambda: sync.ready(o + cnt))

        # the head's input gradient dlogits . w is not written: the last block's BatchNorm backward recomputes it and
        # accumulates the head's dw / db (ops.bn_bwd_head); their bucket is marked with that block's
        w_head = m.outc.weight.data.view(m.n_classes, -1)
        dcur = HeadGrad(dlogits, w_head, g(m.outc.weight).view(m.n_classes, -1), g(m.outc.bias),
                        act=None if FUSE_BN_PASSES else tape.last)
        if gt is not None:   # parity tooling: the tensor b200_head_bwd writes for the same inputs
            dcur.materialized = ActView(new_act(n, *dims[0], ch[0], dev))
            ops.head_bwd(tape.last, w_head, dlogits, dcur.materialized, torch.zeros_like(w_head),
                         torch.zeros_like(m.outc.bias.data))
            gt["head"] = {"dx": dcur.materialized}
        tape.last = None
        dcats = [None] * 4
        for j in (4, 3, 2, 1):
            k = 4 - j
            tp, dc = self.ups[j - 1]
            dcat = new_act(n, *dims[k], 2 * ch[k], dev)
            dcats[k] = dcat
            dc.backward(tape.dcs[f"up{j}"], dcur, ActView(dcat), g, self.scratch, side, gt, f"up{j}")
            tape.dcs[f"up{j}"] = None
            mark(dc.conv1.bias)
            dupper = ActView(dcat, ch[k], ch[k])
            x_in = tape.dec_in[j - 1]
            pad = tape.pads[j - 1]
            d1, h1, w1 = x_in.shape[1:4]
            if (2 * d1, 2 * h1, 2 * w1) == dims[k]:
                ops.channel_sum(dupper, g(tp.up.bias))
            else:  # F.pad border carries no bias gradient: reduce the un-padded core only
                ops.channel_sum_box(dupper, pad, (2 * d1, 2 * h1, 2 * w1), g(tp.up.bias))
            dprev = ActView(new_act(*x_in.shape, dev))
            ops.convt2x_dgrad(dupper, pad, tp.wd, dprev)
            side.run(lambda x_in=x_in, dupper=dupper, pad=pad, tp=tp: ops.convt2x_wgrad(x_in, dupper, pad,
                                                                                      g(tp.up.weight)),
                     keep=(x_in, dupper, dcat))
            mark(tp.up.bias)
            if gt is not None:
                gt[f"up{j}.up"] = {"dout": dupper, "dx": dprev, "x": x_in, "pad": pad}
            dcur = dprev
        tape.dec_in = None
        for k in (4, 3, 2, 1):
            dcobj = self.downs[k - 1]
            dpool = ActView(new_act(n, *dims[k], ch[k - 1], dev))
            dcobj.backward(tape.dcs[f"down{k}"], dcur, dpool, g, self.scratch, side, gt, f"down{k}")
            tape.dcs[f"down{k}"] = None
            mark(dcobj.conv1.bias)
            skip_act = ActView(tape.cats[k - 1], 0, ch[k - 1])
            dskip = ActView(dcats[k - 1], 0, ch[k - 1])
            if gt is not None:
                gt[f"pool{k}"] = {"x": skip_act, "dy": dpool, "dskip_before": dskip.to_ncdhw(), "dx": dskip}
            ops.maxpool3d_bwd(skip_act, dpool, dskip, dskip)  # in place: dskip += scatter(dpool)
            dcur = dskip
        self.inc.backward(tape.dcs["inc"], dcur, None, g, self.scratch, side, gt, "inc")
        mark(self.inc.conv1.bias)
        tape.dcs = tape.cats = tape.pooled = None
        if sync is not None:
            side.sync_point(sync.finish)
        side.join()
        for p, tmp in foreign:
            p.grad.add_(tmp.to(p.grad.dtype))


def total_flops_per_voxel(init_features: int = 64, n_modalities: int = 5, n_classes: int = 1):
    """algorithmic conv FLOPs per input voxel: (forward, forward+backward); SURVEY.md 8(d) accounting"""
    f = init_features
    fwd = 0.0
    first = 2.0 * 27 * n_modalities * f
    fwd += first + 2.0 * 27 * f * f
    c = f
    for k in range(1, 5):
        s = 8.0 ** -k
        fwd += s * 2.0 * 27 * (c * 2 * c + 2 * c * 2 * c)
        c *= 2
    for j in range(1, 5):
        k = 4 - j
        s = 8.0 ** -k
        fwd += (8.0 ** -(k + 1)) * 2.0 * c * (c // 2) * 8   # transposed conv, per coarse voxel
        fwd += s * 2.0 * 27 * (c * (c // 2) + (c // 2) * (c // 2))
        c //= 2
    fwd += 2.0 * f * n_classes
    return fwd, 3.0 * fwd - first
