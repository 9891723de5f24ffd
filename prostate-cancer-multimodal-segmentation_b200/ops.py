"""Python faces of the C-ABI entry points (include/b200_unet3d.h), one function per entry point.

Tensors are only carriers of device memory here: every function takes torch tensors, passes their device pointers
and the current CUDA stream to the sm_100a kernels and returns nothing (outputs are caller-allocated).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import Act, check, ptr, stream_ptr

EPI_PLAIN, EPI_BIAS_STATS, EPI_AFFINE_RELU, EPI_BIAS = 0, 1, 2, 3

# ---- instrumentation (bench.py): kernel launch counter and an optional per-launch timing hook for the GEMM kernels
launch_count = 0
profile_hook = None  # callable(kernel_name, tag, algorithmic_flops, start_event, end_event, shape)
#                      shape = (voxels, in channels, out channels) of the GEMM: identifies the layer


def _launched(n: int = 1):
    global launch_count
    launch_count += n


def _gemm(kernel, tag, flops, call, shape=None):
    _launched(1)
    if profile_hook is None:
        return call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call()
    e1.record()
    profile_hook(kernel, tag, flops, e0, e1, shape)


class ActView:
    """Channel slice [c_off, c_off + c) of an NDHWC bf16 buffer of shape (N, D, H, W, LD)."""

    __slots__ = ("t", "c_off", "c", "_s")

    def __init__(self, t: torch.Tensor, c_off: int = 0, c: int = None):
        if t.dtype != torch.bfloat16 or t.dim() != 5 or not t.is_contiguous():
            raise ValueError("ActView needs a contiguous bf16 (N,D,H,W,C) tensor")
        if not t.is_cuda:
            raise _lib.B200Error("b200 kernels need CUDA tensors: there is no CPU path")
        self.t = t
        self.c_off = c_off
        self.c = t.shape[4] - c_off if c is None else c
        n, d, h, w, ld = t.shape
        self._s = Act(t.data_ptr() + 2 * c_off, n, d, h, w, self.c, ld)

    @property
    def ref(self):
        return C.byref(self._s)

    @property
    def shape(self):
        n, d, h, w, _ = self.t.shape
        return (n, d, h, w, self.c)

    @property
    def voxels(self) -> int:
        n, d, h, w, _ = self.t.shape
        return n * d * h * w

    def as_torch(self):
        """(N,D,H,W,c) strided torch view (debug / tests)"""
        return self.t[..., self.c_off:self.c_off + self.c]

    def to_ncdhw(self) -> torch.Tensor:
        n, d, h, w, c = self.shape
        out = torch.empty((n, c, d, h, w), device=self.t.device, dtype=torch.float32)
        with torch.cuda.device(self.t.device):
            check(_lib.load().b200_unpack_act(self.ref, ptr(out), stream_ptr()), "unpack_act")
        return out


def new_act(n, d, h, w, c, device, zero=False) -> torch.Tensor:
    f = torch.zeros if zero else torch.empty
    return f((n, d, h, w, c), device=device, dtype=torch.bfloat16)


def sm_count() -> int:
    return _lib.load().b200_sm_count()


def set_pdl(on: bool) -> bool:
    """process-wide switch of programmatic dependent launch (csrc/launch.cuh; default on); returns the previous value.
    A CUDA graph keeps the kind of edges it was captured with."""
    return bool(_lib.load().b200_set_pdl(1 if on else 0))


def set_dmarch_pair_mma(on: bool) -> bool:
    """process-wide switch: depth-marching convolutions on cta_group::2 MMAs (csrc/dmarch2.cu; default on)"""
    return bool(_lib.load().b200_set_dmarch_pair_mma(1 if on else 0))


def pack_input(x: torch.Tensor, out: ActView):
    _launched(1)
    n, c, d, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    check(_lib.load().b200_pack_input(ptr(x), n, c, d, h, w, out.ref, stream_ptr()), "pack_input")


def im2col_input(x: torch.Tensor, out: ActView):
    _launched(1)
    n, c, d, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    check(_lib.load().b200_im2col_input(ptr(x), n, c, d, h, w, out.ref, stream_ptr()), "im2col_input")


def pack_rows(w: torch.Tensor, k_pad: int, out):
    _launched(1)
    rows = w.shape[0]
    k = w.numel() // rows
    assert w.dtype == torch.float32 and w.is_contiguous()
    check(_lib.load().b200_pack_rows(ptr(w), rows, k, k_pad, ptr(out), stream_ptr()), "pack_rows")


def conv1_fprop(x: ActView, w_rows, bias, y: ActView, stats=None, mode=EPI_BIAS_STATS, scale=None, shift=None,
                k_real=None):
    lib = _lib.load()
    _gemm("igemm_kernel", "conv1_fprop", 2.0 * x.voxels * y.c * (k_real or x.c),
          lambda: check(lib.b200_conv1_fprop(x.ref, ptr(w_rows), ptr(bias), y.ref, ptr(stats), mode, ptr(scale),
                                             ptr(shift), stream_ptr()), "conv1_fprop"),
          shape=(x.voxels, (k_real or x.c) // 27, y.c))


def conv1_wgrad(x: ActView, dy: ActView, dw: torch.Tensor, k_real: int):
    lib = _lib.load()
    _gemm("wgrad_kernel", "conv1_wgrad", 2.0 * x.voxels * dy.c * k_real,
          lambda: check(lib.b200_conv1_wgrad(x.ref, dy.ref, ptr(dw), k_real, stream_ptr()), "conv1_wgrad"),
          shape=(x.voxels, k_real // 27, dy.c))


def conv1_direct_supported(cin: int, cout: int, w: int = 4) -> bool:
    return bool(_lib.load().b200_conv1_direct_supported(cin, cout, w))


def conv1_direct_stat_rows(n, d, h, w, cout) -> int:
    r = _lib.load().b200_conv1_direct_stat_rows(n, d, h, w, cout)
    if r <= 0:
        raise _lib.B200Error("b200_conv1_direct_stat_rows failed (no CUDA device?)")
    return r


def conv1_direct_fprop(x: torch.Tensor, w_rows, bias, y: ActView, stats=None, mode=EPI_BIAS_STATS, scale=None,
                       shift=None):
    """first conv straight from the fp32 (N, C, D, H, W) input: im2col rows exist only in shared memory"""
    lib = _lib.load()
    n, c, d, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    _gemm("igemm_im2col5_kernel", "conv1_fprop", 2.0 * y.voxels * y.c * 27 * c,
          lambda: check(lib.b200_conv1_direct_fprop(ptr(x), n, c, d, h, w, ptr(w_rows), ptr(bias), y.ref, ptr(stats),
                                                    mode, ptr(scale), ptr(shift), stream_ptr()), "conv1_direct_fprop"),
          shape=(y.voxels, c, y.c))


def conv1_march_supported(cin: int, cout: int) -> bool:
    return bool(_lib.load().b200_conv1_march_supported(cin, cout))


def conv1_march_stat_rows(n, d, h, w, cout) -> int:
    r = _lib.load().b200_conv1_march_stat_rows(n, d, h, w, cout)
    if r <= 0:
        raise _lib.B200Error("b200_conv1_march_stat_rows failed (no CUDA device?)")
    return r


def pack_conv1_slices(w: torch.Tensor, out):
    """(Cout, Cin, 3, 3, 3) fp32 -> bf16 [3][Cout][64] (csrc/conv1_march.cu)"""
    _launched(1)
    assert w.dtype == torch.float32 and w.is_contiguous()
    check(_lib.load().b200_pack_conv1_slices(ptr(w), w.shape[0], w.shape[1], ptr(out), stream_ptr()),
          "pack_conv1_slices")


def conv1_march_fprop(x: torch.Tensor, w_slices, bias, y: ActView, stats=None, mode=EPI_BIAS_STATS, scale=None,
                      shift=None):
    """first conv straight from the fp32 (N, C, D, H, W) input, depth-marching form (one slice image per input slice)"""
    lib = _lib.load()
    n, c, d, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    _gemm("conv1_march_kernel", "conv1_fprop", 2.0 * y.voxels * y.c * 27 * c,
          lambda: check(lib.b200_conv1_march_fprop(ptr(x), n, c, d, h, w, ptr(w_slices), ptr(bias), y.ref, ptr(stats),
                                                   mode, ptr(scale), ptr(shift), stream_ptr()), "conv1_march_fprop"),
          shape=(y.voxels, c, y.c))


def conv1_march_wgrad(x: torch.Tensor, dy: ActView, dw: torch.Tensor):
    """weight gradient of the first conv by the same depth march; dw fp32 (Cout, C*27) / (Cout, C, 3, 3, 3), +="""
    lib = _lib.load()
    n, c, d, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and dw.is_contiguous() and dw.dtype == torch.float32
    _gemm("conv1_march_wgrad_kernel", "conv1_wgrad", 2.0 * dy.voxels * dy.c * 27 * c,
          lambda: check(lib.b200_conv1_march_wgrad(ptr(x), n, c, d, h, w, dy.ref, ptr(dw), stream_ptr()),
                        "conv1_march_wgrad"),
          shape=(dy.voxels, c, dy.c))


def conv1_direct_wgrad(x: torch.Tensor, dy: ActView, dw: torch.Tensor):
    lib = _lib.load()
    n, c, d, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    _gemm("wgrad_im2col5_kernel", "conv1_wgrad", 2.0 * dy.voxels * dy.c * 27 * c,
          lambda: check(lib.b200_conv1_direct_wgrad(ptr(x), n, c, d, h, w, dy.ref, ptr(dw), stream_ptr()),
                        "conv1_direct_wgrad"), shape=(dy.voxels, c, dy.c))


def pack_conv_weight(w: torch.Tensor, cin_pad: int, w_packed):
    _launched(1)
    cout, cin = w.shape[0], w.shape[1]
    assert w.dtype == torch.float32 and w.is_contiguous()
    check(_lib.load().b200_pack_conv_weight(ptr(w), cout, cin, cin_pad, ptr(w_packed), stream_ptr()),
          "pack_conv_weight")


def pack_convt_weight(w: torch.Tensor, bias: torch.Tensor, w_fwd, w_dgrad, bias8):
    _launched(1)
    cin, cout = w.shape[0], w.shape[1]
    assert w.dtype == torch.float32 and w.is_contiguous()
    check(_lib.load().b200_pack_convt_weight(ptr(w), ptr(bias), cin, cout, ptr(w_fwd), ptr(w_dgrad), ptr(bias8),
                                             stream_ptr()), "pack_convt_weight")


def conv3d_workspace_bytes(n, d, h, w, out_cols) -> int:
    """bytes of the fp32 split-K scratch a 3x3x3 conv / dgrad of this shape uses (0: it does not split)"""
    return int(_lib.load().b200_conv3d_workspace_bytes(n, d, h, w, out_cols))


def conv3d_stat_rows(n, d, h, w, cout, ntaps=27, with_workspace=False) -> int:
    r = _lib.load().b200_conv3d_stat_rows(n, d, h, w, cout, ntaps, 1 if with_workspace else 0)
    if r <= 0:
        raise _lib.B200Error("b200_conv3d_stat_rows failed (no CUDA device?)")
    return r


def _conv_kernel(lib, v: ActView, out_cols: int) -> str:
    """name of the kernel a 3x3x3 conv over view v is routed to (only asked for when a profile hook is installed)"""
    if profile_hook is None:
        return "igemm_kernel"
    n, d, h, w, _ = v.shape
    return ("igemm_kernel", "dmarch_kernel", "igemm_pair_kernel",
            "dmarch_pair_kernel")[lib.b200_conv3d_kernel_id(n, d, h, w, out_cols)]


def _ws(workspace):
    return (ptr(workspace), 0 if workspace is None else workspace.numel() * workspace.element_size())


def conv3d_fprop(x: ActView, w_fprop, bias, y: ActView, stats=None, mode=EPI_BIAS_STATS, scale=None, shift=None,
                 k_real=None, workspace=None):
    """workspace: fp32 scratch of >= conv3d_workspace_bytes(...) for the split-K form of the deep levels"""
    lib = _lib.load()
    wp, wb = _ws(workspace)
    split = workspace is not None and lib.b200_conv3d_workspace_bytes(*x.shape[:4], y.c) > 0
    _launched(1 if split else 0)   # + the finalize pass
    kern = "igemm_pair_kernel" if split else _conv_kernel(lib, x, y.c)
    if kern == "dmarch_pair_kernel" and lib.b200_set_dmarch_pair_mma(-1):
        kern = "dmarch2_kernel"   # the forward's weights are K-major: CTA-pair MMAs (csrc/dmarch2.cu)
    _gemm(kern, "conv3d_fprop",
          2.0 * x.voxels * y.c * (k_real or x.c) * 27,
          lambda: check(lib.b200_conv3d_fprop(x.ref, ptr(w_fprop), ptr(bias), y.ref, ptr(stats), mode, ptr(scale),
                                              ptr(shift), wp, wb, stream_ptr()), "conv3d_fprop"),
          shape=(x.voxels, k_real or x.c, y.c))


def conv3d_dgrad(dy: ActView, w_packed, dx: ActView, workspace=None):
    lib = _lib.load()
    wp, wb = _ws(workspace)
    split = workspace is not None and lib.b200_conv3d_workspace_bytes(*dy.shape[:4], dx.c) > 0
    _launched(1 if split else 0)
    _gemm("igemm_pair_kernel" if split else _conv_kernel(lib, dy, dx.c), "conv3d_dgrad",
          2.0 * dy.voxels * dy.c * dx.c * 27,
          lambda: check(lib.b200_conv3d_dgrad(dy.ref, ptr(w_packed), dx.ref, wp, wb, stream_ptr()), "conv3d_dgrad"),
          shape=(dy.voxels, dx.c, dy.c))


def conv3d_dgrad_kmajor_supported(n, d, h, w, cin) -> bool:
    return bool(_lib.load().b200_conv3d_dgrad_kmajor_supported(n, d, h, w, cin))


def conv3d_dgrad_kmajor(dy: ActView, w_packed_t, dx: ActView):
    """input gradient on the CTA-pair depth-marching kernel, from transposed packed weights [27][Cin][Cout]"""
    lib = _lib.load()
    _gemm("dmarch2_kernel", "conv3d_dgrad", 2.0 * dy.voxels * dy.c * dx.c * 27,
          lambda: check(lib.b200_conv3d_dgrad_kmajor(dy.ref, ptr(w_packed_t), dx.ref, stream_ptr()),
                        "conv3d_dgrad_kmajor"),
          shape=(dy.voxels, dx.c, dy.c))


def transpose_taps(src: torch.Tensor, dst: torch.Tensor):
    """bf16 [taps][rows][cols] -> [taps][cols][rows]"""
    _launched(1)
    taps, rows, cols = src.shape
    assert src.dtype == torch.bfloat16 and dst.dtype == torch.bfloat16 and src.is_contiguous() and dst.is_contiguous()
    check(_lib.load().b200_transpose_taps(ptr(src), taps, rows, cols, ptr(dst), stream_ptr()), "transpose_taps")


def conv3d_wgrad(x: ActView, dy: ActView, dw: torch.Tensor, cin_real: int, packed: bool = False):
    """dw (+=): torch's (Cout, Cin, 3, 3, 3), or with packed=True the engine's physical [27][Cout][Cin] buffer"""
    lib = _lib.load()
    kern = "wgrad_halo_kernel" if profile_hook and lib.b200_conv3d_wgrad_kernel_id(x.shape[2], x.shape[3]) else "wgrad_kernel"
    _gemm(kern, "conv3d_wgrad", 2.0 * x.voxels * dy.c * cin_real * 27,
          lambda: check(lib.b200_conv3d_wgrad(x.ref, dy.ref, ptr(dw), cin_real, 1 if packed else 0, stream_ptr()),
                        "conv3d_wgrad"), shape=(x.voxels, cin_real, dy.c))


def convt2x_fwd(x: ActView, w_fwd, bias8, y: ActView, pads=(0, 0, 0)):
    lib = _lib.load()
    _gemm("igemm_kernel", "convt2x_fwd", 2.0 * x.voxels * x.c * y.c * 8,
          lambda: check(lib.b200_convt2x_fwd(x.ref, ptr(w_fwd), ptr(bias8), y.ref, pads[0], pads[1], pads[2],
                                             stream_ptr()), "convt2x_fwd"), shape=(x.voxels, x.c, y.c))


def convt2x_dgrad(dy: ActView, pads, w_dgrad, dx: ActView):
    lib = _lib.load()
    _gemm("igemm_kernel", "convt2x_dgrad", 2.0 * dx.voxels * dx.c * dy.c * 8,
          lambda: check(lib.b200_convt2x_dgrad(dy.ref, pads[0], pads[1], pads[2], ptr(w_dgrad), dx.ref,
                                               stream_ptr()), "convt2x_dgrad"), shape=(dx.voxels, dx.c, dy.c))


def convt2x_wgrad(x: ActView, dy: ActView, pads, dw: torch.Tensor):
    lib = _lib.load()
    _gemm("wgrad_kernel", "convt2x_wgrad", 2.0 * x.voxels * x.c * dy.c * 8,
          lambda: check(lib.b200_convt2x_wgrad(x.ref, dy.ref, pads[0], pads[1], pads[2], ptr(dw), stream_ptr()),
                        "convt2x_wgrad"), shape=(x.voxels, x.c, dy.c))


def bn_finalize(stats, rows, count, c, gamma, beta, eps, momentum, running_mean, running_var, mean, rstd, scale,
                shift, num_batches_tracked=None):
    _launched(1)
    if num_batches_tracked is not None:
        assert num_batches_tracked.dtype == torch.int64 and num_batches_tracked.is_cuda
    check(_lib.load().b200_bn_finalize(ptr(stats), rows, count, c, ptr(gamma), ptr(beta), eps, momentum,
                                       ptr(running_mean), ptr(running_var), ptr(num_batches_tracked), ptr(mean),
                                       ptr(rstd), ptr(scale), ptr(shift), stream_ptr()), "bn_finalize")


def bn_fold_eval(gamma, beta, running_mean, running_var, conv_bias, eps, scale, shift):
    _launched(1)
    check(_lib.load().b200_bn_fold_eval(ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), ptr(conv_bias),
                                        eps, gamma.numel(), ptr(scale), ptr(shift), stream_ptr()), "bn_fold_eval")


def bn_apply_relu(y: ActView, scale, shift, out: ActView):
    _launched(1)
    check(_lib.load().b200_bn_apply_relu(y.ref, ptr(scale), ptr(shift), out.ref, stream_ptr()), "bn_apply_relu")


def bn_bwd_max_blocks() -> int:
    return _lib.load().b200_bn_bwd_max_blocks()


def bn_bwd(dout: ActView, y: ActView, scale, shift, mean, rstd, gamma, partial, coef, dgamma, dbeta, dy: ActView,
           dbias):
    """BatchNorm3d(train)+ReLU backward: reduce -> finalize -> apply (three launches)."""
    _launched(3)
    lib = _lib.load()
    nblk = C.c_int(0)
    s = stream_ptr()
    check(lib.b200_bn_bwd_reduce(dout.ref, y.ref, ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(partial),
                                 C.byref(nblk), s), "bn_bwd_reduce")
    check(lib.b200_bn_bwd_finalize(ptr(partial), nblk.value, y.c, y.voxels, ptr(dgamma), ptr(dbeta), ptr(coef), s),
          "bn_bwd_finalize")
    check(lib.b200_bn_bwd_apply(dout.ref, y.ref, ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(gamma), ptr(coef),
                                dy.ref, ptr(dbias), s), "bn_bwd_apply")


def bn_apply_relu_pool(y: ActView, scale, shift, out: ActView, pooled: ActView):
    """out = relu(y * scale + shift) and pooled = MaxPool3d(2)(out) in one pass (encoder blocks)"""
    _launched(1)
    check(_lib.load().b200_bn_apply_relu_pool(y.ref, ptr(scale), ptr(shift), out.ref, pooled.ref, stream_ptr()),
          "bn_apply_relu_pool")


def bn_bwd_head(dlogits, w, y: ActView, scale, shift, mean, rstd, gamma, partial, coef, dgamma, dbeta, dy: ActView,
                dbias, dw, db):
    """BatchNorm3d(train)+ReLU backward of the network's last BatchNorm from the head's dlogits (N, ncls, D, H, W): the
    head's input gradient is recomputed in both passes, the head's dw / db are accumulated by the first."""
    _launched(3)
    lib = _lib.load()
    nblk = C.c_int(0)
    s = stream_ptr()
    ncls = w.shape[0]
    check(lib.b200_bn_bwd_reduce_head(ptr(dlogits), ptr(w), ncls, y.ref, ptr(scale), ptr(shift), ptr(mean), ptr(rstd),
                                      ptr(partial), C.byref(nblk), ptr(dw), ptr(db), s), "bn_bwd_reduce_head")
    check(lib.b200_bn_bwd_finalize(ptr(partial), nblk.value, y.c, y.voxels, ptr(dgamma), ptr(dbeta), ptr(coef), s),
          "bn_bwd_finalize")
    check(lib.b200_bn_bwd_apply_head(ptr(dlogits), ptr(w), ncls, y.ref, ptr(scale), ptr(shift), ptr(mean), ptr(rstd),
                                     ptr(coef), dy.ref, ptr(dbias), s), "bn_bwd_apply_head")


def maxpool3d_fwd(x: ActView, y: ActView):
    _launched(1)
    check(_lib.load().b200_maxpool3d_fwd(x.ref, y.ref, stream_ptr()), "maxpool3d_fwd")


def maxpool3d_bwd(x: ActView, dy: ActView, dskip, dx: ActView):
    _launched(1)
    check(_lib.load().b200_maxpool3d_bwd(x.ref, None, dy.ref, dskip.ref if dskip is not None else None, dx.ref,
                                         stream_ptr()), "maxpool3d_bwd")


def head_fwd(x: ActView, w, b, logits, probs=None):
    _launched(1)
    check(_lib.load().b200_head_fwd(x.ref, ptr(w), ptr(b), w.shape[0], ptr(logits), ptr(probs), stream_ptr()),
          "head_fwd")


def head_bwd(x: ActView, w, dlogits, dx: ActView, dw, db):
    _launched(1)
    check(_lib.load().b200_head_bwd(x.ref, ptr(w), w.shape[0], ptr(dlogits), dx.ref, ptr(dw), ptr(db), stream_ptr()),
          "head_bwd")


def loss_fwd(logits, target, bce_w, dice_w, smooth, workspace, sums, loss):
    _launched(2)
    check(_lib.load().b200_loss_fwd(ptr(logits), ptr(target), logits.numel(), bce_w, dice_w, smooth, ptr(workspace),
                                    ptr(sums), ptr(loss), stream_ptr()), "loss_fwd")


def loss_bwd(logits, target, bce_w, dice_w, smooth, sums, gout, dlogits):
    _launched(1)
    check(_lib.load().b200_loss_bwd(ptr(logits), ptr(target), logits.numel(), bce_w, dice_w, smooth, ptr(sums),
                                    ptr(gout), ptr(dlogits), stream_ptr()), "loss_bwd")


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0,
              found_inf=None, bf16_shadow=None, dyn_scalars=None):
    _launched(1)
    check(_lib.load().b200_adam_step(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), param.numel(), lr, beta1,
                                     beta2, eps, weight_decay, step, grad_scale, ptr(found_inf), ptr(bf16_shadow),
                                     ptr(dyn_scalars), stream_ptr()), "adam_step")


def cast_bf16(x, out):
    _launched(1)
    check(_lib.load().b200_cast_bf16(ptr(x), x.numel(), ptr(out), stream_ptr()), "cast_bf16")


def sumsq(x, out):
    _launched(1)
    check(_lib.load().b200_sumsq(ptr(x), x.numel(), ptr(out), stream_ptr()), "sumsq")


def fill_zero(v: ActView):
    _launched(1)
    check(_lib.load().b200_fill_zero(v.ref, stream_ptr()), "fill_zero")


def channel_sum(v: ActView, out):
    _launched(1)
    check(_lib.load().b200_channel_sum(v.ref, ptr(out), stream_ptr()), "channel_sum")


def channel_sum_box(v: ActView, origin, extent, out):
    """out[c] += sum over the box origin + [0, extent) of every sample"""
    _launched(1)
    check(_lib.load().b200_channel_sum_box(v.ref, int(origin[0]), int(origin[1]), int(origin[2]), int(extent[0]),
                                           int(extent[1]), int(extent[2]), ptr(out), stream_ptr()), "channel_sum_box")


def window_gather(x: torch.Tensor, origins: torch.Tensor, window) -> torch.Tensor:
    """x (N,C,D,H,W) fp32, origins int32 (nwin,4) on the device = (volume, d0, h0, w0) -> (nwin,C,*window) fp32"""
    _cuda_f32(x, "window_gather")
    n, c, d, h, w = x.shape
    nwin = origins.shape[0]
    out = torch.empty((nwin, c) + tuple(window), device=x.device, dtype=torch.float32)
    _launched(1)
    check(_lib.load().b200_window_gather(ptr(x), n, c, d, h, w, ptr(origins), nwin, window[0], window[1], window[2],
                                         ptr(out), stream_ptr()), "window_gather")
    return out


def window_accumulate(logits: torch.Tensor, origins: torch.Tensor, acc: torch.Tensor, v_lo: int, v_cnt: int):
    """acc (N,K,D,H,W) += logits (nwin,K,wd,wh,ww) at their origins, in list order, for volumes [v_lo, v_lo+v_cnt)"""
    _cuda_f32(logits, "window_accumulate")
    _cuda_f32(acc, "window_accumulate")
    nwin, k, wd, wh, ww = logits.shape
    n, _, d, h, w = acc.shape
    _launched(1)
    check(_lib.load().b200_window_accumulate(ptr(logits), ptr(origins), nwin, k, wd, wh, ww, ptr(acc), n, d, h, w,
                                             int(v_lo), int(v_cnt), stream_ptr()), "window_accumulate")


def window_finalize(acc: torch.Tensor, cover: torch.Tensor, threshold=0.5, probs=None, mask=None):
    """in place: acc (N,K,D,H,W) /= per-voxel window count (cover: int32 [D+H+W] per-axis counts); optional sigmoid
    probabilities and thresholded mask"""
    _cuda_f32(acc, "window_finalize")
    n, k, d, h, w = acc.shape
    _launched(1)
    check(_lib.load().b200_window_finalize(ptr(acc), ptr(cover), n * k, d, h, w, float(threshold), ptr(probs),
                                           ptr(mask), stream_ptr()), "window_finalize")


def _cuda_f32(t, what):
    if not t.is_cuda:
        raise _lib.B200Error(f"{what}: b200 kernels need CUDA tensors: there is no CPU path")
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError(f"{what}: expected a contiguous float32 tensor")


def resample3d(x: torch.Tensor, size, nearest: bool = False, binarize: bool = False) -> torch.Tensor:
    """(..., D, H, W) fp32 -> (..., *size): the reference loader's resample-to-target (script/data_loader.py:240-283
    linear for images, :395-409 nearest + `> 0` for labels) with ITK's index mapping (see include/b200_unet3d.h)."""
    _cuda_f32(x, "resample3d")
    if x.dim() < 3:
        raise ValueError(f"resample3d: expected at least 3 dimensions, got shape {tuple(x.shape)}")
    d, h, w = x.shape[-3:]
    nvol = x.numel() // max(1, d * h * w)
    out = torch.empty(tuple(x.shape[:-3]) + tuple(int(v) for v in size), device=x.device, dtype=torch.float32)
    _launched(1)
    check(_lib.load().b200_resample3d(ptr(x), nvol, d, h, w, ptr(out), int(size[0]), int(size[1]), int(size[2]),
                                      int(nearest), int(binarize), stream_ptr()), "resample3d")
    return out


def minmax_normalize_(x: torch.Tensor) -> torch.Tensor:
    """in place, per leading index: (x - min) / (max - min) over the last three axes (script/predict.py:69-75)"""
    _cuda_f32(x, "minmax_normalize_")
    per = x.shape[-1] * x.shape[-2] * x.shape[-3]
    nvol = x.numel() // max(1, per)
    ws = torch.empty(2 * max(nvol, 1), device=x.device, dtype=torch.int32)
    _launched(3)
    check(_lib.load().b200_minmax_normalize(ptr(x), nvol, per, ptr(ws), stream_ptr()), "minmax_normalize")
    return x


def seg_counts(score: torch.Tensor, label: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    """(N, ...) scores and labels -> int64 (N, 3): |P & T|, |P|, |T| with P = score > threshold, T = label > 0.5"""
    _cuda_f32(score, "seg_counts")
    _cuda_f32(label, "seg_counts")
    if score.shape != label.shape:
        raise ValueError(f"seg_counts: score shape {tuple(score.shape)} != label shape {tuple(label.shape)}")
    n = score.shape[0]
    counts = torch.zeros(n, 3, device=score.device, dtype=torch.int64)
    _launched(1)
    check(_lib.load().b200_seg_counts(ptr(score), ptr(label), n, score.numel() // max(n, 1), float(threshold),
                                      ptr(counts), stream_ptr()), "seg_counts")
    return counts


# ---- every op runs on the device that owns its tensors -------------------------------------------------------------
# The C ABI launches on the CURRENT device with the stream it is handed; a tensor on cuda:1 while cuda:0 is current
# would otherwise get device 0's stream and device 0's kernel attributes.  The wrapper switches the current device to
# the first CUDA argument's device for the duration of the call (a no-op in the usual one-process-per-GPU setting).
# Optional timeline (dev tool): `timeline = []` records CUDA events around every op on whatever stream it runs, to
# inspect cross-stream overlap without an external profiler; entries are (name, stream, start, end).
timeline = None


def _device_of(args):
    for a in args:
        if isinstance(a, ActView):
            return a.t.device
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a.device
    return None


def _run(fn, a, k):
    if timeline is None:
        return fn(*a, **k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn(*a, **k)
    e1.record()
    timeline.append((fn.__name__, torch.cuda.current_stream().cuda_stream, e0, e1))
    return r


def _op(fn):
    def wrapper(*a, **k):
        dev = _device_of(a)
        if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):
                return _run(fn, a, k)
        return _run(fn, a, k)
    wrapper.__name__ = fn.__name__
    wrapper.__doc__ = fn.__doc__
    return wrapper


for _n in ("pack_input", "im2col_input", "pack_rows", "pack_conv_weight", "pack_convt_weight", "conv3d_fprop",
           "conv1_fprop", "conv3d_dgrad", "conv3d_wgrad", "conv1_wgrad", "conv1_direct_fprop", "conv1_direct_wgrad", "convt2x_fwd", "convt2x_dgrad",
           "convt2x_wgrad", "bn_finalize", "bn_fold_eval", "bn_apply_relu", "bn_bwd", "bn_apply_relu_pool", "bn_bwd_head",
           "maxpool3d_fwd",
           "maxpool3d_bwd", "head_fwd", "head_bwd", "loss_fwd", "loss_bwd", "adam_step", "cast_bf16", "sumsq",
           "fill_zero", "channel_sum", "channel_sum_box", "window_gather", "window_accumulate", "window_finalize",
           "resample3d", "minmax_normalize_", "seg_counts"):
    globals()[_n] = _op(globals()[_n])
