"""ctypes binding of the C ABI in include/b200_unet3d.h.

There is no fallback: if the library is missing, or a call fails, an exception is raised.
"""
import ctypes as C
import os

import torch

from . import build as _build

_i64, _i32, _f32, _f64, _vp = C.c_int64, C.c_int, C.c_float, C.c_double, C.c_void_p


class B200Error(RuntimeError):
    pass


class Act(C.Structure):
    """b200_act: NDHWC bf16 view (pointer to channel 0 of voxel 0, extents, voxel pitch)."""
    _fields_ = [("ptr", _vp), ("n", _i64), ("d", _i64), ("h", _i64), ("w", _i64), ("c", _i64), ("ld", _i64)]


_AP = C.POINTER(Act)

# name -> (restype, argtypes); mirrors include/b200_unet3d.h one to one
SIGNATURES = {
    "b200_last_error": (C.c_char_p, []),
    "b200_abi_version": (_i32, []),
    "b200_set_pdl": (_i32, [_i32]),
    "b200_set_dmarch_pair_mma": (_i32, [_i32]),
    "b200_sm_count": (_i32, []),
    "b200_pack_input": (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _AP, _vp]),
    "b200_im2col_input": (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _AP, _vp]),
    "b200_pack_rows": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "b200_conv1_fprop": (_i32, [_AP, _vp, _vp, _AP, _vp, _i32, _vp, _vp, _vp]),
    "b200_conv1_wgrad": (_i32, [_AP, _AP, _vp, _i32, _vp]),
    "b200_conv1_direct_supported": (_i32, [_i64, _i64, _i64]),
    "b200_conv1_direct_stat_rows": (_i32, [_i64, _i64, _i64, _i64, _i64]),
    "b200_conv1_direct_fprop": (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _AP, _vp, _i32, _vp, _vp, _vp]),
    "b200_conv1_direct_wgrad": (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _AP, _vp, _vp]),
    "b200_conv1_march_supported": (_i32, [_i64, _i64]),
    "b200_conv1_march_stat_rows": (_i32, [_i64, _i64, _i64, _i64, _i64]),
    "b200_pack_conv1_slices": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "b200_conv1_march_fprop": (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _AP, _vp, _i32, _vp, _vp, _vp]),
    "b200_conv1_march_wgrad": (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _AP, _vp, _vp]),
    "b200_pack_conv_weight": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "b200_pack_convt_weight": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b200_conv3d_mtiles": (_i64, [_i64, _i64, _i64, _i64]),
    "b200_conv3d_stat_rows": (_i32, [_i64, _i64, _i64, _i64, _i64, _i32, _i32]),
    "b200_conv3d_workspace_bytes": (_i64, [_i64, _i64, _i64, _i64, _i64]),
    "b200_conv3d_fprop": (_i32, [_AP, _vp, _vp, _AP, _vp, _i32, _vp, _vp, _vp, _i64, _vp]),
    "b200_conv3d_dgrad": (_i32, [_AP, _vp, _AP, _vp, _i64, _vp]),
    "b200_conv3d_dgrad_kmajor_supported": (_i32, [_i64, _i64, _i64, _i64, _i64]),
    "b200_conv3d_dgrad_kmajor": (_i32, [_AP, _vp, _AP, _vp]),
    "b200_transpose_taps": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "b200_conv3d_wgrad": (_i32, [_AP, _AP, _vp, _i32, _i32, _vp]),
    "b200_convt2x_fwd": (_i32, [_AP, _vp, _vp, _AP, _i32, _i32, _i32, _vp]),
    "b200_convt2x_dgrad": (_i32, [_AP, _i32, _i32, _i32, _vp, _AP, _vp]),
    "b200_convt2x_wgrad": (_i32, [_AP, _AP, _i32, _i32, _i32, _vp, _vp]),
    "b200_bn_finalize": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200_bn_fold_eval": (_i32, [_vp, _vp, _vp, _vp, _vp, _f32, _i32, _vp, _vp, _vp]),
    "b200_bn_apply_relu": (_i32, [_AP, _vp, _vp, _AP, _vp]),
    "b200_bn_bwd_max_blocks": (_i32, []),
    "b200_bn_bwd_reduce": (_i32, [_AP, _AP, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i32), _vp]),
    "b200_bn_bwd_finalize": (_i32, [_vp, _i32, _i32, _i64, _vp, _vp, _vp, _vp]),
    "b200_bn_bwd_apply": (_i32, [_AP, _AP, _vp, _vp, _vp, _vp, _vp, _vp, _AP, _vp, _vp]),
    "b200_bn_apply_relu_pool": (_i32, [_AP, _vp, _vp, _AP, _AP, _vp]),
    "b200_bn_bwd_reduce_head": (_i32, [_vp, _vp, _i32, _AP, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i32), _vp, _vp, _vp]),
    "b200_bn_bwd_apply_head": (_i32, [_vp, _vp, _i32, _AP, _vp, _vp, _vp, _vp, _vp, _AP, _vp, _vp]),
    "b200_maxpool3d_fwd": (_i32, [_AP, _AP, _vp]),
    "b200_maxpool3d_bwd": (_i32, [_AP, _AP, _AP, _AP, _AP, _vp]),
    "b200_head_fwd": (_i32, [_AP, _vp, _vp, _i32, _vp, _vp, _vp]),
    "b200_head_bwd": (_i32, [_AP, _vp, _i32, _vp, _AP, _vp, _vp, _vp]),
    "b200_loss_fwd": (_i32, [_vp, _vp, _i64, _f32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "b200_loss_bwd": (_i32, [_vp, _vp, _i64, _f32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "b200_adam_step": (_i32, [_vp, _vp, _vp, _vp, _i64, _f64, _f64, _f64, _f64, _f64, _i64, _f64, _vp, _vp, _vp, _vp]),
    "b200_cast_bf16": (_i32, [_vp, _i64, _vp, _vp]),
    "b200_sumsq": (_i32, [_vp, _i64, _vp, _vp]),
    "b200_fill_zero": (_i32, [_AP, _vp]),
    "b200_channel_sum": (_i32, [_AP, _vp, _vp]),
    "b200_channel_sum_box": (_i32, [_AP, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "b200_window_gather": (_i32, [_vp, _i64, _i64, _i64, _i64, _i64, _vp, _i32, _i64, _i64, _i64, _vp, _vp]),
    "b200_window_accumulate": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _i32,
                                      _i32, _vp]),
    "b200_window_finalize": (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _f32, _vp, _vp, _vp]),
    "b200_unpack_act": (_i32, [_AP, _vp, _vp]),
    "b200_resample3d": (_i32, [_vp, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _vp]),
    "b200_minmax_normalize": (_i32, [_vp, _i64, _i64, _vp, _vp]),
    "b200_seg_counts": (_i32, [_vp, _vp, _i64, _i64, _f32, _vp, _vp]),
    "b200_conv3d_kernel_id": (_i32, [_i64, _i64, _i64, _i64, _i64]),
    "b200_conv3d_wgrad_kernel_id": (_i32, [_i64, _i64]),
}

# development library only (build.build(dev=True), -DB200_DEV): tcgen05 micro-probes and kernel ablation switches
DEV_SIGNATURES = {
    "b200_probe_pair": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp]),
    "b200_probe_mma": (_i32, [_i32, _i32, _i32, _vp, _i32, _vp]),
    "b200_probe_mma2": (_i32, [_i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "b200_dev_set_ablation": (_i32, [_i32, _i32, _i32, _i32]),
    "b200_dev_set_variant": (_i32, [_i32, _i32]),
}

ABI_VERSION = 5   # must equal b200_abi_version() of the loaded library

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def _bind(lib, table):
    for name, (res, args) in table.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args


def load():
    """dlopen the C-ABI library, (re)building it first when it is missing or its sources changed.

    No fallback: a failed rebuild of a stale library raises (an old binary with other kernels or another ABI is never
    loaded silently), and so does an ABI version mismatch.  Only when no compiler exists at all (a deployment box
    without nvcc) is an existing library loaded as it is — with a warning if its recorded source hash differs.
    B200_DEV=1 selects the development variant (ablation switches, probes)."""
    global _lib
    if _lib is not None:
        return _lib
    dev = os.environ.get("B200_DEV") == "1"
    path = _build.DEV_LIB_PATH if dev else _build.LIB_PATH
    try:
        path = _build.build(dev=dev)  # no-op unless the library is missing or its sources changed
    except FileNotFoundError as e:    # nvcc itself is absent
        if not os.path.exists(path):
            raise B200Error(f"{path} is missing and cannot be built here ({e}); there is no fallback") from e
        if _build.is_stale(dev):
            import warnings
            warnings.warn(f"{path}: sources differ from the ones it was built from and no nvcc is available to "
                          "rebuild it; loading the existing binary", RuntimeWarning)
    lib = C.CDLL(path)
    _bind(lib, SIGNATURES)
    if dev:
        _bind(lib, DEV_SIGNATURES)
    got = lib.b200_abi_version()
    if got != ABI_VERSION:
        raise B200Error(f"{path}: ABI version {got}, this package needs {ABI_VERSION} (stale library?)")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().b200_last_error().decode("utf-8", "replace")
        raise B200Error(f"{what or 'b200 call'} failed (status {rc}): {msg}")


def stream_ptr(device=None) -> int:
    """raw cudaStream_t of torch's current stream on `device` (default: the current device)"""
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    """device pointer of a tensor (None -> NULL)"""
    return 0 if t is None else t.data_ptr()
