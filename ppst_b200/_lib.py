"""ctypes binding of libppst_b200.so (include/ppst_b200.h).

There is NO fallback: if the shared library is missing or a CUDA tensor is not supplied, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libppst_b200.so")

F32, F16, BF16 = 0, 1, 2
_DT = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}

c_i64, c_int, c_float, c_double, c_void_p = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_void_p


class ConvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("n", "in_h", "in_w", "cin", "out_h", "out_w", "cout", "kh", "kw", "stride", "off_y", "off_x",
                 "phases_y", "phases_x", "phase_off_dy", "phase_off_dx", "full_h", "full_w", "cout_stride", "c_off",
                 "has_act")] + \
               [("slope", c_float), ("gain", c_float), ("noise_w", c_float),
                ("in_dtype", ctypes.c_int32), ("out_dtype", ctypes.c_int32),
                ("in_row_pitch", ctypes.c_int64), ("in_img_pitch", ctypes.c_int64),
                ("w_img_rows", ctypes.c_int32), ("split_terms", ctypes.c_int32),
                ("out_pad", ctypes.c_int32), ("bias_img_stride", ctypes.c_int32), ("out_pad_trim", ctypes.c_int32)]


class WgradDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("n", "in_h", "in_w", "cin", "out_h", "out_w", "cout", "kh", "kw", "stride", "off_y", "off_x")]


class ApplyDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("n", "h", "w", "c", "up", "pad_t", "pad_b", "pad_l", "pad_r", "pad_mode")] + \
               [("pre_slope", c_float), ("pre_gain", c_float), ("post_slope", c_float), ("post_gain", c_float)] + \
               [(n, ctypes.c_int32) for n in ("res_mode", "x_dtype", "res_dtype", "out_dtype", "out_c_stride",
                                              "out_c_off", "out_border", "res_border")]


_SIGS = {
    "ppst_last_error": (ctypes.c_char_p, []),
    "ppst_version": (c_int, []),
    "ppst_launch_count": (ctypes.c_longlong, []),
    "ppst_fused_bias_act": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_int,
                                    c_float, c_float, c_int, c_void_p]),
    "ppst_bias_grad_reduce": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_void_p]),
    "ppst_act_grad_sum_scratch_elems": (c_i64, [c_i64, c_int]),
    "ppst_act_grad_sum": (c_int, [c_void_p] * 5 + [c_i64, c_int, c_float, c_float, c_int, c_int, c_void_p]),
    "ppst_upfirdn2d": (c_int, [c_void_p, c_void_p, c_void_p, c_i64] + [c_int] * 14 + [c_void_p]),
    "ppst_blur_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_i64] + [c_int] * 11 + [c_void_p]),
    "ppst_plane_reduce_scratch_elems": (c_i64, [c_i64, c_i64, c_i64, c_int]),
    "ppst_plane_reduce": (c_int, [c_void_p] * 7 + [c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ppst_upsample_reduce_scratch_elems": (c_i64, [c_i64, c_int, c_int]),
    "ppst_upsample_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_i64] + [c_int] * 6 + [c_void_p, c_void_p]),
    "ppst_plane_dot_scratch_elems": (c_i64, [c_i64, c_i64, c_i64]),
    "ppst_plane_dot": (c_int, [c_void_p] * 5 + [c_i64, c_i64, c_i64, c_int, c_void_p]),
    "ppst_plane_count_eq": (c_int, [c_void_p] * 4 + [c_i64, c_i64, c_i64, c_int, c_void_p]),
    "ppst_pool_backward": (c_int, [c_void_p] * 5 + [c_i64, c_i64, c_i64, c_int, c_void_p]),
    "ppst_labels_pool_or": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_void_p]),
    "ppst_labels_to_onehot": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_int, c_void_p]),
    "ppst_nchw_to_nhwc": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_void_p]),
    "ppst_nhwc_to_nchw": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_i64, c_i64, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p]),
    "ppst_affine_apply_stats": (c_int, [c_void_p] * 7 + [ctypes.POINTER(ApplyDesc), c_void_p, c_void_p]),
    "ppst_affine_apply_torgb": (c_int, [c_void_p] * 8 + [ctypes.POINTER(ApplyDesc), c_void_p]),
    "ppst_conv_simt": (c_int, [c_void_p] * 5 + [ctypes.POINTER(ConvDesc), c_void_p]),
    "ppst_conv_tc_stats_elems": (c_i64, [ctypes.POINTER(ConvDesc)]),
    "ppst_conv_tc": (c_int, [c_void_p] * 6 + [ctypes.POINTER(ConvDesc), c_void_p]),
    "ppst_conv_tc_available": (c_int, []),
    "ppst_fold_fir": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ppst_pack_conv_weights": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                       ctypes.POINTER(ctypes.c_int), c_int, c_float, c_void_p]),
    "ppst_conv_wgrad_workspace_bytes": (c_i64, [ctypes.POINTER(WgradDesc)]),
    "ppst_conv_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.POINTER(WgradDesc), c_void_p, c_i64, c_void_p]),
    "ppst_scale_weights": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_i64, c_int, c_int, c_void_p]),
    "ppst_modulate_weights": (c_int, [c_void_p] * 6 + [c_i64, c_i64, c_int, c_int, c_void_p]),
    "ppst_fill_border": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ppst_pad_border": (c_int, [c_void_p, c_i64] + [c_int] * 9 + [c_void_p]),
    "ppst_conv1x1_in_nchw": (c_int, [c_void_p] * 4 + [c_i64, c_int, c_i64, c_int, c_int, c_int, c_int, c_float, c_float,
                                                     c_int, c_int, c_int, c_void_p]),
    "ppst_conv1x1_thin_wgrad_scratch_elems": (c_i64, [c_i64, c_i64, c_int]),
    "ppst_conv1x1_thin_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_i64, c_int, c_int, c_void_p, c_void_p]),
    "ppst_conv1x1_thin_out": (c_int, [c_void_p] * 4 + [c_i64, c_int, c_int, c_int, c_void_p]),
    "ppst_in_coeffs": (c_int, [c_void_p] * 6 + [c_i64, c_void_p, c_i64, c_i64, c_i64, c_double, c_float, c_int, c_void_p]),
    "ppst_affine_apply": (c_int, [c_void_p] * 7 + [ctypes.POINTER(ApplyDesc), c_void_p]),
    "ppst_resample": (c_int, [c_void_p, c_void_p, c_i64] + [c_int] * 10 + [c_void_p, c_void_p, c_float, c_int, c_int,
                                                                           c_void_p]),
    "ppst_upsample2x_backward": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ppst_reflect_pad_backward": (c_int, [c_void_p, c_void_p, c_i64] + [c_int] * 8 + [c_void_p]),
    "ppst_linear": (c_int, [c_void_p] * 4 + [c_i64] * 4 + [c_float, c_float, c_int, c_void_p]),
    "ppst_linear_dgrad_scratch_elems": (c_i64, [c_i64, c_i64, c_i64]),
    "ppst_linear_dgrad": (c_int, [c_void_p] * 4 + [c_i64, c_i64, c_i64, c_void_p]),
    "ppst_linear_wgrad": (c_int, [c_void_p] * 3 + [c_i64, c_i64, c_i64, c_void_p]),
    "ppst_l2_normalize": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_float, c_int, c_void_p]),
    "ppst_lerp": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_i64, c_void_p]),
    "ppst_rselfcorr": (c_int, [c_void_p, c_void_p, c_i64] + [c_int] * 7 + [c_void_p]),
    "ppst_corr_prepare": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_void_p]),
    "ppst_corr_prepare_split": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_void_p]),
    "ppst_split_bf16": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_i64, c_int, c_void_p]),
    "ppst_softmax_rows_split": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_float, c_void_p]),
    "ppst_gemm_f32": (c_int, [c_void_p] * 3 + [c_i64] * 10 + [c_int, c_float, c_void_p]),
    "ppst_softmax_rows": (c_int, [c_void_p, c_i64, c_i64, c_float, c_void_p]),
    "ppst_resample_u8": (c_int, [c_void_p] * 4 + [c_i64] + [c_int] * 6 + [c_void_p]),
    "ppst_u8_to_nchw_norm": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_int, c_float, c_float, c_void_p]),
    "ppst_tensor2im": (c_int, [c_void_p, c_void_p, c_i64, c_i64, c_int, c_int, c_void_p]),
    "ppst_guided_smooth_workspace_bytes": (c_i64, [c_i64, c_int, c_int]),
    "ppst_guided_smooth": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_float, c_void_p, c_i64,
                                   c_void_p]),
}

EXPORTED = sorted(_SIGS)

_lib = None


def load():
    """Load the shared library (once).  Raises ImportError with the build hint if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "ppst_b200: %s is missing -- build it with `python -m ppst_b200.build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("ppst_b200: %s (code %d)" % (load().ppst_last_error().decode(), rc))


def dt(t):
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError("ppst_b200: unsupported dtype %s (float32 / float16 / bfloat16)" % t.dtype)


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    """Every tensor of a launch must live on the CURRENT CUDA device: the launch goes to that device's current stream
    with raw pointers (SwapModel / Engine enter `torch.cuda.device(self.device)` themselves; direct callers of the
    operators set the device as they would for any torch op)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("ppst_b200: expected a CUDA tensor (there is no CPU fallback); got device %s" % t.device)
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError("ppst_b200: tensor on %s but the current CUDA device is cuda:%d -- wrap the call in "
                               "torch.cuda.device(tensor.device)" % (t.device, cur))


def launch_count():
    return int(load().ppst_launch_count())
