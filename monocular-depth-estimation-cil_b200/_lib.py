"""ctypes binding of libdepth_b200.so (the C ABI declared in include/depth_b200.h).

There is no CPU fallback: if the shared library is missing or a call is made without a CUDA
device, this module raises.  PyTorch is used only for device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdepth_b200.so")

c_int, c_uint, c_float, c_size_t, c_void_p, c_u64 = (ctypes.c_int, ctypes.c_uint, ctypes.c_float, ctypes.c_size_t,
                                                      ctypes.c_void_p, ctypes.c_uint64)
c_ll = ctypes.c_longlong
c_double = ctypes.c_double


class DepthB200Error(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DepthB200Error(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback for the hot path.")
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
        if os.environ.get("DP_TRACE"):
            _lib = _Traced(_lib)
        if _lib.dp_abi_version() != 3:
            raise DepthB200Error("libdepth_b200.so ABI version mismatch")
    return _lib


class _Traced:
    """DP_TRACE=1: print every C-ABI call (name + scalar arguments) to stderr and synchronise after it, so a faulting
    kernel is attributed to the call that launched it.  Diagnostics only."""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith("dp_") or name in ("dp_last_error", "dp_launch_count", "dp_abi_version"):
            return fn

        def call(*a):
            import sys
            sys.stderr.write(f"[dp] {name} {[x for x in a if isinstance(x, (int, float)) and abs(x) < (1 << 32)]}\n")
            sys.stderr.flush()
            r = fn(*a)
            if torch.cuda.is_available() and not torch.cuda.is_current_stream_capturing():
                torch.cuda.synchronize()
            return r
        return call


def _declare(L):
    L.dp_abi_version.restype = c_int
    L.dp_last_error.restype = ctypes.c_char_p
    L.dp_launch_count.restype = c_u64
    for name in dir(_Sig):
        if name.startswith("dp_"):
            fn = getattr(L, name)
            res, args = getattr(_Sig, name)
            fn.restype = res
            fn.argtypes = args


P = c_void_p


class _Sig:
    dp_depth_moments_workspace = (c_size_t, [c_int, c_int, c_int])
    dp_rgb_minmax_bytes = (c_size_t, [])
    dp_rgb_gradmag_minmax = (c_int, [P, c_int, c_int, c_int, P, P])
    dp_depth_moments = (c_int, [P, P, P, P, c_int, c_int, c_int, c_uint, c_float, P, P, c_size_t, P])
    dp_loss_combine = (c_int, [P, c_int, c_int, c_int, c_uint, c_float, c_float, c_float, c_float, c_float, c_int,
                               P, P, P])
    dp_loss_backward = (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_uint, c_float, c_float, c_float, c_float,
                                c_float, c_float, P, P])
    dp_delta_counts = (c_int, [P, P, P, c_int, c_int, c_int, ctypes.POINTER(c_float), c_int, c_int, c_float, P, P,
                               c_size_t, P])
    dp_metrics_combine = (c_int, [P, P, c_int, c_int, c_int, c_int, P, P])
    dp_per_pixel_si = (c_int, [P, P, P, c_int, c_int, c_int, P, P])
    dp_eval_metrics_workspace = (c_size_t, [c_int, c_int, c_int])
    dp_eval_metrics_plan = (c_int, [ctypes.c_longlong, c_int, c_int, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                    ctypes.POINTER(c_int), ctypes.POINTER(c_size_t)])
    dp_eval_metrics = (c_int, [P, P, c_int, c_int, c_int, ctypes.POINTER(c_float), c_int, c_float, c_int, P, P, P, P, c_size_t, P])
    dp_conv2d_tc_grid = (c_int, [c_int, c_int, c_int, c_int, c_int, c_int])
    dp_conv2d_tc = (c_int, [P, c_ll, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, P, P, c_ll, P, c_ll, c_int, P,
                            c_ll, P, c_ll, c_int, P, P])
    dp_conv2d_tc_caps = (c_int, [c_int, c_int, c_int, c_int, c_int, c_int])
    dp_conv2d_tc_fused = (c_int, [P, c_ll, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, P, P, c_ll, P, c_ll, c_int,
                                  P, c_ll, P, c_ll, c_int, P, P, P])
    dp_conv2d_wgrad_tc_fused = (c_int, [P, c_ll, P, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, P, c_size_t,
                                        P, c_int, P])
    dp_conv2d_tc_down2_grid = (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int])
    dp_conv2d_tc_down2 = (c_int, [P, c_ll, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, c_int, P, c_int, P, c_ll,
                                  c_int, c_int, P, P])
    dp_conv2d_tc_up2 = (c_int, [P, c_ll, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, c_int, P, c_int, P, c_ll,
                                c_int, c_int, P])
    dp_conv2d_wgrad_tc_workspace = (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int])
    dp_conv2d_wgrad_tc = (c_int, [P, c_ll, P, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, P, c_size_t, P])
    dp_dwconv_fwd_blocks = (c_int, [c_int, c_int, c_int, c_int])
    dp_dwconv_fwd_blocks_s = (c_int, [c_int, c_int, c_int, c_int, c_int, c_int])
    dp_dwconv_fwd = (c_int, [P, c_ll, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, c_int, P, c_ll, c_int, c_int, P, P])
    dp_dwconv_dgrad_s2 = (c_int, [P, c_ll, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, P, c_ll, c_int, c_int, P])
    dp_dwconv_wgrad_workspace = (c_size_t, [c_int, c_int, c_int, c_int, c_int])
    dp_dwconv_wgrad = (c_int, [P, c_ll, c_int, c_int, c_int, c_int, P, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, P,
                               c_int, P, c_size_t, P])
    dp_nchw_f32_to_nhwc_bf16 = (c_int, [P, c_int, c_int, c_int, c_int, P, c_ll, P])
    dp_nhwc_bf16_to_nchw_f32 = (c_int, [P, c_ll, c_int, c_int, c_int, c_int, P, P])
    dp_cast_f32_to_bf16 = (c_int, [P, P, c_size_t, P])
    dp_cast_bf16_to_f32 = (c_int, [P, P, c_size_t, P])
    dp_pack_conv_weight = (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, P])
    dp_pack_conv_weights_batched = (c_int, [P, c_int, c_int, P])
    dp_add_relu_bwd = (c_int, [P, P, P, P, c_size_t, P])
    dp_add_bf16 = (c_int, [P, P, P, P, c_size_t, P])
    dp_relu_bf16 = (c_int, [P, P, c_size_t, P])
    dp_copy_channels = (c_int, [P, c_ll, P, c_ll, c_size_t, c_int, P])
    dp_resize_bilinear_nhwc = (c_int, [P, c_ll, c_int, c_int, c_int, c_int, P, c_ll, c_int, c_int, c_int, P])
    dp_resize_bilinear_nhwc_bwd = (c_int, [P, c_ll, c_int, c_int, c_int, c_int, P, c_ll, c_int, c_int, c_int, P])
    dp_resize_bilinear_planes_f32 = (c_int, [P, c_int, c_int, c_int, P, c_int, c_int, c_int, P])
    dp_chan_reduce_blocks = (c_int, [])
    dp_chan_reduce = (c_int, [c_int, P, c_ll, P, c_ll, P, c_ll, P, c_size_t, c_int, P, P])
    dp_sum_partials = (c_int, [P, c_int, c_int, c_int, P, c_int, P])
    dp_bn_finalize = (c_int, [P, c_int, c_int, c_double, P, P, c_float, c_float, P, P, P, P, P, P])
    dp_bn_eval_coeffs = (c_int, [P, P, P, P, c_float, c_int, P, P, P])
    dp_bn_apply = (c_int, [P, c_ll, P, P, c_ll, P, P, c_ll, c_size_t, c_int, c_int, P, c_ll, P])
    dp_bn_bwd_apply = (c_int, [P, c_ll, P, c_ll, P, P, c_ll, P, P, P, c_double, c_int, c_size_t, c_int, P, c_ll, P, c_ll,
                               P, P, c_int, P])
    dp_conv_gather = (c_int, [P, c_ll, c_int, c_int, c_int, c_int, P, P, P, c_ll, c_int, c_int, c_int, c_int, c_int,
                              c_int, c_int, c_int, c_int, P])
    dp_conv_wgrad_direct_workspace = (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int])
    dp_conv_wgrad_direct = (c_int, [P, c_ll, c_int, c_int, c_int, P, c_ll, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_int, c_int, c_int, P, c_int, P, c_size_t, P])
    dp_head_conv_fwd = (c_int, [P, c_ll, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, P])
    dp_head_conv_bwd_workspace = (c_size_t, [c_int, c_int])
    dp_head_conv_bwd = (c_int, [P, P, c_int, P, c_ll, c_int, c_int, c_int, c_int, c_int, P, P, c_ll, P, P, c_int, P,
                                c_size_t, P])
    dp_lnl_blocks = (c_int, [])
    dp_lnl_partial_floats = (c_int, [])
    dp_ln_linear_fwd = (c_int, [P, c_ll, c_int, c_size_t, c_int, P, P, c_float, P, P, P, c_ll, c_int, P])
    dp_ln_linear_bwd = (c_int, [P, c_ll, c_int, c_size_t, c_int, P, P, c_float, P, P, c_ll, c_int, P, c_ll, P, P, P, P,
                                P, c_int, P])
    dp_attn_fwd = (c_int, [P, P, P, c_int, c_int, c_float, P, c_int, P, P, P])
    dp_attn_bwd = (c_int, [P, P, P, P, P, P, c_int, c_int, c_float, P, c_int, P, c_int, P, P, P, P, P])
    dp_conv2d_wgrad_tc_s2_workspace = (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int])
    dp_conv2d_wgrad_tc_s2 = (c_int, [P, c_ll, c_int, c_int, c_int, P, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, P,
                                     c_int, P, c_size_t, P])
    dp_debug_set_buffer = (None, [P])



class ConvFuse(ctypes.Structure):
    """dp_conv_fuse_t (include/depth_b200.h)"""
    _fields_ = [("pre_scale_shift", c_void_p), ("pre_act", c_int), ("mask_x", c_void_p), ("mask_ld", c_ll),
                ("mask_scale_shift", c_void_p), ("mask_act", c_int)]


CAP_PROLOGUE, CAP_BN_BACKWARD = 1, 2


def check(code):
    if code != 0:
        raise DepthB200Error(f"libdepth_b200 error {code}: {lib().dp_last_error().decode()}")


def ptr(t):
    """device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise DepthB200Error("depth_b200 kernels need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def build_id():
    """source hash of the library that was built (build.py writes it next to the .so)"""
    try:
        with open(os.path.join(_HERE, "build_id.txt")) as f:
            return f.read().strip()
    except OSError:
        return "unknown"


def build_files():
    """per-file source hashes of the library that was built"""
    import json
    try:
        with open(os.path.join(_HERE, "build_files.json")) as f:
            return json.load(f)
    except OSError:
        return {}


def launch_count():
    return int(lib().dp_launch_count())


# flags (include/depth_b200.h)
F_SI, F_SILOG, F_GRAD, F_EDGE, F_ABSREL, F_M4 = 1, 2, 4, 8, 16, 32
NMOM, NLOSS, MAX_THR = 16, 8, 8
M_S1, M_S2, M_M0, M_M1, M_M2, M_GX, M_GY, M_EX, M_EY, M_AR, M_AB, M_SQ, M_V0, M_V1, M_V2 = range(15)
L_TOTAL, L_SI, L_SILOG, L_GRAD, L_EDGE, L_ABSREL, L_SI_RAW, L_SILOG_RAW = range(8)
