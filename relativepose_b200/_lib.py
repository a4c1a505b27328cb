"""ctypes binding of librp_b200.so (the C ABI in include/rp_b200.h).

There is NO CPU fallback: if the library is missing or no CUDA device is usable,
the product path raises.  (The numpy oracle under oracle/ is test infrastructure
and is never imported from here.)
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RP_B200_LIB", os.path.join(HERE, "librp_b200.so"))   # env override: tuning variants only

RP_OK = 0
ERRORS = {-1: "RP_ERR_INVALID_ARG", -2: "RP_ERR_WORKSPACE_TOO_SMALL", -3: "RP_ERR_UNSUPPORTED",
          -4: "RP_ERR_CUDA", -5: "RP_ERR_NO_DEVICE"}

STAGE_TOPK, STAGE_AFFINITY, STAGE_SOLVE = 1, 2, 3
METHODS = {"horn87": 0, "spectral": 1, "irls": 2, "irls+sm": 3}
MAX_TOPK = 8
STATS_STRIDE = 8
STATUS_EDGE_OVERFLOW = -2
STATUS_UNSUPPORTED = -3


class RpParams(ctypes.Structure):
    _fields_ = [("feat_den", ctypes.c_double), ("feat_den_obs", ctypes.c_double),
                ("dist_thre_sq", ctypes.c_double), ("sep_thre", ctypes.c_double),
                ("angle_thre_sq", ctypes.c_double), ("den_dist", ctypes.c_double),
                ("den_a1", ctypes.c_double), ("den_a2", ctypes.c_double),
                ("mu", ctypes.c_double), ("power_tol", ctypes.c_double),
                ("topk", ctypes.c_int32), ("method", ctypes.c_int32),
                ("max_power_iters", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class RpDebug(ctypes.Structure):
    _fields_ = [("topk_idx", ctypes.c_void_p), ("topk_f", ctypes.c_void_p),
                ("dij", ctypes.c_void_p), ("dij_off", ctypes.c_void_p),
                ("edge_rc", ctypes.c_void_p), ("edge_w", ctypes.c_void_p), ("edge_cap", ctypes.c_int64),
                ("u", ctypes.c_void_p), ("u_stride", ctypes.c_int64), ("phase_clk", ctypes.c_void_p)]


class RpConvSrc(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("pitch", ctypes.c_int32), ("ch_off", ctypes.c_int32),
                ("C", ctypes.c_int32), ("act", ctypes.c_int32), ("scale", ctypes.c_void_p),
                ("shift", ctypes.c_void_p), ("sstride", ctypes.c_int32), ("s_off", ctypes.c_int32),
                ("slope", ctypes.c_float), ("dtype", ctypes.c_int32)]


class RpConvDesc(ctypes.Structure):
    _fields_ = [("src", RpConvSrc * 2), ("nsrc", ctypes.c_int32), ("transposed", ctypes.c_int32),
                ("k", ctypes.c_int32), ("s", ctypes.c_int32), ("p", ctypes.c_int32), ("G", ctypes.c_int32),
                ("Hin", ctypes.c_int32), ("Win", ctypes.c_int32), ("Hout", ctypes.c_int32), ("Wout", ctypes.c_int32),
                ("Cout", ctypes.c_int32), ("W", ctypes.c_void_p), ("out", ctypes.c_void_p),
                ("out_pitch", ctypes.c_int32), ("out_ch_off", ctypes.c_int32), ("psum", ctypes.c_void_p),
                ("psq", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("tanh_out", ctypes.c_int32),
                ("imgs_per_group", ctypes.c_int32), ("out_dtype", ctypes.c_int32)]


class RpNetOp(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("reserved", ctypes.c_int32), ("conv", RpConvDesc), ("arg", ctypes.c_uint64 * 16)]


# op kinds of rp_scnet_forward / rp_resnet18_8s_forward (include/rp_b200.h: RP_OP_*), keyed by the layer entry point
NET_OPS = {"rp_conv_layer": 1, "rp_conv_layer_halo": 3, "rp_bn_finalize": 4, "rp_bn_finalize_split": 5,
           "rp_scnet_resize_in": 6, "rp_scnet_resize_in_split": 7, "rp_scnet_resize_out_map": 8, "rp_im2col_bf16": 9,
           "rp_bn_relu_maxpool": 10, "rp_bn_add_relu": 11, "rp_resize_nhwc": 12, "rp_resize_to_nchw": 13, "rp_space_to_depth_h16": 14, "rp_scnet_resize_out_sub": 15}

EXPORTS = ("rp_abi_version", "rp_device_info", "rp_solve_workspace_bytes", "rp_solve_batch",
           "rp_solve_batch_ex", "rp_solve_default_slots", "rp_solver_wide_max", "rp_solve_pair_host", "rp_h16_format", "rp_match_topk", "rp_launch_count", "rp_spectral_irls_solve", "rp_spectral_irls_workspace_bytes",
           "rp_conv_nparts", "rp_conv_layer", "rp_bn_finalize", "rp_bn_finalize_split", "rp_im2col_bf16", "rp_space_to_depth_h16", "rp_scnet_resize_in", "rp_scnet_resize_in_split", "rp_scnet_resize_out", "rp_scnet_resize_out_map", "rp_scnet_resize_out_sub",
           "rp_conv_launch_count", "rp_tc_gemm_test",
           "rp_conv_halo_plan", "rp_conv_halo_fits", "rp_conv_layer_halo", "rp_conv_halo_debug", "rp_conv_halo_prof", "rp_conv_halo_tma_count",
           "rp_affinity_build", "rp_scnet_forward", "rp_resnet18_8s_forward", "rp_gather_primitives", "rp_match_sample_workspace_bytes", "rp_match_sample", "rp_heat_sample", "rp_warp_workspace_bytes", "rp_warp_views", "rp_warp_views_ex", "rp_pano2pc", "rp_blend_completion",
           "rp_bn_relu_maxpool", "rp_bn_add_relu", "rp_resize_nhwc", "rp_resize_to_nchw", "rp_interpolate")

_lib = None


def h16_is_fp16():
    """True when the library was built with IEEE-half tensor-core operands (csrc/rp_h16.cuh), False for bfloat16."""
    return load().rp_h16_format() == 1


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    """Load the CUDA library or raise (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            "%s not found: build it with `python -m relativepose_b200.build` (nvcc, sm_100a). "
            "relativepose_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.rp_abi_version.restype = i32
    lib.rp_launch_count.restype = i64
    lib.rp_device_info.restype = i32
    lib.rp_device_info.argtypes = [ctypes.POINTER(ctypes.c_int)] * 3 + [ctypes.POINTER(ctypes.c_size_t)]
    lib.rp_solve_workspace_bytes.restype = i32
    lib.rp_h16_format.restype = i32
    lib.rp_conv_halo_tma_count.restype = i64
    lib.rp_solve_default_slots.restype = i32
    lib.rp_solver_wide_max.restype = i32
    lib.rp_solve_pair_host.restype = i32
    lib.rp_solve_pair_host.argtypes = [i32, i32] + [vp] * 8 + [i32, vp, vp, i32, i32, i64, vp, vp, vp, vp]
    lib.rp_solver_wide_max.argtypes = [i32]
    lib.rp_solve_default_slots.argtypes = [i32, i32, i32, i32, ctypes.POINTER(ctypes.c_int)]
    lib.rp_solve_workspace_bytes.argtypes = [i32, i32, i32, i32, i32, i64, ctypes.POINTER(ctypes.c_size_t)]
    common = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i64, vp,
              ctypes.c_size_t, vp, vp, vp]
    lib.rp_solve_batch.restype = i32
    lib.rp_solve_batch.argtypes = common + [vp]
    lib.rp_solve_batch_ex.restype = i32
    lib.rp_solve_batch_ex.argtypes = common + [i32, ctypes.POINTER(RpDebug), vp]
    lib.rp_match_topk.restype = i32
    lib.rp_match_topk.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, vp,
                                  ctypes.c_size_t, vp, vp, vp, vp]
    lib.rp_spectral_irls_workspace_bytes.restype = i32
    lib.rp_spectral_irls_workspace_bytes.argtypes = [i32, i32, i64, ctypes.POINTER(ctypes.c_size_t)]
    lib.rp_spectral_irls_solve.restype = i32
    lib.rp_spectral_irls_solve.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i64, vp, ctypes.c_size_t, vp, vp, vp, vp]
    lib.rp_conv_nparts.restype = i32
    lib.rp_conv_nparts.argtypes = [ctypes.POINTER(RpConvDesc), ctypes.POINTER(ctypes.c_int)]
    lib.rp_conv_layer.restype = i32
    lib.rp_conv_layer.argtypes = [ctypes.POINTER(RpConvDesc), vp]
    lib.rp_bn_finalize.restype = i32
    lib.rp_bn_finalize.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, i32, i32, vp]
    lib.rp_bn_finalize_split.restype = i32
    lib.rp_bn_finalize_split.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp]
    lib.rp_space_to_depth_h16.restype = i32
    lib.rp_space_to_depth_h16.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp]
    lib.rp_conv_halo_fits.restype = i32
    lib.rp_conv_halo_fits.argtypes = [ctypes.POINTER(RpConvDesc), i32, i32, i32]
    lib.rp_im2col_bf16.restype = i32
    lib.rp_im2col_bf16.argtypes = [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.rp_scnet_resize_in.restype = i32
    lib.rp_scnet_resize_in.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.rp_scnet_resize_in_split.restype = i32
    lib.rp_scnet_resize_in_split.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.rp_scnet_resize_out.restype = i32
    lib.rp_scnet_resize_out.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    lib.rp_scnet_resize_out_sub.restype = i32
    lib.rp_scnet_resize_out_sub.argtypes = [vp, i32, i32, vp, vp, i32, i32, i32, vp, i32, vp]
    lib.rp_scnet_resize_out_map.restype = i32
    lib.rp_scnet_resize_out_map.argtypes = [vp, i32, i32, vp, i32, i32, i32, vp, vp]
    lib.rp_conv_launch_count.restype = i64
    lib.rp_conv_halo_plan.restype = i32
    lib.rp_conv_halo_plan.argtypes = [ctypes.POINTER(RpConvDesc), i32, i32, i32, ctypes.POINTER(ctypes.c_int),
                                      ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    lib.rp_conv_halo_debug.restype = i32
    lib.rp_conv_halo_debug.argtypes = [ctypes.POINTER(RpConvDesc), i32, i32, i32, ctypes.POINTER(ctypes.c_int)]
    lib.rp_conv_layer_halo.restype = i32
    lib.rp_conv_layer_halo.argtypes = [ctypes.POINTER(RpConvDesc), vp, i32, i32, i32, vp]
    lib.rp_bn_relu_maxpool.restype = i32
    lib.rp_bn_relu_maxpool.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, i32, i32, vp]
    lib.rp_bn_add_relu.restype = i32
    lib.rp_bn_add_relu.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.rp_resize_nhwc.restype = i32
    lib.rp_resize_nhwc.argtypes = [vp, i32, i32, i32, i32, vp, i32, i32, i32, vp]
    lib.rp_resize_to_nchw.restype = i32
    lib.rp_resize_to_nchw.argtypes = [vp, i32, i32, i32, i32, vp, i32, i32, i32, vp]
    lib.rp_interpolate.restype = i32
    lib.rp_interpolate.argtypes = [vp, i32, i32, i32, vp, i32, vp, vp]
    lib.rp_warp_workspace_bytes.restype = i32
    lib.rp_warp_workspace_bytes.argtypes = [i32, ctypes.POINTER(ctypes.c_size_t)]
    lib.rp_warp_views.restype = i32
    lib.rp_warp_views.argtypes = [vp, vp, i32, i32, vp, vp, ctypes.c_size_t, vp]
    lib.rp_warp_views_ex.restype = i32
    lib.rp_warp_views_ex.argtypes = [vp, ctypes.c_longlong, vp, vp, i32, i32, vp, ctypes.c_longlong, vp, ctypes.c_size_t, vp]
    lib.rp_pano2pc.restype = i32
    lib.rp_pano2pc.argtypes = [vp, i32, i32, vp, vp, vp]
    lib.rp_blend_completion.restype = i32
    lib.rp_blend_completion.argtypes = [vp, i32, vp, vp, vp, i32, i32, vp, vp, vp]
    lib.rp_match_sample.restype = i32
    lib.rp_match_sample_workspace_bytes.restype = i32
    lib.rp_match_sample_workspace_bytes.argtypes = [i32, ctypes.POINTER(ctypes.c_size_t)]
    lib.rp_match_sample.argtypes = [vp, i32, i32, vp, i32, i32, i32, i32, vp, vp, ctypes.c_size_t, vp]
    lib.rp_heat_sample.restype = i32
    lib.rp_heat_sample.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, ctypes.c_size_t, vp]
    lib.rp_gather_primitives.restype = i32
    lib.rp_gather_primitives.argtypes = [vp, i32, ctypes.c_longlong, vp, vp, vp, i32, i32, i32, vp, vp, vp, vp]
    lib.rp_affinity_build.restype = i32
    lib.rp_affinity_build.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, i32, i32, i32, i64, vp,
                                      ctypes.c_size_t, vp, vp, vp, vp, vp, vp]
    lib.rp_scnet_forward.restype = i32
    lib.rp_scnet_forward.argtypes = [ctypes.POINTER(RpNetOp), i32, vp]
    lib.rp_resnet18_8s_forward.restype = i32
    lib.rp_resnet18_8s_forward.argtypes = [ctypes.POINTER(RpNetOp), i32, vp]
    lib.rp_tc_gemm_test.restype = i32
    lib.rp_tc_gemm_test.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    if lib.rp_abi_version() != 1:
        raise RuntimeError("librp_b200.so ABI mismatch")
    _lib = lib
    return lib


def check(rc, what):
    if rc != RP_OK:
        raise RuntimeError("%s failed: %s (%d)" % (what, ERRORS.get(rc, "?"), rc))
