"""ctypes binding of ``libdns_slam_b200.so`` (C ABI declared in ``include/dns_slam_b200.h``).

The library is a plain nvcc-built shared object living next to this file; it is loaded with
``ctypes`` (no torch extension machinery at run time).  Tensors cross the boundary as raw
device pointers (``tensor.data_ptr()``) plus the current CUDA stream.  There is no CPU
fallback: if the library is missing, or a tensor is not a contiguous CUDA tensor of the
expected dtype, the call raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DNS_SLAM_B200_LIB") or os.path.join(_HERE, "libdns_slam_b200.so")   # override: scratch builds only

MAX_LEVELS = 16
MODE_TRACK, MODE_MAP = 0, 1


class Grid(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("n_features", C.c_int32),
                ("scale", C.c_float * MAX_LEVELS), ("res", C.c_uint32 * MAX_LEVELS),
                ("size", C.c_uint32 * MAX_LEVELS), ("offset", C.c_uint32 * (MAX_LEVELS + 1)),
                ("hashed", C.c_uint32 * MAX_LEVELS)]


_P = C.c_void_p


class RenderArgs(C.Structure):
    _fields_ = [("mode", C.c_int32), ("n_rays", C.c_int32), ("n_samples", C.c_int32),
                ("n_class", C.c_int32), ("n_experts", C.c_int32), ("need_dparams", C.c_int32),
                ("need_drays", C.c_int32), ("need_dfeat", C.c_int32),
                ("bound", (C.c_double * 2) * 3),
                ("lambda_p", C.c_float), ("lambda_d", C.c_float), ("lambda_l", C.c_float),
                ("lambda_lt", C.c_float), ("lambda_fs", C.c_float), ("lambda_op", C.c_float),
                ("opacity_trunc", C.c_float), ("opacity_sigma", C.c_float),
                ("grid", Grid),
                ("rays_o", _P), ("rays_d", _P), ("z_vals", _P), ("gt_color", _P), ("gt_depth", _P),
                ("gt_label", _P), ("mask", _P), ("features", _P),
                ("table", _P), ("coarse", _P), ("color", _P), ("logit", _P), ("experts", _P),
                ("class_to_expert", _P), ("n_class_ids", C.c_int32),
                ("pred_color", _P), ("pred_depth", _P), ("pred_var", _P), ("pred_logits", _P),
                ("fine", _P), ("coarse_out", _P), ("losses", _P),
                ("d_table", _P), ("d_coarse", _P), ("d_color", _P), ("d_logit", _P),
                ("d_experts", _P), ("d_rays_o", _P), ("d_rays_d", _P), ("d_features", _P),
                ("workspace", _P), ("workspace_bytes", C.c_int64),
                ("n_rays_total", C.c_int64), ("ray_offset", C.c_int64), ("gt_label_all", _P),
                ("global_counts", _P), ("forward_only", C.c_int32), ("use_simt", C.c_int32),
                ("features_band_only", C.c_int32), ("reserved_", C.c_int32)]


class TvArgs(C.Structure):
    _fields_ = [("n", C.c_int32), ("smooth_pts", C.c_int32), ("voxel", C.c_double),
                ("bound", (C.c_double * 2) * 3), ("offset", C.c_double * 3),
                ("jitter", C.c_double * 3), ("lambda_sm", C.c_float), ("need_dparams", C.c_int32),
                ("grid", Grid), ("table", _P), ("coarse", _P), ("loss", _P), ("d_table", _P),
                ("d_coarse", _P), ("workspace", _P), ("workspace_bytes", C.c_int64), ("offset_jitter_dev", _P),
                ("use_simt", C.c_int32), ("reserved_", C.c_int32)]


class SampleArgs(C.Structure):
    _fields_ = [("n", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("H0", C.c_int32),
                ("W0", C.c_int32), ("Ww", C.c_int32), ("n_uniform", C.c_int32),
                ("n_surface", C.c_int32), ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float),
                ("cy", C.c_float), ("bound", (C.c_double * 2) * 3),
                ("color", _P), ("depth", _P), ("label", _P), ("index", _P), ("R", _P), ("T", _P),
                ("t_lin", _P), ("t_surface", _P), ("t_zero", _P),
                ("gt_color", _P), ("gt_depth", _P), ("gt_label", _P), ("rays_o", _P),
                ("rays_d", _P), ("z_vals", _P), ("pts", _P), ("inside", _P), ("scratch", _P),
                ("order", _P), ("slot_base", _P), ("n_direct", C.c_int32), ("phase", C.c_int32), ("pixel", _P)]


MAX_FRAMES = 8


class FeatMergeArgs(C.Structure):
    _fields_ = [("n_rays", C.c_int32), ("n_samples", C.c_int32), ("n_frames", C.c_int32), ("n_views", C.c_int32),
                ("ray_start", C.c_int32 * (MAX_FRAMES + 1)), ("H", C.c_int32), ("W", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("apply_trunc", C.c_int32), ("need_dparams", C.c_int32), ("need_drays", C.c_int32),
                ("no_zero_fill", C.c_int32), ("bound", (C.c_double * 2) * 3), ("K", _P), ("w2c", _P), ("cam_o", _P), ("feats", _P * MAX_FRAMES),
                ("rays_o", _P), ("rays_d", _P), ("z_vals", _P), ("gt_depth", _P), ("params", _P), ("features", _P),
                ("d_features", _P), ("d_params", _P), ("d_rays_o", _P), ("d_rays_d", _P), ("workspace", _P),
                ("workspace_bytes", C.c_int64), ("stash", _P), ("stash_bytes", C.c_int64)]


_lib = None

# every symbol include/dns_slam_b200.h declares
SYMBOLS = ["dns_last_error", "dns_version", "dns_struct_sizes", "dns_profile_enable", "dns_profile_read", "dns_debug_gemm_tc", "dns_debug_gemm_fmt", "dns_debug_gemm_img", "dns_oneblob_fwd", "dns_oneblob_bwd",
           "dns_hashgrid_fwd", "dns_hashgrid_bwd", "dns_hashgrid_indices", "dns_mlp_fwd",
           "dns_mlp_bwd", "dns_render_workspace_bytes", "dns_render_fwd_bwd", "dns_render_counts",
           "dns_tv_workspace_bytes", "dns_tv_fwd_bwd", "dns_sample_rays", "dns_feature_gather",
           "dns_adam_step", "dns_merge_workspace_bytes", "dns_merge_fwd", "dns_merge_bwd",
           "dns_stem_workspace_bytes", "dns_stem_fwd", "dns_adam_multi", "dns_featmerge_workspace_bytes",
           "dns_featmerge_fwd", "dns_featmerge_bwd", "dns_pose_prepare", "dns_pose_grad",
           "dns_class_tables_workspace_bytes", "dns_class_tables", "dns_track_best", "dns_map_step_result", "dns_sample_rays_batch"]


def lib():
    """Loads the shared library once; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C dns_slam_b200/csrc`). dns_slam_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.dns_last_error.restype = C.c_char_p
    L.dns_render_workspace_bytes.restype = C.c_int64
    L.dns_tv_workspace_bytes.restype = C.c_int64
    L.dns_merge_workspace_bytes.restype = C.c_int64
    i64, i32, f32 = C.c_int64, C.c_int, C.c_float
    L.dns_struct_sizes.argtypes = [C.POINTER(C.c_int64)]
    L.dns_profile_enable.argtypes = [C.c_int]
    L.dns_debug_gemm_tc.argtypes = [_P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int64, _P, _P]
    L.dns_debug_gemm_fmt.argtypes = [_P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, _P, _P]
    L.dns_debug_gemm_img.argtypes = [_P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int64, C.c_int, _P, _P]
    L.dns_profile_read.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]
    L.dns_oneblob_fwd.argtypes = [_P, i64, i32, i32, _P, _P]
    L.dns_oneblob_bwd.argtypes = [_P, _P, i64, i32, i32, _P, _P]
    L.dns_hashgrid_fwd.argtypes = [C.POINTER(Grid), _P, _P, i64, _P, _P]
    L.dns_hashgrid_bwd.argtypes = [C.POINTER(Grid), _P, _P, _P, i64, _P, _P, _P]
    L.dns_hashgrid_indices.argtypes = [C.POINTER(Grid), _P, i64, _P, _P]
    L.dns_mlp_fwd.argtypes = [_P, _P, i64, i32, i32, _P, _P, _P]
    L.dns_mlp_bwd.argtypes = [_P, _P, _P, _P, i64, i32, i32, _P, _P, _P, _P]
    L.dns_render_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
    L.dns_render_fwd_bwd.argtypes = [C.POINTER(RenderArgs), _P]
    L.dns_render_counts.argtypes = [C.POINTER(RenderArgs), _P, _P]
    L.dns_tv_workspace_bytes.argtypes = [i32]
    L.dns_tv_fwd_bwd.argtypes = [C.POINTER(TvArgs), _P]
    L.dns_sample_rays.argtypes = [C.POINTER(SampleArgs), _P]
    L.dns_sample_rays_batch.argtypes = [C.POINTER(SampleArgs), i32, _P]
    L.dns_feature_gather.argtypes = [_P, i64, _P, i32, _P, i32, i32, _P, i32, i32, i32, _P, _P, _P, _P]
    L.dns_adam_step.argtypes = [_P, _P, _P, _P, i64, f32, f32, f32, f32, i32, _P]
    bound_t = (C.c_double * 2) * 3
    L.dns_merge_workspace_bytes.argtypes = [i64]
    L.dns_merge_fwd.argtypes = [_P, _P, _P, i64, i32, bound_t, _P, i32, _P, i64, _P]
    L.dns_merge_bwd.argtypes = [_P, _P, i64, i32, bound_t, _P, _P, _P, i64, _P]
    L.dns_adam_multi.argtypes = [_P, i32, i64, _P, f32, f32, f32, _P]
    L.dns_stem_workspace_bytes.restype = C.c_int64
    L.dns_stem_workspace_bytes.argtypes = []
    L.dns_stem_fwd.argtypes = [_P, i32, i32, i32, _P, _P, _P, f32, f32, i32, _P, _P, _P, _P, i64, _P]
    L.dns_featmerge_workspace_bytes.restype = C.c_int64
    L.dns_featmerge_workspace_bytes.argtypes = [i32, i32]
    L.dns_featmerge_fwd.argtypes = [C.POINTER(FeatMergeArgs), _P]
    L.dns_featmerge_bwd.argtypes = [C.POINTER(FeatMergeArgs), _P]
    L.dns_pose_prepare.argtypes = [_P, _P, i32, _P, _P, _P, i32, _P, _P, _P, _P]
    L.dns_pose_grad.argtypes = [_P, _P, _P, i32, C.POINTER(C.c_int32), i32, i32, i32, f32, f32, f32, f32, _P, _P, _P, _P, _P]
    L.dns_track_best.argtypes = [_P, _P, _P, _P, _P, _P, _P, i32, _P, _P]
    L.dns_map_step_result.argtypes = [i32, _P, _P, f32, f32, f32, _P, i32, _P, _P, _P]
    L.dns_class_tables_workspace_bytes.restype = C.c_int64
    L.dns_class_tables_workspace_bytes.argtypes = [i64, i32]
    L.dns_class_tables.argtypes = [_P, i64, i32, _P, _P, _P, _P, _P, i64, _P]
    sizes = (C.c_int64 * 5)()
    L.dns_struct_sizes(sizes)
    mine = [C.sizeof(Grid), C.sizeof(RenderArgs), C.sizeof(TvArgs), C.sizeof(SampleArgs), C.sizeof(FeatMergeArgs)]
    if list(sizes) != mine:
        raise RuntimeError(f"ctypes struct layout {mine} != C layout {list(sizes)}; rebuild the library")
    _lib = L
    return L


PHASES = ("prep", "class_prep", "point_fwd", "ray", "point_bwd", "dw_gemm", "finalize", "adam", "tv_fwd",
          "tv_bwd", "sample", "feature", "ops")


def profile_enable(on):
    lib().dns_profile_enable(int(on))


def profile_read(reset=True):
    """(ms per phase, launches per phase) accumulated since the last reset."""
    ms = (C.c_double * 16)()
    ln = (C.c_longlong * 16)()
    n = lib().dns_profile_read(ms, ln, int(reset))
    return {PHASES[i]: ms[i] for i in range(n)}, {PHASES[i]: ln[i] for i in range(n)}


def check(rc):
    if rc != 0:
        raise RuntimeError("dns_slam_b200: " + lib().dns_last_error().decode())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, dtype=None, allow_none=False):
    """Raw device pointer of a contiguous CUDA tensor (raises otherwise: no host fallback)."""
    if t is None:
        if allow_none:
            return None
        raise ValueError("dns_slam_b200: required tensor is None")
    if not t.is_cuda:
        raise RuntimeError("dns_slam_b200 runs on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("dns_slam_b200: tensor must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"dns_slam_b200: expected {dtype}, got {t.dtype}")
    return t.data_ptr()


def fill_bound(dst, bound):
    """float64 scene bound -> C struct.  The bound never changes after construction, so its host copy is
    cached ON the tensor object (a ``.cpu()`` per call would put a device sync in every iteration; a cache
    keyed by data_ptr could alias a freed tensor)."""
    b = getattr(bound, "_dns_host", None)
    if b is None:
        b = bound.detach().double().cpu().tolist()      # plain floats: indexing a tensor six times per call costs 15 us
        try:
            bound._dns_host = b
        except AttributeError:
            pass
    for a in range(3):
        dst[a][0] = b[a][0]
        dst[a][1] = b[a][1]
