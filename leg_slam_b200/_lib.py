"""ctypes binding of liblgs.so (include/lgs.h).  No fallback: if the CUDA library is missing
or fails to load, importing a compute entry point raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblgs.so")

c_void_p, c_int, c_float, c_double, c_size_t, c_int64 = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t, ctypes.c_int64)


class LgsError(RuntimeError):
    pass


class BinningView(ctypes.Structure):
    _fields_ = [("point_list", c_void_p), ("keys_sorted32", c_void_p)]


class ReferenceKeysView(ctypes.Structure):
    _fields_ = [("keys_unsorted", c_void_p), ("values_unsorted", c_void_p), ("keys_sorted", c_void_p),
                ("point_offsets", c_void_p)]


class ImageView(ctypes.Structure):
    _fields_ = [("ranges", c_void_p), ("final_T", c_void_p), ("n_contrib", c_void_p)]


class GeomView(ctypes.Structure):
    _fields_ = [("records", c_void_p), ("cov3D", c_void_p), ("tiles_touched", c_void_p),
                ("internal_radii", c_void_p), ("clamped", c_void_p)]


# symbol -> (restype, argtypes); the CPU test-suite checks every one of these is exported
SIGNATURES = {
    "lgs_status_string": (ctypes.c_char_p, [c_int]),
    "lgs_last_cuda_error": (c_int, []),
    "lgs_abi_version": (c_int, []),
    "lgs_geom_bytes": (c_size_t, [c_int]),
    "lgs_image_bytes": (c_size_t, [c_int, c_int]),
    "lgs_binning_bytes": (c_size_t, [c_int]),
    "lgs_forward_stage1": (c_int, [c_int, c_int, c_int, c_int, c_int] + [c_void_p] * 5 + [c_float] +
                           [c_void_p] * 5 + [c_float, c_float, c_int] + [c_void_p] * 4),
    "lgs_forward_stage1_split_sh": (c_int, [c_int, c_int, c_int, c_int, c_int] + [c_void_p] * 5 + [c_float] +
                                    [c_void_p] * 5 + [c_float, c_float, c_int] + [c_void_p] * 4),
    "lgs_forward_stage2": (c_int, [c_int, c_int, c_int, c_int] + [c_void_p] * 8 + [c_int, c_void_p]),
    "lgs_backward": (c_int, [c_int] * 6 + [c_void_p] * 6 + [c_float] + [c_void_p] * 5 + [c_float, c_float] +
                     [c_void_p] * 18 + [c_int, c_int, c_void_p, c_void_p]),
    "lgs_backward_split_sh": (c_int, [c_int] * 6 + [c_void_p] * 6 + [c_float] + [c_void_p] * 5 + [c_float, c_float] +
                              [c_void_p] * 17 + [c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "lgs_backward_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "lgs_mark_visible": (c_int, [c_int] + [c_void_p] * 5),
    "lgs_view_binning": (c_int, [c_void_p, c_int, ctypes.POINTER(BinningView)]),
    "lgs_view_image": (c_int, [c_void_p, c_int, c_int, ctypes.POINTER(ImageView)]),
    "lgs_view_geom": (c_int, [c_void_p, c_int, ctypes.POINTER(GeomView)]),
    "lgs_debug_keys_bytes": (c_size_t, [c_int, c_int]),
    "lgs_debug_reference_keys": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         ctypes.POINTER(ReferenceKeysView), c_void_p]),
    "lgs_forward_status": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "lgs_profile_enable": (c_int, [c_int]),
    "lgs_profile_read": (c_int, [ctypes.POINTER(c_float), c_int]),
    "lgs_bench_fma": (c_int, [c_int, c_int, c_void_p, c_void_p]),
    "lgs_adam_multi": (c_int, [c_int] + [c_void_p] * 6 + [c_double, c_double, c_double, c_int, c_void_p]),
    "lgs_dp_adam_shard": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int] + [c_void_p] * 4 + [c_int64, c_int64, c_void_p,
                                  c_void_p, c_double, c_double, c_double, c_int, c_int, c_void_p]),
    "lgs_dp_adam_shard_sparse": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                         c_int64, c_int64, c_void_p, c_void_p, c_double, c_double, c_double, c_int, c_int, c_void_p]),
    "lgs_dp_rows_mark": (c_int, [c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "lgs_dp_rows_table_bytes": (c_size_t, [c_int, c_int]),
    "lgs_dp_rows_publish": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "lgs_dp_rows_combine": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "lgs_stream_hooks": (c_int, [c_void_p, c_void_p]),
    "lgs_activations_fwd": (c_int, [c_int, c_int] + [c_void_p] * 10),
    "lgs_activations_bwd": (c_int, [c_int, c_int, c_int] + [c_void_p] * 13),
    "lgs_mapping_loss_scratch_bytes": (c_size_t, [c_int, c_int]),
    "lgs_mapping_loss": (c_int, [c_int] * 4 + [c_void_p] * 7 + [c_float, c_int] + [c_void_p] * 6),
    "lgs_densify_stats": (c_int, [c_int] + [c_void_p] * 6),
    "lgs_densify_plan_bytes": (c_size_t, [c_int]),
    "lgs_densify_plan": (c_int, [c_int] + [c_void_p] * 4 + [c_float] * 4 + [c_int, c_void_p, c_void_p, c_void_p]),
    "lgs_densify_apply": (c_int, [c_int, c_void_p, c_void_p, c_int] + [c_void_p] * 9),
    "lgs_reproject_depth_pinhole": (c_int, [c_int, c_int] + [c_float] * 4 + [c_void_p] * 4),
    "lgs_transform_points": (c_int, [c_int] + [c_void_p] * 4),
    "lgs_knn_scratch_bytes": (c_size_t, [c_int]),
    "lgs_knn_mean_dist2": (c_int, [c_int] + [c_void_p] * 4),
    "lgs_scale_transform_mark_visible": (c_int, [c_int, c_float] + [c_void_p] * 6 + [c_int, c_void_p, c_void_p]),
    "lgs_inactive_geo_scratch_bytes": (c_size_t, [c_int]),
    "lgs_inactive_geo_densify": (c_int, [c_int, c_int] + [c_float] * 5 + [c_void_p] * 4 + [ctypes.c_longlong] + [c_void_p] * 5),
    "lgs_ply_pack": (c_int, [ctypes.c_longlong, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lgs_ply_unpack": (c_int, [ctypes.c_longlong, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lgs_cosine_query": (c_int, [c_int, c_int] + [c_void_p] * 4),
    "lgs_cosine_query_simt": (c_int, [c_int, c_int] + [c_void_p] * 4),
    "lgs_minmax_invert": (c_int, [c_int64] + [c_void_p] * 3),
    "lgs_cosine_image": (c_int, [c_int64, c_int] + [c_void_p] * 4),
    "lgs_heat_colors": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p]),
}

_lib = None


def lib():
    """The loaded library; raises LgsError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LgsError(
                f"{LIB_PATH} not found: build it with `python -m leg_slam_b200.build` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        try:
            L = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise LgsError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status, what):
    if status != 0:
        L = lib()
        msg = L.lgs_status_string(status).decode()
        raise LgsError(f"{what}: {msg} (status {status}, cudaError {L.lgs_last_cuda_error()})")


def ptr(t):
    """Device pointer of a tensor, NULL for None / empty tensors (the reference's
    `torch::tensor({})` sentinel whose data_ptr is nullptr, gaussian_rasterizer.cpp:207-218)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()
