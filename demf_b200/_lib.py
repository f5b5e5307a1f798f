"""ctypes binding of libdemf_b200.so (include/demf_b200.h).

There is no CPU fallback and no alternative backend: if the library is missing or a launch
fails, the caller gets an exception.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libdemf_b200.so")

_c_int = ctypes.c_int
_c_float = ctypes.c_float
_ptr = ctypes.c_void_p

# name -> argtypes ; every function returns int unless listed in _RESTYPES
_SIGNATURES = {
    "demf_version": [],
    "demf_last_error_string": [],
    "demf_launch_count": [],
    "demf_trace_set": [_ptr, _ptr, ctypes.c_uint],
    "demf_fps_workspace_bytes": [_c_int, _c_int, _c_int],
    "demf_fps": [_ptr, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr],
    "demf_fps_grid": [_ptr, _ptr, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr],
    "demf_fps_grid_prefix": [_ptr, _ptr, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _c_int, _ptr],
    "demf_fps_prefix": [_ptr, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr],
    "demf_chain_indices": [_c_int, _c_int] + [_ptr, _c_int] * 4 + [_ptr] * 4 + [_ptr],
    "demf_interp_cat_rows_fwd": [_ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr],
    "demf_decode_boxes": [_ptr, _c_int] * 6 + [_c_int] * 6 + [_ptr, _ptr, _ptr, _ptr],
    "demf_ball_query": [_ptr, _ptr, _c_int, _c_int, _c_int, _c_float, _c_float, _c_int, _ptr, _ptr],
    "demf_group_fwd": [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr],
    "demf_group_bwd": [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr],
    "demf_gather_fwd": [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr],
    "demf_gather_bwd": [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr],
    "demf_query_and_group_fwd": [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float,
                                 _c_int, _c_int, _c_int, _ptr, _ptr, _ptr],
    "demf_three_nn": [_ptr, _ptr, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr],
    "demf_three_interpolate_fwd": [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr],
    "demf_three_interpolate_bwd": [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr],
    "demf_group_rows_width": [_c_int],
    "demf_query_and_group_rows_fwd": [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_float,
                                      _c_float, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr],
    "demf_ball_grid_workspace_bytes": [_c_int, _c_int],
    "demf_ball_grid_build": [_ptr, _c_int, _c_int, _c_float, _ptr, _ptr],
    "demf_ball_query_grid": [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_float, _c_float, _c_int, _ptr,
                             _ptr],
    "demf_group_rows_bwd": [_ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _ptr, _ptr,
                            _ptr, _ptr],
    "demf_three_interpolate_rows_fwd": [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr],
    "demf_three_interpolate_rows_bwd": [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr],
    "demf_sa_pack_floats": [_c_int, _c_int],
    "demf_sa_pack_weights": [_ptr, _c_int, _c_int, _ptr, _ptr],
    "demf_sa_fused_supported": [_c_int] * 5,
    "demf_sa_fused_fwd": [_ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float, _c_int,
                          _c_int, _c_int, _ptr, _ptr, _c_int, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr],
    "demf_sa_fused_error": [],
    "demf_sa_pipe_supported": [_c_int] * 6,
    "demf_sa_pipe_error": [],
    "demf_sa_pipe_fwd": [_ptr, _ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_float, _c_int, _ptr, _ptr, _ptr,
                         _ptr],
    "demf_sa_fused_set_profile": [_ptr],
    "demf_sa_fused_tune": [_c_int, _c_int],
    "demf_msda_fwd": [_ptr, _ptr, _ptr, _ptr, _ptr] + [_c_int] * 7 + [_ptr, _ptr],
    "demf_bn_rows_supported": [_c_int],
    "demf_bn_rows_state_bytes": [_c_int],
    "demf_bn_rows_fwd": [_ptr, ctypes.c_long, _c_int, _ptr, _ptr, _c_float, _c_float, _c_int, _ptr, _ptr, _ptr, _ptr,
                         _ptr, _ptr, _ptr],
    "demf_bn_rows_bwd": [_ptr, _ptr, _ptr, ctypes.c_long, _c_int, _ptr, _ptr, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr,
                         _ptr, _ptr],
    "demf_bn_max_rows_fwd": [_ptr, ctypes.c_long, _c_int, _c_int, _ptr, _ptr, _c_float, _c_float, _ptr, _ptr, _ptr,
                             _ptr, _ptr, _ptr, _ptr, _ptr],
    "demf_bn_max_rows_bwd": [_ptr, _ptr, _ptr, _ptr, ctypes.c_long, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr,
                             _ptr, _ptr, _ptr, _ptr],
    "demf_bn_rows_apply": [_ptr, ctypes.c_long, _c_int, _ptr, _ptr, _ptr, _ptr, _c_int, _ptr, _ptr],
    "demf_bn_max_rows_apply": [_ptr, ctypes.c_long, _c_int, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "demf_mha_supported": [_c_int],
    "demf_mha_fwd": [_ptr, ctypes.c_long, _ptr, ctypes.c_long, _ptr, ctypes.c_long, _ptr, ctypes.c_long, _c_int, _c_int,
                     _c_int, _c_int, _c_int, _c_float, _c_int, _ptr],
    "demf_col_sum_add": [_ptr, ctypes.c_long, _c_int, ctypes.c_long, _ptr, _ptr],
    "demf_bn_rows_bwd_apply": [_ptr, _ptr, ctypes.c_long, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "demf_gemm_rows_dgrad_bn": [_ptr, ctypes.c_long, _ptr, ctypes.c_long, ctypes.c_long, _c_int, _c_int, _ptr,
                                ctypes.c_long, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, ctypes.c_long, _ptr],
    "demf_bn_bwd_finalize": [_ptr, ctypes.c_long, _c_int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "demf_gemm_supported": [_c_int, _c_int],
    "demf_gemm_error": [],
    "demf_gemm_tune": [_c_int],
    "demf_gemm_debug_mn": [_c_int, _c_int, _c_int],
    "demf_gemm_rows_fwd": [_ptr, ctypes.c_long, _ptr, ctypes.c_long, _ptr, ctypes.c_long, _c_int, _c_int, _c_int, _ptr,
                           _ptr, ctypes.c_long, _ptr],
    "demf_gemm_rows_dgrad": [_ptr, ctypes.c_long, _ptr, ctypes.c_long, ctypes.c_long, _c_int, _c_int, _ptr,
                             ctypes.c_long, _ptr],
    "demf_gemm_wgrad": [_ptr, ctypes.c_long, _ptr, ctypes.c_long, ctypes.c_long, _c_int, _c_int, _ptr, _c_int, _ptr],
    "demf_bn_finalize": [_ptr, ctypes.c_long, _c_int, _c_float, _c_float, _ptr, _ptr, _ptr, _ptr, _ptr],
    "demf_stage_loss_fwd": [_ptr] * 14 + [ctypes.c_long, _c_int, _c_int, _ptr, _ptr, _ptr],
    "demf_stage_loss_bwd": [_ptr] * 14 + [ctypes.c_long, _c_int, _c_int, _ptr, _ptr] + [_ptr] * 6 + [_ptr],
    "demf_box_point_count": [_ptr, _c_int, _ptr, _c_int, _c_int, _c_int, _c_int, _ptr, _ptr],
    "demf_nms_select": [_ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float, _ptr, _ptr, _ptr,
                        _ptr, _ptr, _ptr],
    "demf_aligned_3d_nms": [_ptr, _ptr, _ptr, _ptr, _c_int, _c_int, _c_float, _ptr, _ptr],
    "demf_vote_tail": [_ptr, _c_int, _ptr, _ptr, ctypes.c_long, _c_int, _ptr, _c_int, _ptr, _ptr, _ptr, _ptr],
    "demf_project_points": [_ptr, _ptr, _ptr, _c_int, _c_int, _ptr, _ptr],
    "demf_levels_to_rows": [_ptr, _ptr, _c_int, _c_int, _c_int, _ptr, _ptr],
    "demf_bias_layer_norm_rows": [_ptr, _ptr, _ptr, _ptr, _ptr, ctypes.c_long, _c_int, _c_float, _ptr, _ptr, _ptr,
                                  _ptr],
    "demf_msda_proj_fwd_supported": [_c_int] * 3,
    "demf_msda_proj_fwd": [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr] + [_c_int] * 8 + [_ptr, _ptr],
    "demf_msda_bwd": [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr] + [_c_int] * 7 + [_ptr, _ptr, _ptr, _ptr],
}
_RESTYPES = {
    "demf_last_error_string": ctypes.c_char_p,
    "demf_launch_count": ctypes.c_uint64,
    "demf_fps_workspace_bytes": ctypes.c_size_t,
    "demf_sa_pack_floats": ctypes.c_long,
    "demf_ball_grid_workspace_bytes": ctypes.c_size_t,
    "demf_bn_rows_state_bytes": ctypes.c_long,
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class DemfLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library once; raises DemfLibraryError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DemfLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m demf_b200.build` "
            "(there is no CPU or PyTorch fallback for the DeMF hot-path ops)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale: fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, _c_int)
    _lib = lib
    return lib


def check(rc, name):
    if rc != 0:
        msg = load().demf_last_error_string().decode("utf-8", "replace")
        raise DemfLibraryError(f"{name} failed (code {rc}): {msg}")


def launch_count():
    return int(load().demf_launch_count())
