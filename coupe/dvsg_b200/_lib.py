"""ctypes binding of libdvsg_warp.so -- the C ABI declared in include/dvsg_warp.h.

There is deliberately no fallback: if the CUDA library is missing, or a tensor is not on a
CUDA device, the call fails loudly.  Build with `python -m coupe.dvsg_b200._build` (or
`__graft_entry__.build()`); the .so is kept in-tree next to this file.
"""
import ctypes
import os
from ctypes import c_char_p, c_int, c_longlong, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# DVSG_LIB: A/B experiments with an alternative build of the same sources (tools/); never a different implementation
LIB_PATH = os.environ.get('DVSG_LIB') or os.path.join(HERE, 'libdvsg_warp.so')

OK, ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4
FLAG_FORCE_DIRECT = 1
FLAG_TPS_EXACT = 2

_P = c_void_p
# name -> (restype, argtypes); mirrors include/dvsg_warp.h one to one
PROTOTYPES = {
    'dvsg_version': (c_int, []),
    'dvsg_last_error': (c_char_p, []),
    'dvsg_launch_count': (c_longlong, []),
    'dvsg_tps_solve_workspace_bytes': (c_size_t, [c_int, c_int, c_longlong]),
    'dvsg_tps_solve': (c_int, [_P, c_longlong, _P, _P, c_int, c_int, _P, c_size_t, _P]),
    'dvsg_tps_solve_bwd': (c_int, [_P, c_longlong, _P, _P, c_int, c_int, _P, c_size_t, _P]),
    'dvsg_tps_prepare_workspace_bytes': (c_size_t, [c_int, c_int, c_longlong]),
    'dvsg_tps_prepare': (c_int, [_P, c_longlong, c_int, c_int, _P, c_size_t, _P]),
    'dvsg_tps_prepare_status': (c_int, [_P, c_size_t, c_int, c_int, c_longlong, _P, ctypes.POINTER(c_int)]),
    'dvsg_tps_solve_prepared': (c_int, [_P, c_longlong, _P, _P, c_int, c_int, _P, c_size_t, _P]),
    'dvsg_tps_solve_offsets_prepared': (c_int, [_P, c_longlong, _P, _P, c_int, c_int, _P, c_size_t, _P]),
    'dvsg_tps_solve_bwd_prepared': (c_int, [_P, c_longlong, _P, _P, c_int, c_int, _P, c_size_t, _P]),
    'dvsg_tps_coord_bwd': (c_int, [_P, c_longlong, _P, _P, _P, _P, _P] + [c_int] * 4 + [_P, c_size_t, _P]),
    'dvsg_tps_warp_fwd': (c_int, [_P, _P, c_longlong, _P, _P, _P, _P, _P] + [c_int] * 8 + [_P]),
    'dvsg_tps_warp_frames': (c_int, [_P, _P, _P, _P, c_size_t, _P, _P, _P, _P, _P] + [c_int] * 7 + [_P]),
    'dvsg_tps_warp_frames_offsets': (c_int, [_P, _P, _P, _P, c_size_t, _P, _P, _P, _P, _P] + [c_int] * 7 + [_P]),
    'dvsg_tps_coords_mode': (c_int, [c_int] * 7),
    'dvsg_tps_warp_bwd_ex': (c_int, [_P, _P, c_longlong, _P, _P, _P, _P, _P, _P, _P, _P] + [c_int] * 8 + [_P]),
    'dvsg_tps_warp_bwd': (c_int, [_P, _P, c_longlong, _P, _P, _P, _P, _P, _P, _P, _P] + [c_int] * 7 + [_P]),
    'dvsg_bilinear_fwd': (c_int, [_P, _P, _P, _P] + [c_int] * 7 + [_P]),
    'dvsg_bilinear_bwd': (c_int, [_P, _P, _P, _P, _P, _P, _P] + [c_int] * 6 + [_P]),
    'dvsg_flow_warp_fwd': (c_int, [_P, _P, _P] + [c_int] * 5 + [_P]),
    'dvsg_flow_warp_bwd': (c_int, [_P, _P, _P, _P, _P] + [c_int] * 4 + [_P]),
    'dvsg_st_meshgrid': (c_int, [_P, c_int, c_int, _P]),
    'dvsg_homography_warp_fwd': (c_int, [_P, _P, c_int, _P, _P, _P] + [c_int] * 6 + [_P]),
    'dvsg_homography_grid_bwd': (c_int, [_P, _P, _P, c_int, _P] + [c_int] * 3 + [_P]),
    'dvsg_host_pipeline_create': (c_int, [ctypes.POINTER(c_void_p)] + [c_int] * 7),
    'dvsg_host_pipeline_destroy': (None, [_P]),
    'dvsg_host_tps_warp': (c_int, [_P, _P, _P, _P, _P, c_int]),
    'dvsg_host_flow_warp': (c_int, [_P, _P, _P, _P, c_int]),
    'dvsg_host_pipeline_set_async': (c_int, [_P, c_int]),
    'dvsg_host_pipeline_sync': (c_int, [_P]),
    'dvsg_frames_u8_to_f32': (c_int, [_P, _P, c_longlong, c_int, _P]),
    'dvsg_frames_f32_to_u8': (c_int, [_P, _P, c_longlong, c_int, _P]),
    'dvsg_frames_u8_resize_to_f32': (c_int, [_P, _P] + [c_int] * 6 + [_P]),
    'dvsg_flow_resize_scale': (c_int, [_P, _P] + [c_int] * 5 + [_P]),
    'dvsg_elastic_workspace_bytes': (c_size_t, [c_int]),
    'dvsg_elastic_prepare': (c_int, [_P, c_int, _P, c_size_t, _P]),
    'dvsg_elastic_solve': (c_int, [_P, _P, c_int, c_int, _P, c_size_t, _P]),
    'dvsg_elastic_solve_bwd': (c_int, [_P, _P, c_int, c_int, _P, c_size_t, _P]),
    'dvsg_elastic_grid': (c_int, [_P, _P, _P, _P] + [c_int] * 4 + [_P]),
    'dvsg_elastic_grid_bwd': (c_int, [_P, _P, _P, _P] + [c_int] * 4 + [_P]),
    'dvsg_host_tps_warp_u8': (c_int, [_P, _P, _P, _P, _P, c_int, c_int]),
    'dvsg_tps_eval_points': (c_int, [_P, c_longlong, _P, _P, _P, _P] + [c_int] * 5 + [_P]),
    'dvsg_tps_eval_points_bwd': (c_int, [_P, c_longlong, _P, _P, _P, _P] + [c_int] * 5 + [_P]),
    'dvsg_masked_mse_workspace_bytes': (c_size_t, [c_int, c_longlong]),
    'dvsg_masked_mse_fwd': (c_int, [_P, _P, _P, c_int, _P, _P, _P, _P, c_size_t, c_int, c_longlong, c_int, _P]),
    'dvsg_masked_mse_bwd': (c_int, [_P, _P, _P, c_int, _P, _P, ctypes.c_float, _P, _P, _P, c_int, c_longlong, c_int, _P]),
}
# tuning knobs used by bench / profiling experiments (exported, but not in the public header)
TUNING_PROTOTYPES = {
    'dvsg_set_bwd_tuning': (c_int, [c_int]),
    'dvsg_set_tile_tuning': (c_int, [c_int, c_int, c_int]),
    'dvsg_debug_seg_len': (c_int, [c_longlong, c_int, c_int, ctypes.c_double]),
}

_lib = None


class DvsgError(RuntimeError):
    pass


def load():
    """Load the shared library once; raises ImportError (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "coupe.dvsg_b200: %s is missing -- the CUDA library has not been built. "
            "Run `python -m coupe.dvsg_b200._build` (needs nvcc, targets sm_100a). "
            "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for table in (PROTOTYPES, TUNING_PROTOTYPES):
        for name, (res, args) in table.items():
            fn = getattr(lib, name)      # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc == OK:
        return
    msg = load().dvsg_last_error().decode('utf-8', 'replace')
    if rc == ERR_INVALID:
        raise ValueError('%s: %s' % (what, msg))
    raise DvsgError('%s failed (code %d): %s' % (what, rc, msg))


def launch_count():
    return int(load().dvsg_launch_count())
