"""ctypes binding of include/stereotrack_b200.h.

There is no fallback: if the library is missing or the ABI version differs
this raises, and every op raises when the library reports an error.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libstereotrack_b200.so')
ABI_VERSION = 7

ST_STATUS = dict(CAND_OVERFLOW=1, DET_OVERFLOW=2, TRACK_OVERFLOW=4, LAP_GUARD=8)


class StConfig(C.Structure):
    """st_config of include/stereotrack_b200.h (field order is the ABI)."""
    _fields_ = [
        ('max_sequences', C.c_int32), ('max_tracks', C.c_int32), ('max_dets', C.c_int32),
        ('max_candidates', C.c_int32), ('obj_score_thr', C.c_float), ('init_track_thr', C.c_float),
        ('weight_iou_with_det_scores', C.c_int32), ('match_iou_thr', C.c_double),
        ('num_tentatives', C.c_int32), ('vel_consist_weight', C.c_float), ('vel_delta_t', C.c_int32),
        ('num_frames_retain', C.c_int32), ('baseline', C.c_float), ('focal_length', C.c_float),
        ('score_thr', C.c_float), ('nms_iou_thr', C.c_float)]


class StFrameIO(C.Structure):
    """st_frame_io of include/stereotrack_b200.h."""
    _fields_ = (
        [('frame_ids', C.c_void_p), ('raw', C.c_void_p)] +
        [(n, C.c_int32) for n in ('H', 'W', 'Hr', 'Wr', 'Hp', 'Wp')] +
        [(n, C.c_void_p) for n in ('cls', 'bbox', 'obj', 'level_h', 'level_w', 'strides')] +
        [(n, C.c_int32) for n in ('num_levels', 'num_classes', 'ori_h', 'ori_w')] +
        [('scale_w', C.c_float), ('scale_h', C.c_float)] +
        [(n, C.c_void_p) for n in (
            'depth_map', 'disp3', 'mask', 'det_boxes', 'det_scores', 'det_labels', 'det_prior',
            'det_keep', 'det_counts', 'det_depth', 'det_scales', 'det_inflated', 'trk_boxes',
            'trk_depth', 'out_scale2', 'out_src', 'k3_workspace', 'out_boxes', 'out_scores', 'out_labels', 'out_scales',
            'out_depth', 'out_ids', 'out_counts', 'warp')] +
        [('depth_from_raw', C.c_int32)])


class StereoTrackError(RuntimeError):
    pass


_P = C.c_void_p
_I = C.c_int
_SIGNATURES = {
    'st_abi_version': (C.c_int, []),
    'st_error_string': (C.c_char_p, [_I]),
    'st_default_config': (None, [C.POINTER(StConfig)]),
    'st_create': (_I, [C.POINTER(StConfig), C.POINTER(_P)]),
    'st_destroy': (_I, [_P]),
    'st_reset': (_I, [_P, _P, _P]),
    'st_read_status': (_I, [_P, _P, _I]),
    'st_read_live_counts': (_I, [_P, _P, _I, _P]),
    'st_clear_status': (_I, [_P, _P]),
    'st_read_debug': (_I, [_P, _P, _I]),
    'st_debug_set_skip': (_I, [_I]),
    'st_debug_read_cost': (_I, [_P, _I, _P, _I]),
    'st_disp2depth': (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, C.c_float, C.c_float, _P]),
    'st_disp_f32_to_depth': (_I, [_P, C.c_longlong, _P, C.c_float, C.c_float, _P]),
    'st_decode_nms': (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _I, _I, C.c_float, C.c_float, _I, _I, _I,
                           _P, _P, _P, _P, _P, _P, _P]),
    'st_box_depth': (_I, [_P, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P]),
    'st_scale_boxes': (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    'st_pack_image': (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P]),
    'st_resize_pack_image': (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P]),
    'st_box_depth_mode': (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    'st_box_depth_reuse': (_I, [_P, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    'st_box_depth_raw': (_I, [_P, _I, _I, _I, _I, _I, C.c_float, C.c_float, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    'st_track_step': (_I, [_P, _I] + [_P] * 7 + [_I] + [_P] * 10),
    'st_frame_step': (_I, [_P, _I, C.POINTER(StFrameIO), _P]),
    'st_pipeline_create': (_I, [_P, _I, C.POINTER(_P)]),
    'st_pipeline_destroy': (_I, [_P]),
    'st_pipeline_submit': (_I, [_P, _I, C.POINTER(StFrameIO), _P, _P]),
    'st_pipeline_use_graphs': (_I, [_P, _I]),
    'st_pipeline_stats': (_I, [_P, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    'st_pipeline_partition': (_I, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'st_pipeline_fork': (_I, [_P, _P]),
    'st_pipeline_join': (_I, [_P, _P]),
    'st_pipeline_trace': (_I, [_P, _I]),
    'st_pipeline_read_trace': (_I, [_P, _P, _I]),
    'st_export_state': (_I, [_P, _I] + [_P] * 8),
    'st_kalman_batch': (_I, [_I, _I, _P, _P, _P, _P]),
    'st_snapshot_bytes': (C.c_longlong, [_P]),
    'st_snapshot_save': (_I, [_P, _P, C.c_longlong, _P]),
    'st_snapshot_load': (_I, [_P, _P, C.c_longlong, _P]),
    'st_kalman_project': (_I, [_I, _P, _P, _P, _P, _P]),
    'st_kalman_project_scored': (_I, [_I, _P, _P, _P, _P, _P, _P]),
    'st_kalman_batch_scored': (_I, [_I, _I, _P, _P, _P, _P, _P]),
    'st_kalman_gating': (_I, [_I, _P, _P, _I, _P, _I, _P, _P]),
    'st_lapjv_batch': (_I, [_P, _I, _I, _I, _P, C.c_double, _I, _P, _P, _P]),
}

_lib = None


def lib():
    """The loaded CUDA library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise StereoTrackError(
                f'{LIB_PATH} is missing: build it with `python -m stereotracking_b200.build` '
                '(there is no CPU or PyTorch fallback for this path)')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError = symbol missing from the .so
            fn.restype = res
            fn.argtypes = args
        if handle.st_abi_version() != ABI_VERSION:
            raise StereoTrackError('ABI version mismatch between _lib.py and the built library')
        _lib = handle
    return _lib


def exported_symbols():
    return list(_SIGNATURES)


def check(code, what=''):
    if code < 0:
        msg = lib().st_error_string(code).decode()
        raise StereoTrackError(f'{what or "stereotrack_b200"}: {msg} ({code})')
    return code
