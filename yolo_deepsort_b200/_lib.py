"""ctypes binding of libydst.so (include/ydst.h).  There is NO fallback: if the CUDA library is missing or
does not load, importing any compute entry point raises -- the product path never routes through CPU code.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libydst.so")


class YdstError(RuntimeError):
    pass


class LayerDesc(ctypes.Structure):          # ydst_layer_desc
    _fields_ = [("type", ctypes.c_int), ("filters", ctypes.c_int), ("size", ctypes.c_int), ("stride", ctypes.c_int),
                ("batch_normalize", ctypes.c_int), ("activation", ctypes.c_int), ("n_src", ctypes.c_int),
                ("src", ctypes.c_int * 4), ("groups", ctypes.c_int), ("group_id", ctypes.c_int), ("classes", ctypes.c_int),
                ("anchors", ctypes.c_float * 6)]


CONV, MAXPOOL, UPSAMPLE, ROUTE, SHORTCUT, YOLO = range(6)
ACT = {"linear": 0, "leaky": 1, "mish": 2, "relu": 3}

_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_SZ = ctypes.c_size_t

# name -> (restype, argtypes); every symbol include/ydst.h declares
SIGNATURES = {
    "ydst_last_error": (ctypes.c_char_p, []),
    "ydst_version": (_I, []),
    "ydst_launch_count": (ctypes.c_longlong, []),
    "ydst_profile_begin": (_I, []),
    "ydst_profile_end": (_I, [_I, _P, _P, _P, _P, _P, ctypes.POINTER(_I)]),
    "ydst_detector_create": (_I, [ctypes.POINTER(LayerDesc), _I, _P, _SZ, _I, _I, _I, ctypes.POINTER(_P)]),
    "ydst_detector_destroy": (_I, [_P]),
    "ydst_detector_shape": (_I, [_P, ctypes.POINTER(_I), ctypes.POINTER(_I)]),
    "ydst_detector_forward_nchw": (_I, [_P, _P, _I, _P, _P]),
    "ydst_detector_forward_u8": (_I, [_P, _P, _P, _P]),
    "ydst_detector_nms": (_I, [_P, _F, _F, _P, _P, _P]),
    "ydst_detector_layer_shape": (_I, [_P, _I] + [ctypes.POINTER(_I)] * 5),
    "ydst_detector_layer_output": (_I, [_P, _I, _P, _P]),
    "ydst_detector_flops": (ctypes.c_double, [_P]),
    "ydst_detector_launches": (_I, [_P]),
    "ydst_nms": (_I, [_P, _I, _I, _F, _F, _P, ctypes.POINTER(_I), _P]),
    "ydst_nms_ex": (_I, [_P, _I, _I, _F, _F, _I, _I, _I, _P, _I, _P, ctypes.POINTER(_I), _P]),
    "ydst_window_boxes": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "ydst_conv2d": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P, _I, _P, _I, _P, _I, _P]),
    "ydst_conv_tiling": (_I, [_I, _I, _I, _I, _I, _I] + [ctypes.POINTER(_I)] * 4 + [ctypes.POINTER(ctypes.c_double)]),
    "ydst_reid_create": (_I, [_P, _SZ, _I, ctypes.POINTER(_P)]),
    "ydst_reid_destroy": (_I, [_P]),
    "ydst_reid_extract": (_I, [_P, _P, _I, _I, _P, _I, _P, _P]),
    "ydst_reid_forward": (_I, [_P, _P, _I, _P, _P]),
    "ydst_crop_resize": (_I, [_P, _I, _I, _P, _I, _P, _P]),
    "ydst_reid_flops_per_crop": (ctypes.c_double, []),
    "ydst_kf_initiate": (_I, [_P, _I, _P, _P, _P]),
    "ydst_kf_predict": (_I, [_P, _P, _I, _P]),
    "ydst_kf_update": (_I, [_P, _P, _P, _I, _P]),
    "ydst_gate_position": (_I, [_P, _P, _I, _P, _I, _P, _P]),
    "ydst_appearance_cost": (_I, [_P, _P, _I, _P, _I, _P, _P, _P, ctypes.c_double, _P, _P]),
    "ydst_iou_cost": (_I, [_P, _P, _I, _P, _I, ctypes.c_double, _P, _P]),
    "ydst_lsap": (_I, [_P, _I, _I, _F, _P, _P, _P, _P]),
    "ydst_action_create": (_I, [_I, _I, _P, _P, _P, _P, _I, _I, ctypes.POINTER(_P)]),
    "ydst_action_destroy": (_I, [_P]),
    "ydst_action_update": (_I, [_P, _P, _I, ctypes.c_double, _P, ctypes.POINTER(_I), _P]),
    "ydst_tracker_create": (_I, [ctypes.c_double, ctypes.c_double, _I, _I, _I, _I, _I, ctypes.POINTER(_P)]),
    "ydst_tracker_destroy": (_I, [_P]),
    "ydst_tracker_update": (_I, [_P, _P, _P, _P, _I, _P, ctypes.POINTER(_I), _P]),
    "ydst_tracker_update_dev": (_I, [_P, _P, _P, _P, _I, _P, ctypes.POINTER(_I), _P]),
    "ydst_tracker_tracks": (_I, [_P, _P, _P, _I, ctypes.POINTER(_I), _P]),
    "ydst_tracker_last_matches": (_I, [_P, _P, _I, ctypes.POINTER(_I)]),
    "ydst_pipeline_create": (_I, [_P, _P, _P, _F, _F, _P, _I, ctypes.POINTER(_P)]),
    "ydst_pipeline_destroy": (_I, [_P]),
    "ydst_pipeline_submit": (_I, [_P, _P, _I, _I, _P]),
    "ydst_pipeline_submit_frame": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "ydst_resize_u8": (_I, [_P, _I, _I, _P, _I, _I, _I, _P]),
    "ydst_resize_u8_roi": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _P]),
    "ydst_pipeline_collect": (_I, [_P, _P, ctypes.POINTER(_I), _P, ctypes.POINTER(_I)]),
    "ydst_pipeline_last_inputs": (_I, [_P, _P, _P, _P, _I, ctypes.POINTER(_I)]),
    "ydst_pipeline_in_flight": (_I, [_P]),
    "ydst_pipeline_can_submit": (_I, [_P]),
    "ydst_pipeline_step": (_I, [_P, _P, _P, ctypes.POINTER(_I), _P, ctypes.POINTER(_I), _P]),
    "ydst_pipeline_step_dev": (_I, [_P, _P, _P, ctypes.POINTER(_I), _P, ctypes.POINTER(_I), _P]),
}

_lib = None


def lib():
    """Load libydst.so (once).  Raises YdstError if it is absent -- build it with
    `python -m yolo_deepsort_b200.build` (nvcc, sm_100a)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise YdstError(f"{LIB_PATH} is missing: the CUDA library has not been built "
                            "(python -m yolo_deepsort_b200.build); there is no CPU fallback")
        try:
            L = ctypes.CDLL(LIB_PATH)
        except OSError as e:
            raise YdstError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise YdstError(lib().ydst_last_error().decode("utf-8", "replace") + f" (status {status})")


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        assert t.flags["C_CONTIGUOUS"]
        return t.ctypes.data
    assert t.is_contiguous()
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise YdstError("no CUDA device: yolo_deepsort_b200 runs on B200 (sm_100a) only and has no CPU fallback")
