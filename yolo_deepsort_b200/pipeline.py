"""Fused per-frame path: one C call per frame = ImageDetector.detect (yolo3/detect/img_detect.py:61-95) + the tracker
hand-off of VideoDetector.detect (yolo3/detect/video_detect.py:134-149) + DeepSort.update (deep_sort/deep_sort.py:46-88).
One H2D copy (the frame) and one small D2H copy (the (K,6) rows) per frame.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


class FramePipeline:
    def __init__(self, model, deepsort, thres=0.5, nms_thres=0.4, class_mask=None):
        self.model, self.deepsort = model, deepsort
        self.device = model._device
        mask = np.ascontiguousarray(class_mask if class_mask is not None else [], dtype=np.int32)
        self._mask = mask
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().ydst_pipeline_create(model.handle(1), deepsort.extractor.handle, deepsort.tracker.handle, float(thres),
                                             float(nms_thres), mask.ctypes.data if mask.size else None, int(mask.size),
                                             ctypes.byref(self._h)))
        self._out = np.zeros((deepsort.tracker.cap_tracks, 6), np.int32)
        self._dets = np.zeros((300, 6), np.float32)

    def __del__(self):
        try:
            if self._h:
                lib().ydst_pipeline_destroy(self._h)
        except Exception:
            pass

    def step(self, frame, want_dets=True):
        """frame: (H,W,3) uint8 RGB at the network size; numpy (host, ideally pinned) or a CUDA tensor.
        Returns (tracks, dets): tracks = np.int32 (K,6) or [] -- or None if nothing was detected at all (the reference
        then skips tracker.update) -- and dets = (n,6) float32 post-NMS detections (or None)."""
        k, nd = ctypes.c_int(), ctypes.c_int()
        dets_ptr = self._dets.ctypes.data if want_dets else None
        with torch.cuda.device(self.device):
            if isinstance(frame, torch.Tensor) and frame.is_cuda:
                check(lib().ydst_pipeline_step_dev(self._h, ptr(frame), self._out.ctypes.data, ctypes.byref(k), dets_ptr,
                                                   ctypes.byref(nd), stream_ptr()))
            else:
                f = frame.numpy() if isinstance(frame, torch.Tensor) else np.ascontiguousarray(frame)
                assert f.dtype == np.uint8 and f.shape == (self.model.img_size[0], self.model.img_size[1], 3)
                check(lib().ydst_pipeline_step(self._h, f.ctypes.data, self._out.ctypes.data, ctypes.byref(k), dets_ptr,
                                               ctypes.byref(nd), stream_ptr()))
        dets = self._dets[:nd.value].copy() if want_dets else None
        if k.value < 0:
            return None, dets
        return (self._out[:k.value].copy() if k.value else []), dets
