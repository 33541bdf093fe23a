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
    def __init__(self, model, deepsort, thres=0.5, nms_thres=0.4, class_mask=None, micro_batch=1):
        """micro_batch > 1: that many consecutive frames share one Darknet forward and one ReID forward (run() / submit() /
        collect() then keep up to 2*micro_batch frames in flight); per-frame results are unchanged."""
        self.model, self.deepsort = model, deepsort
        self.micro_batch = int(micro_batch)
        self.device = model._device
        mask = np.ascontiguousarray(class_mask if class_mask is not None else [], dtype=np.int32)
        self._mask = mask
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().ydst_pipeline_create(model.handle(self.micro_batch), deepsort.extractor.handle, deepsort.tracker.handle, float(thres),
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

    # ---- software-pipelined form: the detector half of frame t+1 runs under the ReID + association half of frame t ----
    def submit(self, frame, want_dets=True, bgr=False):
        """Enqueue the detector half of `frame` and return at once (up to 3 micro-batches may be in flight).  `frame` is
        (h,w,3) uint8, RGB (or BGR with bgr=True), of ANY size: colour swap and the cv2-exact resize to the network size run
        on the device, detections come back in the frame's own pixels and the ReID crops are cut from the frame itself."""
        with torch.cuda.device(self.device):
            if isinstance(frame, torch.Tensor) and frame.is_cuda:
                assert frame.dtype == torch.uint8 and frame.dim() == 3 and frame.shape[2] == 3 and frame.is_contiguous()
                check(lib().ydst_pipeline_submit_frame(self._h, ptr(frame), int(frame.shape[0]), int(frame.shape[1]), 0, int(bgr),
                                                       int(want_dets), stream_ptr()))
            else:
                f = frame.numpy() if isinstance(frame, torch.Tensor) else np.ascontiguousarray(frame)
                assert f.dtype == np.uint8 and f.ndim == 3 and f.shape[2] == 3
                if f.shape[:2] != tuple(self.model.img_size) or bgr:
                    self._keep = (getattr(self, '_keep', ()) + (f,))[-4 * self.micro_batch - 2:]
                    check(lib().ydst_pipeline_submit_frame(self._h, f.ctypes.data, int(f.shape[0]), int(f.shape[1]), 1, int(bgr),
                                                           int(want_dets), stream_ptr()))
                    return
                self._keep = (getattr(self, '_keep', ()) + (f,))[-4 * self.micro_batch - 2:]   # async H2D copies read them until collected
                check(lib().ydst_pipeline_submit(self._h, f.ctypes.data, 1, int(want_dets), stream_ptr()))

    def collect(self, want_dets=True):
        """Finish the oldest submitted frame; same return value as step()."""
        k, nd = ctypes.c_int(), ctypes.c_int()
        dets_ptr = self._dets.ctypes.data if want_dets else None
        with torch.cuda.device(self.device):
            check(lib().ydst_pipeline_collect(self._h, self._out.ctypes.data, ctypes.byref(k), dets_ptr, ctypes.byref(nd)))
        dets = self._dets[:nd.value].copy() if want_dets else None
        if k.value < 0:
            return None, dets
        return (self._out[:k.value].copy() if k.value else []), dets

    def last_inputs(self):
        """(tlwh (m,4) f32, features (m,512) f32, class ids (m,) i32): what the tracker was handed for the frame returned by the
        last collect()/step() (deep_sort/deep_sort.py:55-60).  Parity aid."""
        cap = self._dets.shape[0]
        tl, ft, cl = np.zeros((cap, 4), np.float32), np.zeros((cap, 512), np.float32), np.zeros(cap, np.int32)
        m = ctypes.c_int()
        with torch.cuda.device(self.device):
            check(lib().ydst_pipeline_last_inputs(self._h, tl.ctypes.data, ft.ctypes.data, cl.ctypes.data, cap, ctypes.byref(m)))
        return tl[:m.value].copy(), ft[:m.value].copy(), cl[:m.value].copy()

    def in_flight(self):
        return int(lib().ydst_pipeline_in_flight(self._h))

    def drain(self):
        while self.in_flight():
            self.collect(want_dets=False)

    def can_submit(self):
        return bool(lib().ydst_pipeline_can_submit(self._h))

    def run(self, frames, want_dets=True):
        """Generator over an iterable of frames with look-ahead (one frame, or a micro-batch): yields (tracks, dets) per frame,
        in order."""
        it = iter(frames)
        done = False
        while True:
            while not done and self.can_submit():
                try:
                    self.submit(next(it), want_dets)
                except StopIteration:
                    done = True
            if not self.in_flight():
                return
            yield self.collect(want_dets)

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
