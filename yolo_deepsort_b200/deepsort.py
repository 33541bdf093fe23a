"""Host-side mirror of the reference tracker facade: ``DeepSort`` (deep_sort/deep_sort.py:16-88) with
``.update(bbox_tlwh, confidences, ori_img, payload)``, ``.clone()``, ``.tracker.tracks`` and ``.extractor``.
Feature extraction, Kalman filtering, both cost matrices and the assignment run in libydst (CUDA, sm_100a).
"""
import ctypes
from collections import namedtuple

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr
from .reid import Extractor

TrackView = namedtuple("TrackView", "track_id hits age time_since_update state mean")


class TrackerHandle:
    """Owns a ydst_tracker; exposes `.tracks` like deep_sort.sort.tracker.Tracker (read-only views)."""

    def __init__(self, max_dist, max_iou_distance, max_age, n_init, nn_budget, cap_tracks, cap_dets, device):
        self.device = torch.device(device)
        self.cap_tracks, self.cap_dets = int(cap_tracks), int(cap_dets)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().ydst_tracker_create(float(max_dist), float(max_iou_distance), int(max_age), int(n_init), int(nn_budget),
                                            self.cap_tracks, self.cap_dets, ctypes.byref(self._h)))
        self._out = np.zeros((self.cap_tracks, 6), np.int32)

    def __del__(self):
        try:
            if self._h:
                lib().ydst_tracker_destroy(self._h)
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def update(self, tlwh_dev, feat_dev, payload):
        m = int(tlwh_dev.shape[0])
        pay = np.ascontiguousarray(payload, dtype=np.int32) if m else np.zeros(1, np.int32)
        k = ctypes.c_int()
        with torch.cuda.device(self.device):
            check(lib().ydst_tracker_update(self._h, ptr(tlwh_dev) if m else None, ptr(feat_dev) if m else None,
                                            pay.ctypes.data, m, self._out.ctypes.data, ctypes.byref(k), stream_ptr()))
        return self._out[:k.value].copy()

    def table(self):
        """(n,5) int32 [track_id, hits, age, time_since_update, state] and (n,8) float32 means, in list order."""
        tab = np.zeros((self.cap_tracks, 5), np.int32)
        mean = np.zeros((self.cap_tracks, 8), np.float32)
        n = ctypes.c_int()
        with torch.cuda.device(self.device):
            check(lib().ydst_tracker_tracks(self._h, tab.ctypes.data, mean.ctypes.data, self.cap_tracks, ctypes.byref(n), stream_ptr()))
        return tab[:n.value], mean[:n.value]

    def last_matches(self):
        buf = np.zeros((self.cap_tracks, 2), np.int32)
        n = ctypes.c_int()
        check(lib().ydst_tracker_last_matches(self._h, buf.ctypes.data, self.cap_tracks, ctypes.byref(n)))
        return buf[:n.value]

    @property
    def tracks(self):
        tab, mean = self.table()
        return [TrackView(int(r[0]), int(r[1]), int(r[2]), int(r[3]), int(r[4]), mean[i]) for i, r in enumerate(tab)]


class DeepSort:
    def __init__(self, model_path, max_dist=0.2, min_confidence=0.3, nms_max_overlap=1.0, max_iou_distance=0.7, max_age=70,
                 n_init=3, nn_budget=100, use_cuda=False, cap_tracks=4096, cap_dets=2048, device="cuda:0", reid_batch=512):
        _lib.require_cuda()
        if nn_budget is None:
            raise ValueError("nn_budget=None (unbounded galleries) is not supported: pass an integer budget")
        if nms_max_overlap != 1:
            # the reference's numpy NMS (deep_sort/sort/preprocessing.py) is dead code on numpy >= 1.24 (np.float) and is
            # skipped with the demo parameters (deep_sort/deep_sort.py:52)
            raise NotImplementedError("nms_max_overlap != 1 is not supported")
        self.max_dist, self.min_confidence, self.nms_max_overlap = max_dist, min_confidence, nms_max_overlap
        self.max_iou_distance, self.max_age, self.n_init, self.nn_budget = max_iou_distance, max_age, n_init, nn_budget
        self.use_cuda = True
        self.device = torch.device(device)
        self._cap = (cap_tracks, cap_dets)
        if isinstance(model_path, (str, dict)):
            # crops beyond reid_batch go through the net in several forwards (the reference has no limit on detections per frame)
            self.extractor = Extractor(model_path, use_cuda=True, max_batch=min(cap_dets, reid_batch), device=device)
        else:
            self.extractor = model_path                    # injected extractor (deep_sort.py:28-31)
        self.tracker = TrackerHandle(max_dist, max_iou_distance, max_age, n_init, nn_budget, cap_tracks, cap_dets, device)

    def clone(self):
        return DeepSort(self.extractor, self.max_dist, self.min_confidence, self.nms_max_overlap, self.max_iou_distance,
                        self.max_age, self.n_init, self.nn_budget, True, self._cap[0], self._cap[1], str(self.device))

    def _features(self, tlwh_dev, ori_img):
        if isinstance(self.extractor, Extractor):
            frame = ori_img if isinstance(ori_img, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(ori_img))
            return self.extractor.extract(frame.to(self.device), tlwh_dev)
        # duck-typed extractor: hand it host crops exactly like the reference does (deep_sort.py:133-146)
        H, W = ori_img.shape[:2]
        crops = []
        for x, y, w, h in tlwh_dev.cpu():
            x1, x2 = max(int(x), 0), min(int(x + w), W - 1)
            y1, y2 = max(int(y), 0), min(int(y + h), H - 1)
            crops.append(ori_img[y1:y2, x1:x2])
        return self.extractor(crops).to(self.device).float().contiguous()

    def update(self, bbox_xywh, confidences, ori_img, payload):
        """bbox_xywh: (m,4) boxes in (x, y, w, h) = top-left + size (despite the name, video_detect.py:138);
        payload: (m,) class ids.  Returns np.int32 (K,6) [x1,y1,x2,y2,track_id,class_id] or [] (deep_sort.py:86-88)."""
        self.height, self.width = ori_img.shape[:2]
        tlwh = torch.as_tensor(bbox_xywh, dtype=torch.float32).to(self.device).contiguous().view(-1, 4)
        m = tlwh.shape[0]
        feats = self._features(tlwh, ori_img) if m else torch.zeros((0, 512), device=self.device)
        if isinstance(payload, torch.Tensor):
            payload = payload.detach().cpu().numpy()
        pay = np.asarray(payload).astype(np.int32).reshape(-1) if m else np.zeros(0, np.int32)
        out = self.tracker.update(tlwh, feats, pay)
        return out if len(out) else []
