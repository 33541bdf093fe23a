"""Host-side mirror of the reference appearance extractor: ``Extractor`` (deep_sort/deep/feature_extractor.py:12-58)
over ``Net(reid=True)`` (deep_sort/deep/model.py:48-95).  Crop, cv2-exact resize, normalisation and the 20-conv net
all run in libydst (CUDA, sm_100a).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

STAGES = ((1, 64, 64, False), (2, 64, 128, True), (3, 128, 256, True), (4, 256, 512, True))


def flatten_state_dict(sd):
    """'net_dict' -> flat float32 payload in the order libydst consumes it (csrc/net.cu, Reid::Reid):
    stem conv w, b; stem BN gamma, beta, mean, var; then per BasicBlock conv1 w, bn1 (g,b,m,v), conv2 w, bn2 (g,b,m,v),
    and for downsampling blocks downsample.0 w, downsample.1 (g,b,m,v).  The classifier head is unused (reid=True)."""
    def g(k):
        v = sd[k]
        return (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)).astype(np.float32).ravel()

    def bn(p):
        return [g(p + ".weight"), g(p + ".bias"), g(p + ".running_mean"), g(p + ".running_var")]

    parts = [g("conv.0.weight"), g("conv.0.bias")] + bn("conv.1")
    for li, _, _, down in STAGES:
        for bi in range(2):
            p = f"layer{li}.{bi}"
            parts += [g(p + ".conv1.weight")] + bn(p + ".bn1") + [g(p + ".conv2.weight")] + bn(p + ".bn2")
            if bi == 0 and down:
                parts += [g(p + ".downsample.0.weight")] + bn(p + ".downsample.1")
    return np.ascontiguousarray(np.concatenate(parts), dtype=np.float32)


class Extractor:
    """``Extractor(model_path, use_cuda=True)``; ``extractor(list_of_uint8_RGB_crops) -> (m,512) float32`` like the
    reference, plus the fused entry ``extract(frame_dev, tlwh_dev)`` the DeepSort mirror uses (no per-crop copies)."""

    def __init__(self, model_path, use_cuda=True, max_batch=512, device="cuda:0"):
        _lib.require_cuda()
        if not use_cuda:
            raise _lib.YdstError("Extractor(use_cuda=False): this build has no CPU path")
        if isinstance(model_path, dict):
            sd = model_path
        else:
            sd = torch.load(model_path, map_location="cpu")["net_dict"]        # feature_extractor.py:16
        self.device = torch.device(device)
        self.size = (64, 128)
        self.max_batch = int(max_batch)
        w = flatten_state_dict(sd)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().ydst_reid_create(w.ctypes.data, w.size, self.max_batch, ctypes.byref(self._h)))

    def __del__(self):
        try:
            if self._h:
                lib().ydst_reid_destroy(self._h)
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def extract(self, frame_dev, tlwh_dev):
        """frame_dev (H,W,3) uint8 RGB on the device, tlwh_dev (m,4) float32 on the device -> (m,512) float32."""
        m = int(tlwh_dev.shape[0])
        feats = torch.empty((m, 512), dtype=torch.float32, device=self.device)
        if m:
            H, W = int(frame_dev.shape[0]), int(frame_dev.shape[1])
            with torch.cuda.device(self.device):
                check(lib().ydst_reid_extract(self._h, ptr(frame_dev), H, W, ptr(tlwh_dev.contiguous()), m, ptr(feats), stream_ptr()))
        return feats

    def forward_batch(self, x_nhwc):
        """(m,128,64,3) float32 NHWC, already normalised -> (m,512)."""
        m = int(x_nhwc.shape[0])
        feats = torch.empty((m, 512), dtype=torch.float32, device=self.device)
        if m:
            with torch.cuda.device(self.device):
                check(lib().ydst_reid_forward(self._h, ptr(x_nhwc.contiguous()), m, ptr(feats), stream_ptr()))
        return feats

    def __call__(self, im_crops):
        """Reference-compatible entry (feature_extractor.py:53-58): a list of (h,w,3) uint8 RGB crops.  The crops are
        packed side by side into one atlas image so that a single crop+resize kernel reproduces cv2.resize on each."""
        if len(im_crops) == 0:
            return torch.zeros((0, 512), dtype=torch.float32, device=self.device)
        hmax = max(c.shape[0] for c in im_crops) + 1
        wsum = sum(c.shape[1] for c in im_crops) + 1
        atlas = np.zeros((hmax, wsum, 3), np.uint8)
        boxes, x = [], 0
        for c in im_crops:
            h, w = c.shape[:2]
            atlas[:h, x:x + w] = c
            boxes.append([x, 0, w, h])
            x += w
        frame = torch.from_numpy(atlas).to(self.device)
        tlwh = torch.tensor(boxes, dtype=torch.float32, device=self.device)
        return self.extract(frame, tlwh)
