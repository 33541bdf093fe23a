"""Host-side mirror of the reference detector interface: ``Darknet`` (yolo3/models/models.py:277-394),
``parse_model_config`` (yolo3/utils/parse_config.py:1-19) and the NMS / box utilities the video path
uses (yolo3/utils/model_build.py:12-19, 52-137, 317-332).  All arithmetic runs in libydst (CUDA, sm_100a);
this module only parses the cfg, marshals the layer list and owns the handle.
"""
import ctypes
import logging

import numpy as np
import torch

from . import _lib
from ._lib import LayerDesc, check, lib, ptr, stream_ptr


def parse_model_config(path):
    """Same contract as the reference parser (yolo3/utils/parse_config.py:1-19): a list of dicts, values kept
    as strings, '#' and blank lines dropped, conv blocks default batch_normalize=0 (int)."""
    with open(path, "r") as f:
        lines = f.read().split("\n")
    defs = []
    for raw in lines:
        if not raw or raw.startswith("#"):
            continue
        line = raw.strip()
        if not line:
            continue
        if line.startswith("["):
            defs.append({"type": line[1:-1].rstrip()})
            if defs[-1]["type"] == "convolutional":
                defs[-1]["batch_normalize"] = 0
        else:
            key, value = line.split("=")
            defs[-1][key.rstrip()] = value.strip()
    return defs


def build_layer_descs(module_defs):
    """cfg blocks (without the [net] block) -> ctypes array of ydst_layer_desc.  Negative route/shortcut
    indices are resolved to absolute layer indices here (yolo3/models/models.py:300-306 semantics)."""
    n = len(module_defs)
    arr = (LayerDesc * n)()
    for i, d in enumerate(module_defs):
        L = arr[i]
        t = d["type"]
        if t == "convolutional":
            L.type = _lib.CONV
            L.filters, L.size, L.stride = int(d["filters"]), int(d["size"]), int(d["stride"])
            L.batch_normalize = int(d["batch_normalize"])
            L.activation = _lib.ACT.get(d["activation"], 0)        # anything else is linear (models.py:53-56)
        elif t == "maxpool":
            L.type, L.size, L.stride = _lib.MAXPOOL, int(d["size"]), int(d["stride"])
        elif t == "upsample":
            L.type, L.size = _lib.UPSAMPLE, int(d["stride"])
        elif t == "route":
            L.type = _lib.ROUTE
            srcs = [int(x) for x in d["layers"].split(",")]
            if len(srcs) > 4:
                raise ValueError(f"route with {len(srcs)} sources is not supported")
            L.n_src = len(srcs)
            for k, s in enumerate(srcs):
                L.src[k] = s if s >= 0 else i + s
            if "groups" in d:
                L.groups, L.group_id = int(d["groups"]), int(d["group_id"])
        elif t == "shortcut":
            L.type, L.n_src = _lib.SHORTCUT, 1
            s = int(d["from"])
            L.src[0] = s if s >= 0 else i + s
        elif t == "yolo":
            L.type = _lib.YOLO
            mask = [int(x) for x in d["mask"].split(",")]
            a = [int(x) for x in d["anchors"].split(",")]
            if len(mask) != 3:
                raise ValueError("yolo layers with 3 anchors are supported")
            for k, mi in enumerate(mask):
                L.anchors[2 * k], L.anchors[2 * k + 1] = a[2 * mi], a[2 * mi + 1]
            L.classes = int(d["classes"])
        else:
            raise ValueError(f"unsupported cfg block [{t}]")
    return arr


class Darknet:
    """Drop-in for yolo3.models.Darknet on the inference path: ``Darknet(cfg, img_size)``,
    ``.load_darknet_weights(path)``, ``.to(device)``, ``.eval()``, ``.half()``, ``.parameters()``,
    ``.img_size`` and ``model(x) -> (B, sum 3*g*g, 5+nc)`` float32 (xywh-centre boxes in input pixels,
    sigmoid objectness and class scores), cf. yolo3/models/models.py:279-313.  The device handle is created
    lazily on the first forward (weights + device + batch are all known then)."""

    def __init__(self, config_path, img_size=416):
        logging.info("Reading config...")
        self.module_defs = parse_model_config(config_path)
        self.hyperparams = self.module_defs.pop(0)
        self.img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        self._descs = build_layer_descs(self.module_defs)
        self.header_info = np.array([0, 0, 0, 0, 0], dtype=np.int32)
        self.seen = 0
        self._weights = None
        self._device = torch.device("cpu")
        self._handles = {}
        self._token = None
        logging.info("Reading config done")

    # ---- nn.Module-like surface used by ImageDetector (yolo3/detect/img_detect.py:45-50) ----
    def to(self, device):
        self._device = torch.device(device)
        return self

    def cuda(self, device=0):
        return self.to(f"cuda:{device}")

    def eval(self):
        return self

    def half(self):
        return self            # activations are fp16 with fp32 accumulation already; nothing to convert

    def parameters(self):
        if self._token is None or self._token.device != self._device:
            self._token = torch.zeros(1, device=self._device)
        yield self._token

    # ---- weights ----
    def load_darknet_weights(self, weights_path):
        """yolo3/models/models.py:315-366: 5 x int32 header, then the float32 payload."""
        with open(weights_path, "rb") as f:
            self.header_info = np.fromfile(f, dtype=np.int32, count=5)
            self.seen = int(self.header_info[3])
            self.set_weights(np.fromfile(f, dtype=np.float32))

    def set_weights(self, flat_f32):
        self._weights = np.ascontiguousarray(flat_f32, dtype=np.float32)
        self._destroy()

    def _destroy(self):
        for h in self._handles.values():
            lib().ydst_detector_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def handle(self, batch=1):
        _lib.require_cuda()
        if self._weights is None:
            raise _lib.YdstError("Darknet has no weights: call load_darknet_weights() first")
        if self._device.type != "cuda":
            raise _lib.YdstError("Darknet must be moved to a CUDA device (model.to('cuda:0')); there is no CPU path")
        if batch not in self._handles:
            h = ctypes.c_void_p()
            with torch.cuda.device(self._device):
                check(lib().ydst_detector_create(self._descs, len(self._descs), self._weights.ctypes.data, self._weights.size,
                                                 int(self.img_size[0]), int(self.img_size[1]), int(batch), ctypes.byref(h)))
            self._handles[batch] = h
        return self._handles[batch]

    def _shape(self, h):
        rows, fields = ctypes.c_int(), ctypes.c_int()
        check(lib().ydst_detector_shape(h, ctypes.byref(rows), ctypes.byref(fields)))
        return rows.value, fields.value

    @property
    def out_shape(self):
        return self._shape(self.handle(1))

    def flops(self):
        return lib().ydst_detector_flops(self.handle(1))

    def layer_output(self, layer, batch=1):
        """Parity aid: output of cfg layer `layer` from the last forward as a dense (N,H,W,C) tensor."""
        h = self.handle(batch)
        v = [ctypes.c_int() for _ in range(5)]
        check(lib().ydst_detector_layer_shape(h, int(layer), *[ctypes.byref(x) for x in v]))
        n, hh, ww, c, f32 = (x.value for x in v)
        out = torch.empty((n, hh, ww, c), dtype=torch.float32 if f32 else torch.float16, device=self._device)
        with torch.cuda.device(self._device):
            check(lib().ydst_detector_layer_output(h, int(layer), ptr(out), stream_ptr()))
        return out

    # ---- forward ----
    def forward(self, x):
        assert x.dim() == 4 and x.shape[1] == 3 and tuple(x.shape[2:]) == tuple(self.img_size), \
            f"expected (B,3,{self.img_size[0]},{self.img_size[1]}), got {tuple(x.shape)}"
        assert x.dtype in (torch.float32, torch.float16)
        x = x.to(self._device).contiguous()
        B = x.shape[0]
        h = self.handle(B)
        rows, fields = self._shape(h)
        pred = torch.empty((B, rows, fields), dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            check(lib().ydst_detector_forward_nchw(h, ptr(x), int(x.dtype == torch.float16), ptr(pred), stream_ptr()))
        return pred

    __call__ = forward

    def forward_frame(self, frame_u8_dev):
        """(H,W,3) uint8 RGB device tensor already at the network size -> (1, rows, fields)."""
        rows, fields = self.out_shape
        pred = torch.empty((1, rows, fields), dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            check(lib().ydst_detector_forward_u8(self.handle(1), ptr(frame_u8_dev), ptr(pred), stream_ptr()))
        return pred


def soft_non_max_suppression(prediction, conf_thres=0.1, iou_thres=0.6, merge=False, classes=None, agnostic=False, is_p1p2=False):
    """yolo3/utils/model_build.py:52-137 with all of its keyword options (the video path uses none; the sliding-window mode
    passes merge=True, is_p1p2=True).  `merge` behaves exactly as the reference's block executes, see csrc/nms.cu.
    Returns a list with one (n,6) tensor [x1,y1,x2,y2,conf,cls] (or None) per image."""
    prediction = prediction.float().contiguous()
    cls = np.ascontiguousarray(classes, dtype=np.int32) if classes else None
    out = []
    for x in prediction:
        dets = torch.empty((300, 6), dtype=torch.float32, device=x.device)
        n = ctypes.c_int()
        with torch.cuda.device(x.device):
            check(lib().ydst_nms_ex(ptr(x), x.shape[0], x.shape[1], float(conf_thres), float(iou_thres), int(bool(merge)), int(bool(is_p1p2)),
                                    int(bool(agnostic)), cls.ctypes.data if cls is not None else None, 0 if cls is None else int(cls.size),
                                    ptr(dets), ctypes.byref(n), stream_ptr()))
        out.append(dets[:n.value].clone() if n.value else None)
    return out


def resize_boxes(boxes, current_dim, original_shape):
    """yolo3/utils/model_build.py:12-19 (in place)."""
    h_ratio, w_ratio = original_shape[0] / current_dim[0], original_shape[1] / current_dim[1]
    boxes[..., 0] *= w_ratio
    boxes[..., 1] *= h_ratio
    boxes[..., 2] *= w_ratio
    boxes[..., 3] *= h_ratio
    return boxes


def p1p2Toxywh(x):
    """yolo3/utils/model_build.py:326-332: (x1,y1,x2,y2) -> (x1,y1,w,h)."""
    y = x.new_empty(x.shape)
    y[..., 0], y[..., 1] = x[..., 0], x[..., 1]
    y[..., 2], y[..., 3] = x[..., 2] - x[..., 0], x[..., 3] - x[..., 1]
    return y
