"""Oracle: Darknet cfg interpreter, YOLO decode, NMS hand-off (CPU, torch fp32).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  A functional restatement of
the reference detector: no nn.Module graph, weights live in a flat list of
per-conv dicts, every layer is evaluated with the same ATen fp32 op the
reference module would dispatch to, so on CPU the result is bit-identical to
the reference (verified by oracle/gen_golden.py).

Follows (reference file:line):
  parse_cfg            yolo3/utils/parse_config.py:1-19
  layer semantics      yolo3/models/models.py:25-102  (create_modules)
  forward              yolo3/models/models.py:292-313 (Darknet.forward)
  yolo_decode          yolo3/models/models.py:167-224 (YOLOLayer, inference branch)
  read/write_weights   yolo3/models/models.py:315-394
  postprocess          yolo3/utils/model_build.py:52-137 (soft_non_max_suppression),
                       :317-323 (xywh2p1p2), :12-19 (resize_boxes), :326-332 (p1p2Toxywh),
                       yolo3/detect/img_detect.py:61-95, yolo3/detect/video_detect.py:138-147
"""
import numpy as np
import torch
import torch.nn.functional as F

from .nms_ref import nms_ref


def parse_cfg(path_or_text):
    """List of block dicts; values stay strings, conv blocks default batch_normalize=0
    (yolo3/utils/parse_config.py:5-17)."""
    text = path_or_text
    if "\n" not in path_or_text and "[" not in path_or_text:
        with open(path_or_text, "r") as f:
            text = f.read()
    blocks = []
    for raw in text.split("\n"):
        if not raw or raw.startswith("#"):
            continue
        line = raw.strip()
        if not line:
            continue
        if line.startswith("["):
            blocks.append({"type": line[1:-1].rstrip()})
            if blocks[-1]["type"] == "convolutional":
                blocks[-1]["batch_normalize"] = 0
        else:
            key, value = line.split("=")
            blocks[-1][key.rstrip()] = value.strip()
    return blocks


def layer_channels(blocks):
    """Output channel count per layer (create_modules' output_filters bookkeeping,
    yolo3/models/models.py:30,68-75,83-84,100)."""
    net = blocks[0]
    out = [int(net["channels"])]
    for b in blocks[1:]:
        t = b["type"]
        filters = out[-1]
        if t == "convolutional":
            filters = int(b["filters"])
        elif t == "route":
            idx = [int(x) for x in b["layers"].split(",")]
            filters = sum(out[1:][i] for i in idx)
            if "groups" in b:
                filters //= int(b["groups"])
        elif t == "shortcut":
            filters = out[1:][int(b["from"])]
        out.append(filters)
    return out


BETA_MEAN = 1.5


def init_weights(blocks, seed=0, beta_mean=BETA_MEAN):
    """Seeded synthetic weights: one dict per conv block, in cfg order.
    {'w': (Cout,Cin,k,k), and either 'bn': [gamma,beta,mean,var] or 'b': bias}.
    BN beta ~ N(beta_mean, 0.1): a randomly initialised BN + leaky-ReLU stack with beta = 0 sits in the CHAOTIC phase (mean-field
    perturbation gain E[phi'^2] / Var[phi] = 1.34 per layer: any rounding noise doubles every ~2.4 layers, which no trained
    network does), so its outputs say nothing about an implementation's arithmetic.  beta_mean = 1.5 moves the stack to the
    ordered edge (gain ~1.02 per layer) without changing a single FLOP ("Precision vs parity")."""
    g = torch.Generator().manual_seed(seed)
    chans = layer_channels(blocks)
    ws = []
    for li, b in enumerate(blocks[1:]):
        if b["type"] != "convolutional":
            continue
        cin, cout, k = chans[li], int(b["filters"]), int(b["size"])
        fan_in = cin * k * k
        w = torch.randn(cout, cin, k, k, generator=g) * float(np.sqrt(2.0 / fan_in))
        d = {"w": w.numpy().copy()}
        if int(b["batch_normalize"]):
            gamma = 1.0 + 0.1 * torch.randn(cout, generator=g)
            beta = float(beta_mean) + 0.1 * torch.randn(cout, generator=g)
            mean = 0.1 * torch.randn(cout, generator=g)
            var = 0.5 + torch.rand(cout, generator=g)
            d["bn"] = [t.numpy().copy() for t in (gamma, beta, mean, var)]
        else:
            d["b"] = (0.1 * torch.randn(cout, generator=g)).numpy().copy()
        ws.append(d)
    return ws


def write_weights(path, blocks, ws, header=(0, 2, 0, 0, 0)):
    """Darknet .weights writer (layout of yolo3/models/models.py:368-394): 5 x int32 header,
    then per conv [bn.bias, bn.weight, running_mean, running_var | conv.bias], conv.weight."""
    with open(path, "wb") as f:
        np.asarray(header, dtype=np.int32).tofile(f)
        it = iter(ws)
        for b in blocks[1:]:
            if b["type"] != "convolutional":
                continue
            d = next(it)
            if "bn" in d:
                gamma, beta, mean, var = d["bn"]
                for a in (beta, gamma, mean, var):
                    np.asarray(a, np.float32).tofile(f)
            else:
                np.asarray(d["b"], np.float32).tofile(f)
            np.asarray(d["w"], np.float32).tofile(f)


def read_weights(path, blocks):
    """Darknet .weights reader (yolo3/models/models.py:315-366).  `if batch_normalize`
    is a truthiness test on the parsed value: the string '1'/'0' is truthy, int 0 is not (:336)."""
    with open(path, "rb") as f:
        header = np.fromfile(f, dtype=np.int32, count=5)
        flat = np.fromfile(f, dtype=np.float32)
    chans = layer_channels(blocks)
    ws, p = [], 0
    for li, b in enumerate(blocks[1:]):
        if b["type"] != "convolutional":
            continue
        cin, cout, k = chans[li], int(b["filters"]), int(b["size"])
        d = {}
        if b["batch_normalize"]:
            beta, gamma, mean, var = (flat[p + i * cout:p + (i + 1) * cout].copy() for i in range(4))
            p += 4 * cout
            d["bn"] = [gamma, beta, mean, var]
        else:
            d["b"] = flat[p:p + cout].copy()
            p += cout
        n = cout * cin * k * k
        d["w"] = flat[p:p + n].reshape(cout, cin, k, k).copy()
        p += n
        ws.append(d)
    return header, ws


def _mish(x):
    return x * torch.tanh(F.softplus(x))


def yolo_decode(x, anchors, num_classes, img_dim):
    """YOLOLayer.forward inference branch (yolo3/models/models.py:185-224), including the
    quirk that x is scaled by the *height* stride and scale_x_y is ignored."""
    B, _, gy, gx = x.shape
    na = len(anchors)
    pred = x.view(B, na, num_classes + 5, gy, gx).permute(0, 1, 3, 4, 2)
    xy = torch.sigmoid(pred[..., 0:2])
    wh = pred[..., 2:4]
    conf_cls = torch.sigmoid(pred[..., 4:])
    scale = torch.as_tensor([[img_dim[0] / gy, img_dim[1] / gx]], dtype=x.dtype)
    yy, xx = torch.meshgrid([torch.arange(gy, dtype=torch.int32), torch.arange(gx, dtype=torch.int32)],
                            indexing="ij")
    grid = torch.stack((xx.type(x.dtype).flatten(), yy.type(x.dtype).flatten()), 1).view(1, 1, gy, gx, 2)
    anchor = (torch.tensor(anchors, dtype=x.dtype) / scale).view(1, na, 1, 1, 2)
    boxes = torch.cat([xy + grid, torch.exp(wh) * anchor], dim=-1)
    return torch.cat((boxes.reshape(B, -1, 4) * scale.repeat(1, 2),
                      conf_cls[..., 0].reshape(B, -1, 1),
                      conf_cls[..., 1:].reshape(B, -1, num_classes)), -1)


def forward(blocks, ws, x, return_layers=False, calibrate_bn=False, half_storage=False, teacher=None, fused=()):
    """Darknet.forward (yolo3/models/models.py:292-313).  x: (B,3,H,W) float32 in [0,1].
    calibrate_bn=True is a synthetic-weights helper (not reference behaviour): it overwrites each BN's
    running mean/var in `ws` with the statistics of its input on `x`, so seeded random weights keep
    unit-scale activations through 75+ layers.
    half_storage=True restates the reference's own half=True mode (yolo3/detect/img_detect.py:48-50,79-82: model.half(),
    fp16 activations and weights) on the CPU: weights and every materialised activation are rounded to fp16, the arithmetic
    stays fp32 (what fp16 tensor-core convolutions do).  The first conv keeps fp32 weights and input and the head convs keep
    fp32 outputs, as the CUDA path does.
    teacher: optional list (one entry per layer, None allowed) of layer outputs produced by ANOTHER implementation; every
    layer then reads its inputs from `teacher` instead of from this function's own outputs, so each layer is checked in
    isolation on identical inputs and rounding noise cannot compound through the depth of the net (test aid).
    fused: cfg indices of convolutions whose following shortcut is applied inside the conv (the other implementation's
    buffer for that layer holds the post-add tensor); the shortcut layer then passes its input through."""
    x = torch.as_tensor(x)
    img_dim = (x.shape[2], x.shape[3])
    outs, yolo = [], []
    it = iter(ws)
    h16 = (lambda t_: t_.half().float()) if half_storage else (lambda t_: t_)
    body = blocks[1:]

    def src(i, li):
        i = i if i >= 0 else li + i
        return teacher[i] if teacher is not None and teacher[i] is not None else outs[i]

    with torch.no_grad():
        for li, b in enumerate(body):
            t = b["type"]
            if li > 0 and t in ("convolutional", "maxpool", "upsample"):
                x = src(li - 1, li)
            if t == "convolutional":
                d = next(it)
                k = int(b["size"])
                bias = None if "bn" in d else torch.from_numpy(d["b"])
                w = torch.from_numpy(d["w"])
                x = F.conv2d(x, w if li == 0 else h16(w), bias, stride=int(b["stride"]), padding=(k - 1) // 2)
                if "bn" in d:
                    if calibrate_bn:
                        d["bn"][2] = x.mean(dim=(0, 2, 3)).numpy().copy()
                        d["bn"][3] = x.var(dim=(0, 2, 3), unbiased=False).numpy().copy() + np.float32(1e-3)
                    gamma, beta, mean, var = (torch.from_numpy(a) for a in d["bn"])
                    x = F.batch_norm(x, mean, var, gamma, beta, False, 0.1, 1e-5)
                if b["activation"] == "leaky":
                    x = F.leaky_relu(x, 0.1)
                elif b["activation"] == "mish":
                    x = _mish(x)
                nxt = body[li + 1] if li + 1 < len(body) else {"type": ""}
                if li in fused:
                    x = h16(x + src(li + 1 + int(nxt["from"]), 0))
                elif nxt["type"] not in ("yolo", "shortcut"):
                    x = h16(x)
            elif t == "maxpool":
                k, s = int(b["size"]), int(b["stride"])
                if k == 2 and s == 1:
                    x = F.pad(x, (0, 1, 0, 1))          # ZeroPad2d: zeros, not -inf (models.py:61-62)
                x = F.max_pool2d(x, k, s, (k - 1) // 2)
            elif t == "upsample":
                s = int(b["stride"])
                x = x.repeat_interleave(s, 2).repeat_interleave(s, 3)
            elif t == "route":
                x = torch.cat([src(int(i), li) for i in b["layers"].split(",")], 1)
                if "groups" in b:
                    x = x.chunk(int(b["groups"]), dim=1)[int(b["group_id"])]
            elif t == "shortcut":
                x = src(li - 1, li) if (li - 1) in fused else h16(src(li - 1, li) + src(int(b["from"]), li))
            elif t == "yolo":
                mask = [int(v) for v in b["mask"].split(",")]
                a = [int(v) for v in b["anchors"].split(",")]
                anchors = [(a[2 * i], a[2 * i + 1]) for i in mask]
                x = yolo_decode(src(li - 1, li), anchors, int(b["classes"]), img_dim)
                yolo.append(x)
            outs.append(x)
    y = torch.cat(yolo, 1)
    return (y, outs) if return_layers else y


def _bbox_iou_elementwise(box1, box2):
    """bbox_iou, p1p2=True (yolo3/utils/model_build.py:354-381): ELEMENTWISE (broadcasting) IoU with "+1" extents."""
    inter_mins = torch.max(box1[..., :2], box2[..., :2])
    inter_maxes = torch.min(box1[..., 2:4], box2[..., 2:4])
    inter_wh = torch.clamp(inter_maxes - inter_mins + 1, min=0)
    inter_area = inter_wh[..., 0] * inter_wh[..., 1]
    a1 = (box1[..., 2] - box1[..., 0] + 1) * (box1[..., 3] - box1[..., 1] + 1)
    a2 = (box2[..., 2] - box2[..., 0] + 1) * (box2[..., 3] - box2[..., 1] + 1)
    return inter_area / (a1 + a2 - inter_area + 1e-16)


def postprocess(pred, conf_thres, iou_thres, max_det=300, merge=False, is_p1p2=False, classes=None, agnostic=False):
    """soft_non_max_suppression for one image, multi_label=True (yolo3/utils/model_build.py:52-137).
    pred: (R, 5+nc) float32, xywh-centre boxes (corner boxes with is_p1p2).
    Returns (n,6) float32 [x1,y1,x2,y2,conf,cls] in score-descending order, or None.

    merge=True restates the "Merge NMS" block (:122-131) AS IT EXECUTES: the reference calls its elementwise bbox_iou where
    ultralytics has the pairwise box_iou, inside a bare try/except.  With k kept rows of n candidates the shapes only broadcast
    for k == n or k == 1; otherwise the line raises, the exception is swallowed and nothing is merged.  When they do broadcast,
    `weights` is a (1,n) row, torch.mm yields ONE weighted mean box that the broadcast assignment writes into every kept row,
    and the following `iou.sum(1)` raises on the 1-D tensor (swallowed too), so the `redundant` filter never runs."""
    x = np.array(pred, dtype=np.float32, copy=True)
    x = x[x[:, 4] > np.float32(conf_thres)]
    if not x.shape[0]:
        return None
    x[:, 5:] *= x[:, 4:5]
    if is_p1p2:
        box = x[:, :4].copy()
    else:
        box = np.empty((x.shape[0], 4), np.float32)
        box[:, 0] = x[:, 0] - x[:, 2] / np.float32(2.)
        box[:, 1] = x[:, 1] - x[:, 3] / np.float32(2.)
        box[:, 2] = x[:, 0] + x[:, 2] / np.float32(2.)
        box[:, 3] = x[:, 1] + x[:, 3] / np.float32(2.)
    i, j = np.nonzero(x[:, 5:] > np.float32(conf_thres))          # row-major order
    det = np.concatenate((box[i], x[i, j + 5][:, None], j[:, None].astype(np.float32)), 1)
    if classes:
        det = det[np.isin(det[:, 5], np.asarray(classes, np.float32))]
    n = det.shape[0]
    if not n:
        return None
    c = det[:, 5:6] * np.float32(0 if agnostic else 4096)
    boxes = det[:, :4] + c
    keep = nms_ref(boxes, det[:, 4], iou_thres)[:max_det]
    if merge and 1 < n < 3000 and len(keep) in (1, n):
        tb = torch.from_numpy(boxes)
        iou = _bbox_iou_elementwise(tb[torch.from_numpy(keep)], tb) > iou_thres            # (n,)
        weights = iou * torch.from_numpy(det[:, 4])[None]                                    # (1,n)
        merged = torch.mm(weights, torch.from_numpy(det[:, :4])).float() / weights.sum(1, keepdim=True)
        det[keep, :4] = merged.numpy()
    return det[keep]


def detect_windows(blocks, ws, img_rgb_u8, img_size, win_size, overlap, conf_thres, iou_thres):
    """ImageDetector.detect, sliding-window branch, half=False (yolo3/detect/img_detect.py:97-151): windows of win_size
    (width, height) plus `overlap` of it on the right / bottom, x outer and y inner, each cv2-resized to the network size and
    pushed through the net as ONE batch; boxes to corners, scaled back to the window, shifted by the window origin, all windows
    concatenated, then soft_non_max_suppression(merge=True, is_p1p2=True)."""
    import cv2
    h, w, _ = img_rgb_u8.shape
    win_w, win_h = win_size
    ov_x, ov_y = int(win_w * overlap), int(win_h * overlap)
    tiles, sizes, offsets = [], [], []
    for x in range(0, w, win_w):
        for y in range(0, h, win_h):
            sub = img_rgb_u8[y:y + win_h + ov_y, x:x + win_w + ov_x]
            sizes.append((sub.shape[0], sub.shape[1]))
            tiles.append(cv2.resize(sub, (img_size[1], img_size[0]), interpolation=cv2.INTER_LINEAR))
            offsets.append(torch.tensor([x, y, x, y], dtype=torch.float32))
    xb = torch.from_numpy(np.stack(tiles, 0)).permute(0, 3, 1, 2) / 255.
    det = forward(blocks, ws, xb)
    b = det[..., :4].clone()
    det[..., 0] = b[..., 0] - b[..., 2] / 2
    det[..., 1] = b[..., 1] - b[..., 3] / 2
    det[..., 2] = b[..., 0] + b[..., 2] / 2
    det[..., 3] = b[..., 1] + b[..., 3] / 2
    out = []
    for t in range(det.shape[0]):
        d = det[t]
        h_ratio, w_ratio = sizes[t][0] / img_size[0], sizes[t][1] / img_size[1]
        d[..., 0] *= w_ratio
        d[..., 1] *= h_ratio
        d[..., 2] *= w_ratio
        d[..., 3] *= h_ratio
        d[..., :4] += offsets[t]
        out.append(d)
    allp = torch.cat(out, 0)
    return postprocess(allp.numpy(), conf_thres, iou_thres, merge=True, is_p1p2=True), tiles


def detect(blocks, ws, frame_rgb_u8, img_size, conf_thres, iou_thres):
    """ImageDetector.detect, win_size=None, half=False (yolo3/detect/img_detect.py:61-95).
    The frame must already have the model size (cv2.resize is then a copy)."""
    h, w, _ = frame_rgb_u8.shape
    assert (h, w) == tuple(img_size), "oracle detect(): frame must be pre-sized to the model"
    x = torch.from_numpy(np.ascontiguousarray(frame_rgb_u8)).permute(2, 0, 1) / 255.
    pred = forward(blocks, ws, x.unsqueeze(0))
    det = postprocess(pred[0].numpy(), conf_thres, iou_thres)
    if det is not None:                                   # resize_boxes (model_build.py:12-19)
        det[:, 0] *= np.float32(w / img_size[1]); det[:, 2] *= np.float32(w / img_size[1])
        det[:, 1] *= np.float32(h / img_size[0]); det[:, 3] *= np.float32(h / img_size[0])
    return det


def to_tracker_inputs(det, class_mask=None):
    """p1p2Toxywh + class mask (yolo3/detect/video_detect.py:138-147): returns
    (tlwh (m,4) f32, conf (m,), class_ids (m,) f32)."""
    tlwh = det[:, :4].copy()
    tlwh[:, 2] = det[:, 2] - det[:, 0]
    tlwh[:, 3] = det[:, 3] - det[:, 1]
    cls, conf = det[:, 5], det[:, 4]
    if class_mask is not None:
        m = np.zeros(len(cls), bool)
        for c in class_mask:
            m |= cls == c
        tlwh, conf, cls = tlwh[m], conf[m], cls[m]
    return tlwh, conf, cls
