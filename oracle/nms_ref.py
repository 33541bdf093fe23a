"""Oracle: torchvision.ops.nms restated in numpy fp32.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Third-party algorithm: torchvision 0.26.0 ``torchvision.ops.boxes.nms`` (CPU kernel
``nms_kernel_impl``), called by the reference at yolo3/utils/model_build.py:119.  It is not
under /root/reference, so it is restated here from its published behaviour
and pinned by live differential fuzzing against the installed torchvision in
tests/test_oracle_thirdparty.py:

  * order = stable argsort of scores, descending;
  * areas = (x2-x1)*(y2-y1) in fp32, no +1;
  * walk the order; a box that is still alive is kept and suppresses every later alive box j
    with  inter / (area_i + area_j - inter) > iou_threshold,  inter = max(0,xx2-xx1)*max(0,yy2-yy1);
  * returns kept indices in score-descending order.
"""
import numpy as np


def nms_ref(boxes, scores, iou_threshold):
    boxes = np.asarray(boxes, np.float32)
    scores = np.asarray(scores, np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    order = np.argsort(-scores, kind="stable")
    dead = np.zeros(n, bool)
    thr = np.float32(iou_threshold)
    keep = []
    for pos in range(n):
        i = order[pos]
        if dead[i]:
            continue
        keep.append(i)
        rest = order[pos + 1:]
        xx1 = np.maximum(x1[i], x1[rest]); yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest]); yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        dead[rest[ovr > thr]] = True
    return np.asarray(keep, np.int64)
