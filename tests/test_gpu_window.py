"""ImageDetector's sliding-window mode (yolo3/detect/img_detect.py:97-151) and the merge / is_p1p2 / classes /
agnostic options of soft_non_max_suppression (yolo3/utils/model_build.py:52-137) on the device, against the golden written by the
unmodified reference (tests/golden/window.npz) and against the oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from util import DEV
from yolo_deepsort_b200 import Darknet, ImageDetector, soft_non_max_suppression

pytestmark = pytest.mark.gpu


def test_merge_nms_options_match_reference():
    from oracle import darknet_ref as D
    from oracle.gen_golden import merge_cases
    g = np.load(os.path.join(GOLDEN, "window.npz"))
    for name, p in merge_cases().items():
        got = soft_non_max_suppression(torch.from_numpy(p)[None].to(DEV), 0.5, 0.4, merge=True, is_p1p2=True)[0].cpu().numpy()
        ref = g["merge_" + name]
        assert got.shape == ref.shape, name
        np.testing.assert_array_equal(got[:, 4:], ref[:, 4:], err_msg=name)                # scores, classes: exact
        np.testing.assert_allclose(got[:, :4], ref[:, :4], rtol=1e-6, atol=1e-4, err_msg=name)   # weighted mean: fp32 sum order
    # is_p1p2 without merge, classes, agnostic: bit-exact against the oracle on seeded corner-box predictions
    rng = np.random.default_rng(5)
    R = 600
    pred = np.zeros((R, 85), np.float32)
    xy = rng.uniform(20, 500, (R, 2)); wh = rng.uniform(20, 120, (R, 2))
    pred[:, 0:2] = xy; pred[:, 2:4] = xy + wh
    pred[:, 4] = rng.uniform(0.3, 1.0, R)
    pred[:, 5:] = rng.uniform(0, 0.4, (R, 80))
    for i in range(R):
        pred[i, 5 + rng.integers(0, 6)] = rng.uniform(0.6, 1.0)
    for kw in (dict(is_p1p2=True), dict(is_p1p2=True, classes=[0, 2, 4]), dict(is_p1p2=True, agnostic=True), dict(is_p1p2=True, merge=True)):
        ref = D.postprocess(pred, 0.5, 0.4, **kw)
        got = soft_non_max_suppression(torch.from_numpy(pred)[None].to(DEV), 0.5, 0.4, **kw)[0]
        np.testing.assert_array_equal(got.cpu().numpy(), ref, err_msg=str(kw))


def test_sliding_window_detect_matches_reference(tmp_path):
    from oracle.gen_golden import window_image_and_weights
    from oracle.cv_resize_ref import resize_linear_u8
    from test_gpu_detector import flatten_ws
    g = np.load(os.path.join(GOLDEN, "window.npz"))
    cfg, blocks, ws, img, win, ov, info = window_image_and_weights()
    model = Darknet(cfg, img_size=(416, 416))
    model.set_weights(flatten_ws(blocks, ws))
    model.to(DEV)
    names = str(tmp_path / "coco.names")
    with open(names, "w") as fh:
        fh.write("\n".join(f"c{i}" for i in range(80)) + "\n")
    det = ImageDetector(model, names, thres=0.5, nms_thres=0.4, win_size=tuple(int(v) for v in win), overlap=ov, half=True)
    got = det.detect(img).cpu().numpy()
    ref = g["dets"]
    assert got.shape == ref.shape, f"{len(got)} detections vs {len(ref)} in the reference"
    assert (np.diff(got[:, 4]) <= 0).all(), "output must be score-descending"
    # every window's detections sit on the same ladder of scores, so rows of DIFFERENT windows tie up to rounding: pair the two
    # sets by their box centres (exact in every arithmetic: saturated, then scaled and shifted with the same fp32 operations)
    cg, cr = 0.5 * (got[:, :2] + got[:, 2:4]), 0.5 * (ref[:, :2] + ref[:, 2:4])
    d = np.abs(cg[None] - cr[:, None]).max(-1)
    perm = d.argmin(1)
    assert d.min(1).max() < 1e-2 and len(set(perm.tolist())) == len(perm), "not the reference's set of detections"
    got = got[perm]
    np.testing.assert_array_equal(got[:, 5], ref[:, 5])
    size = np.abs((got[:, 2:4] - got[:, :2]) - (ref[:, 2:4] - ref[:, :2])).max()
    print("sliding window: %d detections from 6 windows, size err %.3g px, score err %.3g" % (len(ref), size, np.abs(got[:, 4] - ref[:, 4]).max()))
    assert size <= 0.3 and np.abs(got[:, 4] - ref[:, 4]).max() < 1e-2
    # a frame smaller than the window takes the single-window path (img_detect.py:68)
    small = img[:300, :350]
    a = det.detect(small)
    b = ImageDetector(model, names, thres=0.5, nms_thres=0.4, half=True).detect(small)
    assert (a is None and b is None) or torch.equal(a, b)
