"""CPU tier: what fp16 STORAGE alone costs against the fp32 reference, measured with the oracle (oracle/darknet_ref.py
half_storage=True restates the reference's own half=True mode, yolo3/detect/img_detect.py:48-50,79-82, with the rounding points of a
fused epilogue; oracle/reid_ref.py half_storage=True applies the same rounding points to the ReID net, which the reference
always runs in fp32).  These are the floors the GPU tier's tolerances are pinned to (north_star: 1e-3 relative): no
implementation that feeds fp16 operands to the tensor cores can be closer to the fp32 reference than this.  Also checks that the
decision margins of the calibrated synthetic detectors hold between the two arithmetics: same detections, same ORDER."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import darknet_ref as D
from oracle import reid_ref as R
from oracle.synth import darknet_weights, frame_to_input, make_frame, reid_state_dict


def _det_pair(blocks, ws, frame):
    x = frame_to_input(frame)
    a = D.postprocess(D.forward(blocks, ws, x)[0].numpy(), 0.5, 0.4)
    b = D.postprocess(D.forward(blocks, ws, x, half_storage=True)[0].numpy(), 0.5, 0.4)
    return a, b


def _box_stats(a, b, same_order=True):
    assert a.shape == b.shape, "fp16 storage changed the number of detections"
    ca, cb = 0.5 * (a[:, :2] + a[:, 2:4]), 0.5 * (b[:, :2] + b[:, 2:4])
    if not same_order:                                  # pair by the centres (exact in both arithmetics)
        d = np.abs(cb[None] - ca[:, None]).max(-1)
        perm = d.argmin(1)
        assert d.min(1).max() < 1e-2 and len(set(perm.tolist())) == len(perm), "fp16 storage changed the set of detections"
        b = b[perm]
    assert np.array_equal(a[:, 5], b[:, 5]) and np.abs(ca - 0.5 * (b[:, :2] + b[:, 2:4])).max() < 1e-2, \
        "fp16 storage changed the detections or their order"
    centre = np.abs(0.5 * (a[:, :2] + a[:, 2:4]) - 0.5 * (b[:, :2] + b[:, 2:4])).max()
    size = np.abs((a[:, 2:4] - a[:, :2]) - (b[:, 2:4] - b[:, :2])).max()
    rel = (np.abs(a[:, :4] - b[:, :4]).max(1) / np.minimum(a[:, 2] - a[:, 0], a[:, 3] - a[:, 1])).max()
    assert np.array_equal(a[:, :4].astype(np.int64), b[:, :4].astype(np.int64)), "crop rectangles differ"
    return float(centre), float(size), float(rel), float(np.abs(a[:, 4] - b[:, 4]).max())


def test_detector_floor_tiny416():
    """The golden's model (yolov3-tiny 416, 13 convs): fp16 storage moves box sizes by ~0.1 px = 2e-3 of the box size and scores
    by ~3e-3; centres are exact (saturated) and the order of the 51 detections is unchanged."""
    blocks = D.parse_cfg(os.path.join(ROOT, "config", "yolov3-tiny.cfg"))
    frames = [make_frame(416, 416, seed=s) for s in (0, 1)]
    ws, info = darknet_weights(blocks, frames, seed=0, target=50)
    g = np.load(os.path.join(GOLDEN, "tiny416.npz"))
    a, b = _det_pair(blocks, ws, frames[0])
    np.testing.assert_array_equal(a, g["dets"])                      # the fp32 oracle IS the reference (golden written by it)
    centre, size, rel, score = _box_stats(a, b)
    print("tiny416 fp16-storage floor: centre %.2g px, size %.3g px, %.3g of the box size, score %.3g" % (centre, size, rel, score))
    assert centre <= 1e-4 and 0.02 < size < 0.25 and 5e-4 < rel < 4e-3 and score < 1e-2


@pytest.mark.parametrize("cfg", ["yolov3", "yolov4"])
def test_detector_margins_hold_on_the_bench_workload(cfg):
    """workload.py's calibrated 608x608 models: on every scene of the clip the fp32 and the fp16-storage oracle return the same
    detections (yolov3: in the same order) with identical crop rectangles (75 / 110 convolutions deep)."""
    import workload as W
    torch.set_num_threads(os.cpu_count())
    blocks = D.parse_cfg(os.path.join(ROOT, "config", cfg + ".cfg"))
    _, ws = W.darknet_workload(cfg, 608)
    worst = np.zeros(4)
    for f in W.scenes(608, 608):
        a, b = _det_pair(blocks, ws, f)
        assert 25 <= len(a) <= 60
        # yolov4 (110 convs, Mish): neighbours on the score ladder may swap between the two arithmetics -- same set, paired by centre
        worst = np.maximum(worst, _box_stats(a, b, same_order=(cfg == "yolov3")))
    print("%s-608 fp16-storage floor: centre %.2g px, size %.3g px, %.3g of the box size, score %.3g" % ((cfg,) + tuple(worst)))
    assert worst[0] <= 1e-4 and worst[1] < 0.95                 # a corner moves by half the size error: below the half-pixel margin


def test_reid_floor():
    """ReID features, fp16 operands vs fp32 (the reference): relative L2 error ~1.2e-3 median, ~1.9e-3 worst crop -- above the
    north-star 1e-3 for ANY fp16-operand implementation of these weights; the appearance COSTS (pairwise cosine distances, what
    the association consumes) move by ~2e-4."""
    from oracle.cv_resize_ref import crops_to_batch
    sd = reid_state_dict(seed=0)
    frame = make_frame(608, 608, seed=5)
    rng = np.random.default_rng(11)
    m = 32
    tlwh = np.stack([rng.uniform(0, 540, m), rng.uniform(0, 470, m), rng.uniform(20, 90, m), rng.uniform(40, 160, m)], 1).astype(np.float32)
    x = crops_to_batch(frame, tlwh)
    a, b = R.net_forward(sd, x).numpy(), R.net_forward(sd, x, half_storage=True).numpy()
    rel = np.linalg.norm(a - b, axis=1) / np.linalg.norm(a, axis=1)
    dcost = np.abs((1 - a @ a.T) - (1 - b @ b.T)).max()
    print("ReID fp16-operand floor: rel L2 median %.3g max %.3g; cosine-distance error max %.3g" % (np.median(rel), rel.max(), dcost))
    assert 5e-4 < np.median(rel) < 2e-3 and rel.max() < 3e-3 and dcost < 1e-3
