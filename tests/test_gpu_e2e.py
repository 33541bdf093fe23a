"""End-to-end parity gate on the workload bench.py times ("(K,6) outputs compared to the oracle before timing
counts"): workload.py's clip through FramePipeline at micro-batch 1 and 8, every stage compared with the oracle on the real
data flow (oracle/clip.py ParityCheck).  yolov3-608 is configs[1] (the headline), yolov4-608 configs[2]."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from util import DEV

pytestmark = pytest.mark.gpu


def run_clip(cfg, mb, n_frames):
    import workload as W
    from oracle import darknet_ref as D
    from oracle.clip import ParityCheck
    from yolo_deepsort_b200 import Darknet, DeepSort, FramePipeline
    blocks = D.parse_cfg(os.path.join(ROOT, "config", cfg + ".cfg"))
    _, ws = W.darknet_workload(cfg, 608)
    sd = W.reid_workload()
    scenes = W.scenes(608, 608)
    model = Darknet(os.path.join(ROOT, "config", cfg + ".cfg"), img_size=(608, 608))
    model.set_weights(W.flatten_darknet(ws))
    model.to(DEV)
    ds = DeepSort(sd, use_cuda=True, device=DEV, **W.TRACKER_KW)
    pipe = FramePipeline(model, ds, W.DETECT_KW["thres"], W.DETECT_KW["nms_thres"], W.DETECT_KW["class_mask"], micro_batch=mb)
    chk = ParityCheck(blocks, ws, sd, scenes, W.DETECT_KW["thres"], W.DETECT_KW["nms_thres"], W.DETECT_KW["class_mask"], W.TRACKER_KW)
    frames_dev = [torch.from_numpy(s).to(DEV) for s in scenes]
    sub = col = 0
    while col < n_frames:                               # look-ahead form: exactly what bench.py's timed loop does
        while sub < n_frames and pipe.can_submit():
            pipe.submit(frames_dev[W.clip_index(sub)]); sub += 1
        rows, dets = pipe.collect()
        chk.frame(W.clip_index(col), dets, rows, pipe.last_inputs())
        col += 1
    return chk


@pytest.mark.parametrize("cfg,mb", [("yolov3", 1), ("yolov3", 8), ("yolov4", 1), ("yolov4", 8)])
def test_workload_parity(cfg, mb):
    import workload as W
    chk = run_clip(cfg, mb, W.CLIP_LEN + 8)             # one full cycle of the schedule and the start of the next
    s = chk.summary()
    print(cfg, "micro-batch", mb, s)
    assert s["detections_equal"], chk.problems[:5]
    if cfg == "yolov3":
        assert s["frames_with_other_detection_order"] == {"fp32": 0, "half": 0}, chk.problems[:5]
    # (yolov4, 110 convolutions with Mish: at ~50 detections per frame the rounding noise of the fitted objectness rows reaches
    #  the gap of the score ladder -- the fp32 and the fp16-storage ORACLES already disagree on the order of some neighbours,
    #  tests/test_precision_floor.py -- so there the order is reported, not asserted; the association is teacher-forced anyway)
    assert s["ids_equal"], chk.problems[:5]
    assert s["track_rows"] > 40 * W.CLIP_LEN * 0.6, "the clip should keep ~50 confirmed tracks alive"
    assert s["max_centre_px"] <= 1e-3                    # saturated centres: exact up to the fp32 rounding of x1 + w/2
    # sizes: the fp16-storage floor of these stacks is 0.73 px (yolov3) / 0.85 px (yolov4) (tests/test_precision_floor.py); measured on
    # B200 0.67 / 1.04 px -- pinned at measured x 1.25.  A corner moves by half of that, and the crop rectangles are asserted equal.
    assert s["max_size_px"] <= (0.85 if cfg == "yolov3" else 1.3) and s["max_score_err"] <= 2e-2
    # ReID: fp16 operands against the fp32 reference -- the storage-format floor is 1.2e-3 worst crop (tests/test_precision_floor.py),
    # measured 1.42e-3 on the clip's ~350 distinct crops
    assert s["max_feature_rel"] <= 1.8e-3
    # And from pixels, free-running: the track ids of the CUDA run equal the fp32 oracle's on every frame of the headline clip, IN
    # ORDER.  (Not a law of nature -- association at scene changes is sensitive to sub-pixel differences, and the fp16-storage
    # oracle parts ways with the fp32 one inside this very clip, `oracles_part_ways_at` -- but deterministic, so it is asserted for
    # the headline config: a regression gate.  yolov4's sizes carry ~1 px of rounding noise and its free run leaves the fp32
    # oracle's somewhere after the first scene changes; there the number is printed, the teacher-forced equality above is the gate.)
    if cfg == "yolov3":
        assert s["e2e_first_id_mismatch"]["fp32"] is None, s
    else:
        assert s["e2e_first_id_mismatch"]["fp32"] is None or s["e2e_first_id_mismatch"]["fp32"] >= 16, s
