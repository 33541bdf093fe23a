"""The reference-side binding: the package names video_deepsort.py imports resolve to the B200 implementations (CPU tier),
and the entry script's own flow -- Darknet(cfg).load_darknet_weights(file), DeepSort(ckpt.t7, ...), VideoDetector(...).detect(video)
-- runs through those names on a clip and matches the oracle (GPU tier)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

DROPIN = os.path.join(ROOT, "dropin")


@pytest.fixture
def dropin_path():
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in ("yolo3", "deep_sort", "action")}
    sys.path.insert(0, DROPIN)
    yield
    sys.path.remove(DROPIN)
    for k in list(sys.modules):
        if k.split(".")[0] in ("yolo3", "deep_sort", "action"):
            del sys.modules[k]
    sys.modules.update(saved)


def test_reference_import_lines_resolve(dropin_path):
    import yolo_deepsort_b200 as Y
    from deep_sort import DeepSort, build_tracker                      # video_deepsort.py:5
    from yolo3.detect.video_detect import VideoDetector                # video_deepsort.py:6
    from yolo3.models import Darknet                                   # video_deepsort.py:7
    from yolo3.detect.img_detect import ImageDetector
    from yolo3.utils.model_build import soft_non_max_suppression, resize_boxes, p1p2Toxywh
    from yolo3.utils.parse_config import parse_model_config
    from action.action_Identify import ActionIdentify                  # video_deepsort.py:3-4
    from action import actions as rules
    assert ActionIdentify is Y.ActionIdentify and all(hasattr(rules, n) for n in ("TakeOff", "Landing", "Glide", "FastCrossing", "BreakInto"))
    assert Darknet is Y.Darknet and DeepSort is Y.DeepSort and VideoDetector is Y.VideoDetector and ImageDetector is Y.ImageDetector
    assert callable(build_tracker) and callable(soft_non_max_suppression) and callable(resize_boxes) and callable(p1p2Toxywh)
    defs = parse_model_config(os.path.join(ROOT, "config", "yolov3-tiny.cfg"))
    assert defs[0]["type"] == "net" and sum(d["type"] == "convolutional" for d in defs) == 13


@pytest.mark.gpu
def test_entry_script_flow_through_dropin_names(dropin_path, tmp_path):
    import cv2
    from oracle import darknet_ref as D
    from oracle.synth import darknet_weights, make_frame, reid_state_dict
    from test_gpu_pipeline import oracle_run
    from deep_sort import DeepSort
    from yolo3.detect.video_detect import VideoDetector
    from yolo3.models import Darknet

    cfg = os.path.join(ROOT, "config", "yolov3-tiny.cfg")
    scenes = [make_frame(416, 416, seed=s) for s in (0, 1)]
    blocks = D.parse_cfg(cfg)
    ws, _ = darknet_weights(blocks, scenes, seed=0, target=50)
    wpath, ckpt, names, video = (str(tmp_path / n) for n in ("tiny.weights", "ckpt.t7", "coco.names", "clip.avi"))
    D.write_weights(wpath, blocks, ws)                                  # darknet .weights file, as the reference loads it
    sd = reid_state_dict(seed=0)
    torch.save({"net_dict": sd, "acc": 0.0, "epoch": 0}, ckpt)          # deep_sort/deep/train.py:137-144 checkpoint layout
    with open(names, "w") as fh:
        fh.write("\n".join(f"c{i}" for i in range(80)) + "\n")
    clip = [scenes[0]] * 9
    wr = cv2.VideoWriter(video, cv2.VideoWriter_fourcc(*"FFV1"), 25, (416, 416))
    for f in clip:
        wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    wr.release()

    # ---- the body of video_deepsort.py:13-52, names and keyword arguments as written there ----
    model = Darknet(cfg, img_size=(416, 416))
    model.load_darknet_weights(wpath)
    model.to("cuda:0")
    tracker = DeepSort(ckpt, min_confidence=1, use_cuda=True, nn_budget=30, n_init=3, max_iou_distance=0.7, max_dist=0.3, max_age=30)
    video_detector = VideoDetector(model, names, thickness=2, skip_frames=2, thres=0.5, class_mask=[0, 2, 4], nms_thres=0.4,
                                   tracker=tracker, half=True)
    # skip_frames=2 (video_deepsort.py:41): the detector + tracker run on every second frame, the rows are held in between
    # (yolo3/detect/video_detect.py:132-157)
    ref_steps, _ = oracle_run(blocks, ws, sd, clip[::2])
    n = 0
    for t, (image, detections, _) in enumerate(video_detector.detect(video, real_show=False, skip_secs=0, show_fps=False)):
        assert image.shape == (416, 416, 3), f"frame {t}"
        np.testing.assert_array_equal(np.asarray(detections, np.int32).reshape(-1, 6), np.asarray(ref_steps[t // 2], np.int32).reshape(-1, 6),
                                      err_msg=f"frame {t}: rows [x1,y1,x2,y2,id,cls] differ from the oracle")
        n += 1
    assert n == len(clip)
