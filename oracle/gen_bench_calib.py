"""Produce bench_data/*.npz: the BatchNorm running statistics and YOLO head biases that make bench.py's seeded
random-weight models well-conditioned on its synthetic clip (~50 detections per frame at thres 0.5 / nms 0.4), plus the
ReID net's BatchNorm statistics measured on the crops of those detections.

Build-container only:   python -m oracle.gen_bench_calib [yolov3 608]
TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).  Runs the CPU oracle over workload.py's frames; bench.py
only LOADS the resulting fixture, so its product arm never touches oracle/.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

import workload as W
from . import darknet_ref as D
from . import reid_ref as R
from .cv_resize_ref import crops_to_batch
from .synth import calibrate_heads, frame_to_input


# detections per frame the heads are calibrated for: ~50 on the headline config; yolov4's 110-conv Mish stack carries ~2.5x the
# rounding noise at its heads, and the noise of a fitted row grows like n^1.5, so its score ladder only holds for ~30 rungs
WANT_DETS = {"yolov3": 50, "yolov4": 30}


def calibrate_darknet(cfg_name, size, seed=0, want_dets=None):
    want_dets = WANT_DETS.get(cfg_name, 50) if want_dets is None else want_dets
    blocks = D.parse_cfg(os.path.join(W.ROOT, "config", cfg_name + ".cfg"))
    defs = blocks[1:]
    ws = W.shape_heads(W.init_darknet_weights(defs, seed))
    frames = W.scenes(size, size)
    xs = torch.cat([frame_to_input(f) for f in frames], 0)
    D.forward(blocks, ws, xs, calibrate_bn=True)                    # BN statistics over all scenes of the clip
    # objectness rows with threshold / score-order / NMS margins on every scene (so that fp16-vs-fp32 rounding noise cannot
    # flip a decision and the track ids of both arms are comparable bit for bit)
    ws, info = calibrate_heads(blocks, ws, frames, want_dets=want_dets, conf_thres=W.DETECT_KW["thres"],
                               iou_thres=W.DETECT_KW["nms_thres"], verbose=True)
    box_obj = W.head_box_obj_rows()
    out = {}
    for i, e in enumerate(ws):
        if "bn" in e:
            out[f"bn_mean_{i}"], out[f"bn_var_{i}"] = e["bn"][2].astype(np.float32), e["bn"][3].astype(np.float32)
        else:
            out[f"head_bias_{i}"] = e["b"].astype(np.float32)
            out[f"head_rows_w_{i}"] = e["w"][box_obj, :, 0, 0].astype(np.float32)
    dets = [D.detect(blocks, ws, f, (size, size), W.DETECT_KW["thres"], W.DETECT_KW["nms_thres"]) for f in frames]
    print(cfg_name, size, info)
    os.makedirs(W.DATA, exist_ok=True)
    np.savez_compressed(os.path.join(W.DATA, f"{cfg_name}_{size}_seed{seed}.npz"), **out)
    # the fixture must reproduce the calibrated model through workload.py alone
    _, ws2 = W.darknet_workload(cfg_name, size, seed)
    for a, b in zip(ws, ws2):
        for k in a:
            assert all(np.array_equal(x, y) for x, y in zip(a[k], b[k])) if k == "bn" else np.array_equal(a[k], b[k]), k
    return frames, dets


def calibrate_reid(frames, dets, seed=0):
    sd = W.init_reid_state_dict(seed)
    crops = []
    for f, d in zip(frames, dets):
        tlwh, conf, cls = D.to_tracker_inputs(d, W.DETECT_KW["class_mask"])
        crops.append(crops_to_batch(f, tlwh))
    x = torch.as_tensor(np.concatenate([np.asarray(c) for c in crops], 0))
    print("reid calibration batch", tuple(x.shape))

    def calib(t, p):
        sd[p + ".running_mean"] = t.mean(dim=(0, 2, 3))
        sd[p + ".running_var"] = t.var(dim=(0, 2, 3), unbiased=False) + 1e-3
        return F.batch_norm(t, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.1, 1e-5)

    with torch.no_grad():
        x = F.max_pool2d(F.relu(calib(F.conv2d(x, sd["conv.0.weight"], sd["conv.0.bias"], 1, 1), "conv.1")), 3, 2, 1)
        for li, cin, cout, down in W.REID_STAGES:
            for bi in range(2):
                p = f"layer{li}.{bi}"
                s = 2 if (bi == 0 and down) else 1
                y = F.relu(calib(F.conv2d(x, sd[p + ".conv1.weight"], None, s, 1), p + ".bn1"))
                y = calib(F.conv2d(y, sd[p + ".conv2.weight"], None, 1, 1), p + ".bn2")
                if bi == 0 and down:
                    x = calib(F.conv2d(x, sd[p + ".downsample.0.weight"], None, 2, 0), p + ".downsample.1")
                x = F.relu(x + y)
    out = {}
    for n in W.reid_bn_names():
        out[n + ".running_mean"] = sd[n + ".running_mean"].numpy().astype(np.float32)
        out[n + ".running_var"] = sd[n + ".running_var"].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(W.DATA, f"reid_seed{seed}.npz"), **out)
    # appearance sanity: features of different detections must not be collinear
    f = R.extract(sd, frames[0], D.to_tracker_inputs(dets[0], W.DETECT_KW["class_mask"])[0])
    f = np.asarray(f)
    c = f @ f.T
    print("reid feature cosine between distinct crops: median %.3f max %.3f" % (np.median(c[~np.eye(len(c), dtype=bool)]), c[~np.eye(len(c), dtype=bool)].max()))


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "yolov3"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 608
    torch.set_num_threads(os.cpu_count())
    frames, dets = calibrate_darknet(cfg, size)
    if cfg == "yolov3":                      # one ReID fixture: the headline config's crops
        calibrate_reid(frames, dets)


if __name__ == "__main__":
    main()
