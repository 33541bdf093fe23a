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
from .synth import frame_to_input


def calibrate_darknet(cfg_name, size, seed=0, target=56, n_pool=12):
    blocks = D.parse_cfg(os.path.join(W.ROOT, "config", cfg_name + ".cfg"))
    defs = blocks[1:]
    ws = W.shape_heads(W.init_darknet_weights(defs, seed))
    pool = list(range(n_pool))
    frames = W.scenes(size, size, seeds=pool)
    xs = torch.cat([frame_to_input(f) for f in frames], 0)
    D.forward(blocks, ws, xs, calibrate_bn=True)                    # BN statistics over the whole candidate pool
    obj = W.head_rows()[0]
    heads = [i for i, e in enumerate(ws) if "b" in e]
    _, outs = D.forward(blocks, ws, xs, return_layers=True)
    per_frame = []
    for n in range(len(frames)):
        per = [outs[li - 1][n][obj].reshape(-1).numpy() for li, b in enumerate(defs) if b["type"] == "yolo"]
        per_frame.append(np.concatenate(per))
    # provisional cut: the median candidate scene passes `target` rows; keep the N_SCENES scenes closest to the target,
    # then place the final cut in the widest logit gap of the kept scenes' pooled logits
    cut0 = float(np.median([np.sort(l)[::-1][target] for l in per_frame]))
    counts = np.array([int((l > cut0).sum()) for l in per_frame])
    keep = np.argsort(np.abs(counts - target), kind="stable")[:W.N_SCENES]
    keep = np.sort(keep)
    print("candidate counts", counts.tolist(), "-> scenes", keep.tolist())
    frames = [frames[i] for i in keep]
    per_frame = [per_frame[i] for i in keep]
    pooled = np.sort(np.concatenate(per_frame))[::-1]
    want = target * len(frames)
    lo, hi = int(want * 0.93), int(want * 1.07)
    gaps = pooled[lo - 1:hi - 1] - pooled[lo:hi]
    k = int(np.argmax(gaps)) + lo
    cut = 0.5 * (float(pooled[k - 1]) + float(pooled[k]))
    for hi_ in heads:
        ws[hi_]["b"][obj] = np.float32(-cut)                         # sigmoid(t - cut) > 0.5  <=>  t > cut
    out = {"scene_seeds": np.asarray([pool[i] for i in keep], np.int32)}
    for i, e in enumerate(ws):
        if "bn" in e:
            out[f"bn_mean_{i}"], out[f"bn_var_{i}"] = e["bn"][2].astype(np.float32), e["bn"][3].astype(np.float32)
        else:
            out[f"head_bias_{i}"] = e["b"].astype(np.float32)
    dets = [D.detect(blocks, ws, f, (size, size), W.DETECT_KW["thres"], W.DETECT_KW["nms_thres"]) for f in frames]
    info = dict(cut=cut, gap=float(gaps.max()), n_pass=[int((l > cut).sum()) for l in per_frame],
                n_dets=[0 if d is None else len(d) for d in dets], logit_std=float(np.std(pooled)))
    print(cfg_name, size, info)
    os.makedirs(W.DATA, exist_ok=True)
    np.savez_compressed(os.path.join(W.DATA, f"{cfg_name}_{size}_seed{seed}.npz"), **out)
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
    calibrate_reid(frames, dets)


if __name__ == "__main__":
    main()
