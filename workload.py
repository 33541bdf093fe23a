"""Synthetic workload for bench.py (BASELINE.json configs[1]: yolov3 608x608 + DeepSort, one stream, ~50 detections/frame).

No datasets or pretrained weights exist offline, so the bench runs seeded random-init weights of the reference
architectures on a seeded synthetic clip -- and both arms of bench.py (this repo's CUDA path and the CPU reference arm)
consume exactly the same frames and weights, built here.  This module does NOT import `oracle/`: seeded random weights alone
would blow up or collapse through 75 layers and produce either zero or thousands of detections, so the BatchNorm running
statistics and the three YOLO head biases that make the clip yield ~50 detections per frame are a small committed fixture
(`bench_data/*.npz`), produced once in the build container by `oracle/gen_bench_calib.py` (which runs the CPU oracle over
these same frames) and merely loaded here.
"""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(ROOT, "bench_data")

# the reference demo's parameters (video_deepsort.py:18-45)
TRACKER_KW = dict(max_dist=0.3, min_confidence=1, max_iou_distance=0.7, max_age=30, n_init=3, nn_budget=30)
DETECT_KW = dict(thres=0.5, nms_thres=0.4, class_mask=[0, 2, 4])
# The clip: (scene, frames it is held) in order, cycled.  Held scenes because a seeded random-weight detector is not a detector:
# moving the rectangles by ONE pixel leaves 0 of 47 detections in place, so on a moving clip no track would ever
# be confirmed.  The schedule exercises confirmation (holds >= n_init), tentative tracks that die (holds of 2), re-identification
# of tracks that were missing for fewer than max_age frames (scenes 0, 1, 2 come back) and deletion by age.
SCHEDULE = ((0, 8), (1, 8), (0, 6), (2, 2), (3, 8), (1, 6), (4, 8), (5, 2), (2, 8), (6, 8))
N_SCENES = 1 + max(s for s, _ in SCHEDULE)
CLIP_LEN = sum(n for _, n in SCHEDULE)                     # 64 frames per cycle
_CLIP = [s for s, n in SCHEDULE for _ in range(n)]
BETA_MEAN = 1.5                # BN beta ~ N(1.5, 0.1): the ordered edge of the random BN + leaky stack (see init_darknet_weights)


def make_scene(h, w, seed, n_rect=50):
    """Noise background plus `n_rect` textured person-sized rectangles, uint8 RGB (h,w,3).  Every scene shows the SAME
    `n_rect` objects (sizes, colours, textures come from a fixed stream) at scene-specific positions over scene-specific
    noise, so all scenes share their global image statistics -- a random-weight detector then fires about equally often
    on each of them."""
    objs = np.random.default_rng(4242)
    rng = np.random.default_rng(1000 + seed)
    img = rng.integers(0, 64, (h, w, 3), dtype=np.uint8)
    s = min(h, w) / 608.0
    for _ in range(n_rect):
        rw, rh = int(objs.integers(30, 71) * s) + 2, int(objs.integers(60, 141) * s) + 2
        base = objs.integers(64, 256, 3)
        tex = objs.integers(-32, 33, (rh, rw, 3))
        x, y = int(rng.integers(0, max(1, w - rw))), int(rng.integers(0, max(1, h - rh)))
        img[y:y + rh, x:x + rw] = np.clip(base[None, None, :] + tex, 0, 255).astype(np.uint8)
    return img


def scenes(h=608, w=608, seeds=None):
    """The clip's distinct scenes (the detector heads in bench_data/ were calibrated on exactly these)."""
    if seeds is None:
        seeds = range(N_SCENES)
    return [make_scene(h, w, int(s)) for s in seeds]


def clip_index(t):
    """Which scene frame `t` of the endless clip shows."""
    return _CLIP[t % CLIP_LEN]


# ---------------------------------------------------------------------------------------------------------------------
# Darknet weights
# ---------------------------------------------------------------------------------------------------------------------
def _channels(module_defs):
    out, prev = [], 3
    for i, d in enumerate(module_defs):
        t = d["type"]
        if t == "convolutional":
            prev = int(d["filters"])
        elif t == "route":
            idx = [int(x) for x in d["layers"].split(",")]
            prev = sum(out[j if j >= 0 else i + j] for j in idx)
            if "groups" in d:
                prev //= int(d["groups"])
        elif t == "shortcut":
            prev = out[i + int(d["from"])]
        out.append(prev)
    return out


def init_darknet_weights(module_defs, seed=0):
    """One dict per conv block in cfg order: {'w': (Cout,Cin,k,k), 'bn': [gamma,beta,mean,var] | 'b': bias}.
    BN beta ~ N(BETA_MEAN, 0.1): with beta = 0 a randomly initialised BN + leaky-ReLU stack is in the chaotic phase (mean-field
    perturbation gain 1.34 per layer -- rounding noise doubles every ~2.4 layers, which no trained network does); beta = 1.5
    puts it at the ordered edge (gain ~1.02) without changing a FLOP, so fp16-vs-fp32 differences reflect the arithmetic."""
    g = torch.Generator().manual_seed(seed)
    ch = _channels(module_defs)
    ws = []
    for i, d in enumerate(module_defs):
        if d["type"] != "convolutional":
            continue
        cin = 3 if i == 0 else ch[i - 1]
        cout, k = int(d["filters"]), int(d["size"])
        w = torch.randn(cout, cin, k, k, generator=g) * float(np.sqrt(2.0 / (cin * k * k)))
        e = {"w": w.numpy().copy()}
        if int(d["batch_normalize"]):
            gamma = 1.0 + 0.1 * torch.randn(cout, generator=g)
            beta = BETA_MEAN + 0.1 * torch.randn(cout, generator=g)
            e["bn"] = [gamma.numpy().copy(), beta.numpy().copy(), np.zeros(cout, np.float32), np.ones(cout, np.float32)]
        else:
            e["b"] = np.zeros(cout, np.float32)
        ws.append(e)
    return ws


def head_rows(nc=80, na=3):
    obj = [a * (nc + 5) + 4 for a in range(na)]
    cls0 = [a * (nc + 5) + 5 for a in range(na)]
    other = [a * (nc + 5) + 5 + c for a in range(na) for c in range(1, nc)]
    wh = [a * (nc + 5) + k for a in range(na) for k in (2, 3)]
    return obj, cls0, other, wh


def head_box_obj_rows(nc=80, na=3):
    """tx, ty, tw, th and objectness rows of the three anchors (the rows the margin calibration refits)."""
    return [a * (nc + 5) + k for a in range(na) for k in range(5)]


def shape_heads(ws):
    """Deterministic head shaping (before calibration): class 0 wins everywhere, box sizes stay tame."""
    obj, cls0, other, wh = head_rows()
    for e in ws:
        if "b" not in e:
            continue
        e["w"][other] *= 0.05
        e["w"][cls0] *= 0.002              # (almost) constant class confidence: score order == objectness order
        e["w"][wh] *= 0.25
        e["b"][:] = 0
        e["b"][cls0] = 8.0
        e["b"][other] = -12.0
    return ws


def apply_calibration(ws, calib):
    """calib: npz with bn_mean_<i>, bn_var_<i> per BN conv i; head_bias_<i> and head_rows_w_<i> (the box and objectness rows
    of the three anchors, refitted with decision margins on the clip's scenes: oracle/synth.py fit_head_margins) per head
    conv i (cfg order)."""
    rows = head_box_obj_rows()
    for i, e in enumerate(ws):
        if "bn" in e:
            e["bn"][2] = calib[f"bn_mean_{i}"].astype(np.float32)
            e["bn"][3] = calib[f"bn_var_{i}"].astype(np.float32)
        else:
            e["b"] = calib[f"head_bias_{i}"].astype(np.float32)
            e["w"][rows, :, 0, 0] = calib[f"head_rows_w_{i}"].astype(np.float32)
    return ws


def flatten_darknet(ws):
    """The float32 payload of a darknet .weights file (yolo3/models/models.py:315-366 layout)."""
    flat = []
    for e in ws:
        if "bn" in e:
            g, b, m, v = e["bn"]
            flat += [b, g, m, v]
        else:
            flat.append(e["b"])
        flat.append(e["w"].ravel())
    return np.concatenate([np.asarray(a, np.float32).ravel() for a in flat])


def darknet_workload(cfg_name="yolov3", size=608, seed=0):
    """(module_defs incl. [net] stripped, ws) for the calibrated bench model."""
    from yolo_deepsort_b200.darknet import parse_model_config
    defs = parse_model_config(os.path.join(ROOT, "config", cfg_name + ".cfg"))[1:]
    ws = shape_heads(init_darknet_weights(defs, seed))
    calib = np.load(os.path.join(DATA, f"{cfg_name}_{size}_seed{seed}.npz"))
    return defs, apply_calibration(ws, calib)


# ---------------------------------------------------------------------------------------------------------------------
# ReID weights (deep_sort/deep/model.py:48-95 state_dict names)
# ---------------------------------------------------------------------------------------------------------------------
REID_STAGES = ((1, 64, 64, False), (2, 64, 128, True), (3, 128, 256, True), (4, 256, 512, True))


def reid_bn_names():
    names = ["conv.1"]
    for li, _, _, down in REID_STAGES:
        for bi in range(2):
            p = f"layer{li}.{bi}"
            names += [p + ".bn1", p + ".bn2"]
            if bi == 0 and down:
                names.append(p + ".downsample.1")
    return names


def init_reid_state_dict(seed=0):
    g = torch.Generator().manual_seed(10_000 + seed)
    sd = {}

    def conv(name, cout, cin, k, bias=False):
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * float(np.sqrt(2.0 / (cin * k * k)))
        if bias:
            sd[name + ".bias"] = 0.1 * torch.randn(cout, generator=g)

    def bn(name, c):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
        sd[name + ".running_mean"] = torch.zeros(c)
        sd[name + ".running_var"] = torch.ones(c)

    conv("conv.0", 64, 3, 3, bias=True)
    bn("conv.1", 64)
    for li, cin, cout, down in REID_STAGES:
        for bi in range(2):
            p = f"layer{li}.{bi}"
            conv(p + ".conv1", cout, cin if bi == 0 else cout, 3)
            bn(p + ".bn1", cout)
            conv(p + ".conv2", cout, cout, 3)
            bn(p + ".bn2", cout)
            if bi == 0 and down:
                conv(p + ".downsample.0", cout, cin, 1)
                bn(p + ".downsample.1", cout)
    return sd


def reid_workload(seed=0):
    sd = init_reid_state_dict(seed)
    calib = np.load(os.path.join(DATA, f"reid_seed{seed}.npz"))
    for n in reid_bn_names():
        sd[n + ".running_mean"] = torch.from_numpy(calib[n + ".running_mean"].astype(np.float32))
        sd[n + ".running_var"] = torch.from_numpy(calib[n + ".running_var"].astype(np.float32))
    return sd
