"""Seeded synthetic inputs for the oracle, the parity tests and bench.py (no datasets or pretrained
weights exist offline).  TEST/BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).

Nothing here restates reference code; it only manufactures inputs with the reference's shapes:
frames (HxWx3 uint8 RGB), darknet weight lists, a ReID checkpoint dict, and association scenarios.
"""
import numpy as np
import torch

from . import darknet_ref, reid_ref


def make_frame(h, w, seed=0, n_rect=50):
    """Noise background plus `n_rect` textured rectangles (w in [30,70], h in [60,140], scaled to the
    frame), uint8 RGB."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 64, (h, w, 3), dtype=np.uint8)
    s = min(h, w) / 608.0
    for _ in range(n_rect):
        rw, rh = int(rng.integers(30, 71) * s) + 2, int(rng.integers(60, 141) * s) + 2
        x, y = int(rng.integers(0, max(1, w - rw))), int(rng.integers(0, max(1, h - rh)))
        base = rng.integers(64, 256, 3)
        tex = rng.integers(-32, 33, (rh, rw, 3))
        img[y:y + rh, x:x + rw] = np.clip(base[None, None, :] + tex, 0, 255).astype(np.uint8)
    return img


def frame_to_input(frame):
    """ImageDetector's tensor prep (yolo3/detect/img_detect.py:71-79): (1,3,H,W) float32 in [0,1]."""
    return (torch.from_numpy(np.ascontiguousarray(frame)).permute(2, 0, 1) / 255.).unsqueeze(0)


def head_channels(nc, na=3):
    obj = [a * (nc + 5) + 4 for a in range(na)]
    cls0 = [a * (nc + 5) + 5 for a in range(na)]
    other = [a * (nc + 5) + 5 + c for a in range(na) for c in range(1, nc)]
    return obj, cls0, other


def darknet_weights(blocks, frames, seed=0, target=50, conf_thres=0.5):
    """Seeded weights whose BN statistics are calibrated on frames[0] and whose head biases are set so
    that, on every frame in `frames`, about `target` candidates clear `conf_thres` as class 0 only --
    with the objectness cut placed in the widest gap of the pooled logits (so reduced-precision
    arithmetic does not flip a borderline candidate).  Returns (ws, info)."""
    ws = darknet_ref.init_weights(blocks, seed)
    darknet_ref.forward(blocks, ws, frame_to_input(frames[0]), calibrate_bn=True)
    conv_blocks = [b for b in blocks[1:] if b["type"] == "convolutional"]
    heads = [i for i, b in enumerate(conv_blocks) if not int(b["batch_normalize"])]
    nc = 80
    obj, cls0, other = head_channels(nc)
    for hi in heads:                               # well-separated class scores, tame box sizes
        d = ws[hi]
        d["b"][:] = 0
        d["b"][cls0] = 8.0
        d["b"][other] = -12.0
        d["w"][other] *= 0.05
        d["w"][cls0] *= 0.05
        wh = [a * (nc + 5) + k for a in range(3) for k in (2, 3)]
        d["w"][wh] *= 0.25
    # pooled objectness logits over all frames and heads (bias currently 0)
    logits = []
    for f in frames:
        _, outs = darknet_ref.forward(blocks, ws, frame_to_input(f), return_layers=True)
        per = []
        for li, b in enumerate(blocks[1:]):
            if b["type"] == "yolo":
                raw = outs[li - 1][0]                 # (255, g, g) head conv output
                per.append(raw[obj].reshape(-1).numpy())
        logits.append(np.concatenate(per))
    pooled = np.sort(np.concatenate(logits))[::-1]
    want = target * len(frames)
    lo, hi = max(1, int(want * 0.8)), min(len(pooled) - 1, int(want * 1.25))
    gaps = pooled[lo - 1:hi - 1] - pooled[lo:hi]
    k = int(np.argmax(gaps)) + lo                    # cut between pooled[k-1] and pooled[k]
    cut = 0.5 * (float(pooled[k - 1]) + float(pooled[k]))
    logit_thr = float(np.log(conf_thres / (1 - conf_thres)))
    for hi_ in heads:
        ws[hi_]["b"][obj] = np.float32(logit_thr - cut)
    info = dict(cut=cut, gap=float(gaps.max()), n_pass=[int((l > cut).sum()) for l in logits],
                logit_scale=float(np.std(pooled)))
    return ws, info


def reid_state_dict(seed=0, calib_batch=None):
    """Seeded ReID checkpoint; BN running stats calibrated on `calib_batch` ((m,3,128,64) f32) so the
    20-layer residual stack keeps unit scale."""
    sd = reid_ref.init_state_dict(seed)
    if calib_batch is None:
        g = torch.Generator().manual_seed(seed + 1)
        calib_batch = torch.randn(8, 3, 128, 64, generator=g)
    import torch.nn.functional as F
    x = torch.as_tensor(calib_batch)

    def calib(t, p):
        sd[p + ".running_mean"] = t.mean(dim=(0, 2, 3))
        sd[p + ".running_var"] = t.var(dim=(0, 2, 3), unbiased=False) + 1e-3
        return F.batch_norm(t, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            False, 0.1, 1e-5)

    with torch.no_grad():
        x = F.max_pool2d(F.relu(calib(F.conv2d(x, sd["conv.0.weight"], sd["conv.0.bias"], 1, 1), "conv.1")), 3, 2, 1)
        for li, cin, cout, down in reid_ref.STAGES:
            for bi in range(2):
                p = f"layer{li}.{bi}"
                s = 2 if (bi == 0 and down) else 1
                y = F.relu(calib(F.conv2d(x, sd[p + ".conv1.weight"], None, s, 1), p + ".bn1"))
                y = calib(F.conv2d(y, sd[p + ".conv2.weight"], None, 1, 1), p + ".bn2")
                if bi == 0 and down:
                    x = calib(F.conv2d(x, sd[p + ".downsample.0.weight"], None, 2, 0), p + ".downsample.1")
                x = F.relu(x + y)
    return sd


def unit_rows(rng, n, d=512):
    v = rng.standard_normal((n, d)).astype(np.float32)
    return v / np.linalg.norm(v, axis=1, keepdims=True)


class Scenario:
    """Association scenario: `n` objects with constant-velocity boxes and a fixed unit appearance
    vector each; per frame a subset is observed with box jitter and appearance noise, plus novel
    objects appearing and old ones leaving.  Yields (tlwh (m,4) f32, feats (m,512) f32, cls (m,) f32)."""

    def __init__(self, n=50, frame_hw=(608, 608), seed=0, p_miss=0.05, p_new=0.02, p_leave=0.01,
                 feat_noise=0.05, box_jitter=0.5):
        self.rng = np.random.default_rng(seed)
        self.hw = frame_hw
        self.p_miss, self.p_new, self.p_leave, self.feat_noise, self.jit = p_miss, p_new, p_leave, feat_noise, box_jitter
        self.objs = [self._spawn() for _ in range(n)]

    def _spawn(self):
        r = self.rng
        H, W = self.hw
        w, h = r.uniform(30, 70), r.uniform(60, 140)
        return dict(x=r.uniform(0, W - w), y=r.uniform(0, H - h), w=w, h=h,
                    vx=r.uniform(-3, 3), vy=r.uniform(-3, 3), f=unit_rows(r, 1)[0], c=float(r.choice([0, 2, 4])))

    def step(self):
        r = self.rng
        H, W = self.hw
        n0 = len(self.objs)
        self.objs = [o for o in self.objs if r.random() > self.p_leave]
        for _ in range(r.binomial(max(n0, 1), self.p_new)):
            self.objs.append(self._spawn())
        tl, ft, cl = [], [], []
        for o in self.objs:
            o["x"] = float(np.clip(o["x"] + o["vx"], 0, W - o["w"] - 1))
            o["y"] = float(np.clip(o["y"] + o["vy"], 0, H - o["h"] - 1))
            if r.random() < self.p_miss:
                continue
            j = r.uniform(-self.jit, self.jit, 4)
            tl.append([o["x"] + j[0], o["y"] + j[1], o["w"] + j[2], o["h"] + j[3]])
            f = o["f"] + self.feat_noise * r.standard_normal(512).astype(np.float32) / np.sqrt(512)
            ft.append(f / np.linalg.norm(f))
            cl.append(o["c"])
        order = r.permutation(len(tl))
        tl = np.asarray(tl, np.float32).reshape(-1, 4)[order]
        ft = np.asarray(ft, np.float32).reshape(-1, 512)[order]
        cl = np.asarray(cl, np.float32)[order]
        return tl, ft, cl
