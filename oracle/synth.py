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


def head_convs(blocks):
    """[(cfg layer index of the conv feeding each yolo layer, its number among the conv blocks)] in cfg order."""
    body = blocks[1:]
    conv_no, ci = {}, 0
    for li, b in enumerate(body):
        if b["type"] == "convolutional":
            conv_no[li] = ci
            ci += 1
    return [(li - 1, conv_no[li - 1]) for li, b in enumerate(body) if b["type"] == "yolo"]


def shape_heads(ws, nc=80, na=3):
    """Deterministic head shaping (before calibration): class 0 wins everywhere with an (almost) constant confidence, so the
    score order of the detections is the order of their objectness; box sizes stay tame."""
    obj, cls0, other = head_channels(nc, na)
    wh = [a * (nc + 5) + k for a in range(na) for k in (2, 3)]
    for d in ws:
        if "b" not in d:
            continue
        d["b"][:] = 0
        d["b"][cls0] = 8.0
        d["b"][other] = -12.0
        d["w"][other] *= 0.05
        d["w"][cls0] *= 0.002
        d["w"][wh] *= 0.25
    return ws


def _box_iou(a, b):
    x1 = np.maximum(a[:, None, 0], b[None, :, 0]); y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2]); y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None, :] - inter)


def _head_geometry(blocks, ws, frame):
    """Per head: (gy, gx, anchors [(w,h)] of its yolo layer)."""
    body = blocks[1:]
    _, outs = darknet_ref.forward(blocks, ws, frame_to_input(frame), return_layers=True)
    geo = []
    for li, _ in head_convs(blocks):
        yb = body[li + 1]
        mask = [int(v) for v in yb["mask"].split(",")]
        anc = [int(v) for v in yb["anchors"].split(",")]
        geo.append((int(outs[li].shape[2]), int(outs[li].shape[3]), [(anc[2 * i], anc[2 * i + 1]) for i in mask]))
    return geo


def _features_and_boxes(blocks, ws, frames):
    heads = head_convs(blocks)
    feats, boxes = [[] for _ in heads], []
    for f in frames:
        pred, outs = darknet_ref.forward(blocks, ws, frame_to_input(f), return_layers=True)
        p = pred[0].numpy()
        boxes.append(np.stack([p[:, 0] - p[:, 2] / 2, p[:, 1] - p[:, 3] / 2, p[:, 0] + p[:, 2] / 2, p[:, 1] + p[:, 3] / 2], 1))
        for h, (li, _) in enumerate(heads):
            x = outs[li - 1][0]
            feats[h].append(x.reshape(x.shape[0], -1).numpy().astype(np.float64).T)           # (cells, C)
    return feats, boxes


def fit_head_margins(blocks, ws, frames, target=50, conf_thres=0.5, iou_thres=0.4, m_thr=0.06, gap=0.04, m_iou=0.04,
                     max_box=0.25, seed=0, noise_metric=False, verbose=False):
    """Decision margins for the synthetic detector ("build e2e synthetic data with margins around thresholds").

    The three YOLO heads end in a 1x1 convolution, i.e. every box / objectness output is LINEAR in that head's weight row.
    Starting from the seeded random rows, the box and objectness rows of the nine (head, anchor) pairs get minimum-norm
    corrections after which, on every frame of `frames`:
      * `target` +- 10 % candidates (picked round-robin over the nine rows from each row's own upper tail on that frame,
        boxes no larger than max_box of the frame and pairwise below IoU iou_thres - m_iou) are the frame's detections;
      * a detection's centre is EXACT: tx = ty = +24 saturates the sigmoid to 1.0f in both fp32 and fp16-storage arithmetic,
        so bx = (cx + 1) * stride bit for bit and the Kalman measurement's position carries no rounding noise; its width and
        height are odd integers, so the four corners sit on the half-pixel lattice k + 0.5 and the ReID crop, cut at int(x),
        int(x + w), int(y), int(y + h) (deep_sort/deep_sort.py:116-122), cannot depend on the last bits of the box;
      * the detections of a frame clear `conf_thres` on a ladder of objectness logits `gap` apart, the lowest m_thr above
        the threshold (both in units of the logits' standard deviation), so the score ORDER of the detections -- which
        decides the order of unmatched detections and with it the track ids (deep_sort/sort/tracker.py:160-161) -- does
        not flip under rounding noise;
      * candidates whose IoU with a detection exceeds iou_thres + m_iou stay `gap` below it, so NMS suppresses them
        whether they fire or not; every other candidate stays at least m_thr below the threshold.
    The class rows make class 0 win with an almost constant confidence (shape_heads), so score order == objectness order.
    Returns (ws, info).  Everything is computed from the fp32 oracle's features of these very frames."""
    heads = head_convs(blocks)
    nc, na = 80, 3
    obj = head_channels(nc, na)[0]
    nf = len(frames)
    H_img, W_img = frames[0].shape[:2]
    rng = np.random.default_rng(seed)
    geo = _head_geometry(blocks, ws, frames[0])
    feats, boxes = _features_and_boxes(blocks, ws, frames)
    noise = []
    for h, (li, _) in enumerate(heads):
        if not noise_metric:
            noise.append(np.ones(feats[h][0].shape[1]))
            continue
        # per-channel rounding-noise scale (fp16-storage oracle minus fp32 oracle): corrections minimum-norm in that metric
        acc = 0.0
        for fi, f in enumerate(frames):
            _, o16 = darknet_ref.forward(blocks, ws, frame_to_input(f), return_layers=True, half_storage=True)
            x16 = o16[li - 1][0]
            acc = acc + ((x16.reshape(x16.shape[0], -1).numpy().astype(np.float64).T - feats[h][fi]) ** 2).mean(0)
        sc = np.sqrt(acc / nf)
        noise.append(np.maximum(sc, 0.05 * np.median(sc)))
    ncell = [feats[h][0].shape[0] for h in range(len(heads))]
    off, o = {}, 0                        # first prediction row of every (head, anchor) block: rows are head-major, then anchor, then cell
    for h in range(len(heads)):
        for a in range(na):
            off[(h, a)] = o
            o += ncell[h]
    rows = o
    row_key = np.empty(rows, np.int64)
    for (h, a), o0 in off.items():
        row_key[o0:o0 + ncell[h]] = h * na + a
    w_row = {(h, a): ws[ci]["w"][obj[a], :, 0, 0].astype(np.float64) for h, (_, ci) in enumerate(heads) for a in range(na)}
    nat = np.stack([np.concatenate([feats[h][fi] @ w_row[(h, a)] for h in range(len(heads)) for a in range(na)])
                    for fi in range(nf)])                                                     # natural logits, bias 0
    sigma = float(nat.std())
    m_thr, gap = m_thr * sigma, gap * sigma
    z = np.empty_like(nat)                # per-row standardisation: every row contributes its own upper tail
    for key, o0 in off.items():
        blk = nat[:, o0:o0 + ncell[key[0]]]
        z[:, o0:o0 + ncell[key[0]]] = (blk - float(blk.mean())) / float(blk.std())

    # ---- 1. pick the detections of every frame: (row index, exact box) ----
    # A fitted row's rounding noise grows with the number of outputs it has to prescribe, so the nine rows share the
    # detections round-robin, each contributing its own upper tail on that frame.
    picks, site_boxes = [], []
    quota = [float(feats[h][0].shape[1]) ** (1.0 / 3.0) for h in range(len(heads))]    # n ~ C^(1/3) balances n^1.5 / sqrt(C)
    lim = max_box * min(H_img, W_img)
    for fi in range(nf):
        n_want = int(round(target * (1.0 + rng.uniform(-0.1, 0.1))))
        prio = np.empty(rows)
        pbox = np.zeros((rows, 4))
        for (h, a), o0 in off.items():
            gy, gx, anchors = geo[h]
            s_h, s_w = H_img / gy, W_img / gx                 # yolo_decode scales x by the HEIGHT stride
            zz = z[fi, o0:o0 + ncell[h]]
            rk = np.empty(ncell[h])
            rk[np.argsort(-zz, kind="stable")] = np.arange(ncell[h])
            prio[o0:o0 + ncell[h]] = (rk + 0.5) / quota[h] - 1e-3 * zz
            cy, cx = np.divmod(np.arange(ncell[h]), gx)
            nb_ = boxes[fi][o0:o0 + ncell[h]]
            bw = 2 * np.floor(0.5 * (nb_[:, 2] - nb_[:, 0])) + 1   # odd integer sizes next to the natural ones
            bh = 2 * np.floor(0.5 * (nb_[:, 3] - nb_[:, 1])) + 1
            bx, by = (cx + 1.0) * s_h, (cy + 1.0) * s_w          # the cell's far corner: sigmoid saturated at 1.0f
            pbox[o0:o0 + ncell[h]] = np.stack([bx - bw / 2, by - bh / 2, bx + bw / 2, by + bh / 2], 1)
        bw, bh = pbox[:, 2] - pbox[:, 0], pbox[:, 3] - pbox[:, 1]
        eligible = (bw <= lim) & (bh <= lim) & (bw >= 9) & (bh >= 9)
        picked = []
        for r in np.argsort(prio, kind="stable")[:40 * target]:
            if len(picked) >= n_want:
                break
            if not eligible[r]:
                continue
            if not picked or _box_iou(pbox[r][None], pbox[picked]).max() < iou_thres - m_iou:
                picked.append(int(r))
        picks.append(picked)
        site_boxes.append(pbox)

    # ---- 2. box rows of every firing site: tx = ty = +24 (saturated), tw / th = log(size / anchor); minimum-norm equalities ----
    for h, (li, ci) in enumerate(heads):
        gy, gx, anchors = geo[h]
        for a in range(na):
            Xs, tgt = [], []
            for fi in range(nf):
                for r in picks[fi]:
                    if row_key[r] != h * na + a:
                        continue
                    aw, ah = anchors[a]
                    b = site_boxes[fi][r]
                    tgt.append([24.0, 24.0, np.log((b[2] - b[0]) / aw), np.log((b[3] - b[1]) / ah)])
                    Xs.append(feats[h][fi][r - off[(h, a)]])
            if not Xs:
                continue
            X = np.asarray(Xs)
            A = np.concatenate([X / noise[h], np.ones((len(X), 1))], 1)          # unknowns: noise-scaled dw, db
            G = A @ A.T
            G += 1e-10 * np.trace(G) / len(G) * np.eye(len(G))
            tgt = np.asarray(tgt)
            for kk in range(4):
                row = a * (nc + 5) + kk
                w0 = ws[ci]["w"][row, :, 0, 0].astype(np.float64)
                b0 = float(ws[ci]["b"][row])
                d = A.T @ np.linalg.solve(G, tgt[:, kk] - (X @ w0 + b0))
                ws[ci]["w"][row, :, 0, 0] = (w0 + d[:-1] / noise[h]).astype(np.float32)
                ws[ci]["b"][row] = np.float32(b0 + d[-1])
    _, boxes = _features_and_boxes(blocks, ws, frames)                      # every candidate's box under the refitted rows
    orders = [np.argsort(-z[fi], kind="stable")[:40 * target] for fi in range(nf)]   # candidates worth an IoU test per frame
    for fi in range(nf):                                                    # (every firing site is among them)
        orders[fi] = np.unique(np.concatenate([orders[fi], np.asarray(picks[fi])]))

    # ---- 3. roles: head (value on the ladder), neighbour (any candidate whose IoU with a head exceeds iou_thres + m_iou: NMS
    #         suppresses it as long as it stays `gap` below that head, so it may fire or not -- only the band around the
    #         threshold is forbidden), everything else negative ----
    value = np.full((nf, rows), np.nan)                       # equalities (heads)
    upper = np.full((nf, rows), -m_thr)                       # upper bound of every other candidate
    lower = np.full((nf, rows), -np.inf)                      # lower bound (set for neighbours that end up firing)
    is_nb = np.zeros((nf, rows), bool)
    for fi in range(nf):
        picked = picks[fi]
        hb = boxes[fi][picked]
        iou_hh = _box_iou(hb, hb) - np.eye(len(picked))
        if iou_hh.max() >= iou_thres - 0.5 * m_iou:
            raise RuntimeError("two heads overlap too much after the box refit")
        n = len(picked)
        for k, r in enumerate(picked):
            value[fi, r] = m_thr + (n - 1 - k) * gap
        cand = orders[fi]
        iou = _box_iou(boxes[fi][cand], hb)
        j = np.argmax(iou, 1)
        top = value[fi, np.asarray(picked)[j]] - gap
        ok = (iou[np.arange(len(cand)), j] > iou_thres + m_iou) & (top >= m_thr) & np.isnan(value[fi, cand])
        upper[fi, cand[ok]] = top[ok]
        is_nb[fi, cand[ok]] = True

    # ---- 4. objectness rows: active-set QP (equalities for the heads, bounds for everything else) ----
    logit_thr = float(np.log(conf_thres / (1 - conf_thres)))
    n_firing_nb = 0
    for (h, a), w0 in w_row.items():
        X = np.concatenate([feats[h][fi] for fi in range(nf)], 0)
        n = ncell[h]
        sl = lambda arr: np.concatenate([arr[fi, off[(h, a)]:off[(h, a)] + n] for fi in range(nf)])
        tgt, hi, lo, nb = sl(value), sl(upper), sl(lower), sl(is_nb)
        is_eq = ~np.isnan(tgt)
        Xa = np.concatenate([X / noise[h], np.ones((len(X), 1))], 1)            # unknowns: noise-scaled dw, db
        b0 = -float(np.quantile(X @ w0, 1.0 - 3.0 * target / (9.0 * n)))      # the row's own firing level sits at the threshold
        base = X @ w0 + b0
        for outer in range(12):
            side = np.zeros(len(X), np.int8)                  # 0 inactive, +1 at the upper bound, -1 at the lower bound
            w, b = w0, b0
            for it in range(3000):
                act = is_eq | (side != 0)
                ia = np.nonzero(act)[0]
                if len(ia):
                    A = Xa[ia]
                    G = A @ A.T
                    rhs = np.where(is_eq[ia], np.nan_to_num(tgt[ia]), np.where(side[ia] > 0, hi[ia], lo[ia])) - base[ia]
                    sol = np.linalg.solve(G + 1e-10 * np.trace(G) / len(G) * np.eye(len(G)), rhs)
                    wrong = (~is_eq[ia]) & (sol * side[ia] > 1e-12)   # a bound whose multiplier says "inactive"
                    if wrong.any():
                        side[ia[wrong]] = 0
                        continue
                    d = A.T @ sol
                    w, b = w0 + d[:-1] / noise[h], b0 + d[-1]
                lg = X @ w + b
                over = (~act) & (lg > hi + 1e-7)
                under = (~act) & (lg < lo - 1e-7)
                if not over.any() and not under.any():
                    break
                idx = np.nonzero(over | under)[0]
                if len(idx) > 32:
                    viol = np.where(over, lg - hi, lo - lg)
                    idx = idx[np.argsort(-viol[idx])[:32]]
                side[idx] = np.where(over[idx], 1, -1)
            else:
                raise RuntimeError(f"head row ({h},{a}) refit did not converge ({int(act.sum())} active constraints)")
            # neighbours inside the forbidden band around the threshold pick the side they are on, then the row is solved again
            band = nb & (lg > -m_thr + 1e-6) & (lg < m_thr - 1e-6) & (lo == -np.inf)
            fire = nb & (lg >= m_thr - 1e-6) & (lo == -np.inf)
            if not band.any():
                lo[fire] = m_thr                               # (already satisfied: recorded for the count only)
                break
            up = band & (lg >= 0)
            lo[up] = m_thr
            hi[band & ~up] = -m_thr
        else:
            raise RuntimeError(f"head row ({h},{a}): neighbours keep landing in the threshold band")
        n_firing_nb += int((nb & (lg >= m_thr - 1e-6)).sum())
        ci = heads[h][1]
        ws[ci]["w"][obj[a], :, 0, 0] = w.astype(np.float32)
        ws[ci]["b"][obj[a]] = np.float32(b + logit_thr)
        if verbose:
            print("  head %d anchor %d: %d heads, %d neighbours, %d active constraints of %d, |dw|/|w| = %.3f" %
                  (h, a, int(is_eq.sum()), int(nb.sum()), int((is_eq | (side != 0)).sum()), len(X), np.linalg.norm(w - w0) / np.linalg.norm(w0)))
    n_riders = n_firing_nb
    # ---- 5. verify on the oracle: exactly the heads survive, in ladder order, corners on the half-pixel lattice ----
    for fi, f in enumerate(frames):
        det = darknet_ref.detect(blocks, ws, f, f.shape[:2], conf_thres, iou_thres)
        want = boxes[fi][picks[fi]]
        if det is None or len(det) != len(want) or np.abs(det[:, :4] - want).max() > 1e-2:
            raise RuntimeError(f"frame {fi}: the calibrated detector does not return exactly the planned detections")
        frac = det[:, :4] - np.floor(det[:, :4])
        if np.abs(frac - 0.5).max() > 0.05:
            raise RuntimeError(f"frame {fi}: a box corner is off the half-pixel lattice")
    info = dict(sigma=sigma, m_thr=float(m_thr), gap=float(gap), n_heads=[len(p) for p in picks], n_riders=n_riders)
    return ws, info


def calibrate_heads(blocks, ws, frames, want_dets=50, conf_thres=0.5, iou_thres=0.4, verbose=False):
    """fit_head_margins + the per-frame detection counts (exactly the planned heads survive NMS)."""
    ws, info = fit_head_margins(blocks, ws, frames, target=want_dets, conf_thres=conf_thres, iou_thres=iou_thres, verbose=verbose)
    info["n_dets"] = list(info["n_heads"])
    return ws, info


def darknet_weights(blocks, frames, seed=0, target=50, conf_thres=0.5):
    """Seeded weights whose BN statistics are calibrated on frames[0] and whose head objectness rows carry decision margins
    on every frame in `frames` (fit_head_margins): about `target` detections per frame survive `conf_thres` + NMS as class 0
    only, with threshold, score-order and NMS margins that reduced-precision arithmetic cannot cross.  Returns (ws, info)."""
    ws = darknet_ref.init_weights(blocks, seed)
    darknet_ref.forward(blocks, ws, frame_to_input(frames[0]), calibrate_bn=True)
    shape_heads(ws)
    return calibrate_heads(blocks, ws, frames, want_dets=target, conf_thres=conf_thres)


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
