"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (imported from /root/reference
through oracle/ref_shims.py), and check that the oracle restatement reproduces it bit-for-bit on CPU.

Build-container only:   python -m oracle.gen_golden
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Golden files (small, committed):
  kalman_demo.npz     the reference's only known-answer snippet (deep_sort/sort/kalman_filter.py:259-273)
  kalman_batch.npz    seeded KalmanFilter.initiate/predict/update/gating_distance in/out
  assoc_seq.npz       24-frame DeepSort.update sequence with an injected extractor: per-frame inputs
                      (boxes, features, class ids) and the reference's (K,6) int32 outputs + track tables
  assoc_seq2.npz      44 frames with nn_budget=4, max_age=3, n_init=2: gallery FIFO truncation, deletion by age, re-identification
  tiny416.npz         yolov3-tiny 416: head outputs digest, soft_non_max_suppression output, one frame
  yolov4_tiny_416.npz, yolov4_416.npz   the same for the grouped-route and the Mish / SPP / shortcut architectures
  video_detector.npz  the reference's own VideoDetector loop (skip_frames=2, tracker, overlay) on a lossless clip: held rows and
                      image digests per frame
  reid.npz            Extractor features for boxes on one frame (crop + cv2.resize + Net)
  overlay.npz         LabelDrawer.draw_labels_by_trackers / draw_labels output (images and digests) on two synthetic frames
  action.npz          ActionIdentify.update (action/) on a seeded (K,6) row sequence with patched time stamps: the emitted
                      (track id, class id, rule) triples per frame, every rule firing, deletions and re-insertions included
  window.npz          ImageDetector.detect in sliding-window mode (win_size, overlap; batched tiles + merge-NMS) on a 700x1000
                      image, and soft_non_max_suppression(merge=True) on hand-built predictions that reach the k == n and
                      k == 1 branches of the merge block
"""
import hashlib
import os
import sys
import tempfile

import numpy as np
import torch

from . import ref_shims

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
CFG_DIR = os.path.join(os.path.dirname(HERE), "config")


def _eq(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b):
        d = np.abs(a.astype(np.float64) - b.astype(np.float64)).max() if a.shape == b.shape else "shape"
        raise AssertionError(f"oracle != reference for {what}: max diff {d}")
    print(f"  oracle == reference (bit-exact): {what}")


def _close(a, b, tol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.abs(a - b).max()
    if a.shape != b.shape or d > tol:
        raise AssertionError(f"oracle != reference for {what}: max diff {d} > {tol}")
    print(f"  oracle ~= reference (max abs diff {d:.3g} <= {tol}): {what}")


def gen_kalman():
    from deep_sort.sort.kalman_filter import KalmanFilter
    from . import sort_ref as S
    kf = KalmanFilter()
    m, c = kf.initiate(torch.tensor([10, 15, 0.5, 10]))
    diag0 = torch.diagonal(c[0]).numpy().copy()
    m, c = kf.predict(m, c)
    diag1 = torch.diagonal(c[0]).numpy().copy()
    m, c = kf.update(m, c, torch.tensor([12, 20, 0.6, 11]))
    m2, c2 = kf.initiate(torch.tensor([12, 13, 0.7, 5]))
    m2, c2 = kf.predict(m2, c2)
    m2, c2 = kf.update(m2, c2, torch.tensor([13, 14, 0.7, 8]))
    meas = torch.tensor([[12, 20, 0.6, 11], [20, 16, 0.4, 18]])
    maha = kf.gating_distance(torch.cat((m, m2), 0), torch.cat((c, c2), 0), meas)
    np.savez(os.path.join(GOLD, "kalman_demo.npz"), diag_init=diag0, diag_pred=diag1, mean_upd=m.numpy(),
             cov_upd=c.numpy(), maha4=maha.numpy())
    # oracle on the same snippet
    om, oc = S.kf_initiate(torch.tensor([10, 15, 0.5, 10]))
    _eq(torch.diagonal(oc[0]), diag0, "kf.initiate")
    om, oc = S.kf_predict(om, oc)
    _eq(torch.diagonal(oc[0]), diag1, "kf.predict")
    om, oc = S.kf_update(om, oc, torch.tensor([12, 20, 0.6, 11]))
    _eq(om, m, "kf.update mean"); _eq(oc, c, "kf.update cov")

    rng = np.random.default_rng(7)
    n, mm = 64, 48
    xyah = np.stack([rng.uniform(0, 600, n), rng.uniform(0, 600, n), rng.uniform(0.3, 0.7, n), rng.uniform(40, 160, n)], 1).astype(np.float32)
    means, covs = zip(*[kf.initiate(torch.from_numpy(x)) for x in xyah])
    mean0, cov0 = torch.cat(means, 0), torch.cat(covs, 0)
    mean1, cov1 = kf.predict(mean0, cov0)
    z = (xyah + rng.normal(0, [2, 2, 0.01, 2], (n, 4))).astype(np.float32)
    mean2, cov2 = kf.update(mean1, cov1, torch.from_numpy(z))
    mean3, cov3 = kf.predict(mean2, cov2)
    dets = np.concatenate([z[:mm - 8] + rng.normal(0, [4, 4, 0.01, 3], (mm - 8, 4)),
                           np.stack([rng.uniform(0, 600, 8), rng.uniform(0, 600, 8), rng.uniform(0.3, 0.7, 8), rng.uniform(40, 160, 8)], 1)], 0).astype(np.float32)
    gate2 = kf.gating_distance(mean3, cov3, torch.from_numpy(dets), only_position=True)
    np.savez(os.path.join(GOLD, "kalman_batch.npz"), xyah=xyah, mean0=mean0.numpy(), cov0=cov0.numpy(),
             mean1=mean1.numpy(), cov1=cov1.numpy(), z=z, mean2=mean2.numpy(), cov2=cov2.numpy(),
             mean3=mean3.numpy(), cov3=cov3.numpy(), dets_xyah=dets, gate2=gate2.numpy())
    o1 = S.kf_predict(mean0, cov0); _eq(o1[0], mean1, "batch predict mean"); _eq(o1[1], cov1, "batch predict cov")
    o2 = S.kf_update(mean1, cov1, torch.from_numpy(z)); _eq(o2[0], mean2, "batch update mean"); _eq(o2[1], cov2, "batch update cov")
    _eq(S.kf_gating_position(mean3, cov3, torch.from_numpy(dets)), gate2, "gating (position only)")


def _assoc_sequence(name, sc, n_frames, kw):
    from deep_sort import DeepSort
    from . import sort_ref as S
    frames = [sc.step() for _ in range(n_frames)]
    feats_now = {}

    class Inject:                                    # extractor injection point, deep_sort/deep_sort.py:28-31
        def __call__(self, crops):
            return torch.from_numpy(feats_now["f"])

    frame_img = np.zeros((608, 608, 3), np.uint8)
    ref = DeepSort(Inject(), min_confidence=1, use_cuda=False, **kw)
    orc = S.DeepSortRef(lambda fr, tl: torch.from_numpy(feats_now["f"]), **kw)
    out = {}
    stats = dict(max_tracks=0, deleted=0, rows=0, budget_hit=0)
    seen = set()
    for t, (tl, ft, cl) in enumerate(frames):
        ft = ft.astype(np.float16).astype(np.float32)      # stored as fp16 in the fixture: keep it lossless
        feats_now["f"] = ft
        r = ref.update(torch.from_numpy(tl.copy()), torch.ones(len(tl)), frame_img, torch.from_numpy(cl))
        o = orc.update(tl.copy(), None, frame_img, torch.from_numpy(cl))
        r = np.asarray(r, np.int32).reshape(-1, 6)
        o = np.asarray(o, np.int32).reshape(-1, 6)
        _eq(o, r, f"DeepSort.update frame {t} ({len(tl)} dets -> {len(r)} rows)")
        trk = ref.tracker.tracks
        tab = np.array([[k.track_id, k.hits, k.age, k.time_since_update, k.state] for k in trk], np.int32).reshape(-1, 5)
        st = orc.tracker.state_arrays()
        _eq(np.stack([st["ids"], st["hits"], st["age"], st["tsu"], st["state"]], 1).reshape(-1, 5), tab, f"track table frame {t}")
        if len(trk):
            _eq(st["mean"], torch.cat([k.mean for k in trk], 0).numpy(), f"track means frame {t}")
            _eq(st["cov"], torch.cat([k.covariance for k in trk], 0).numpy(), f"track covs frame {t}")
        ids = set(int(k.track_id) for k in trk)
        stats["deleted"] += len(seen - ids)
        seen = ids
        stats["max_tracks"] = max(stats["max_tracks"], len(trk)); stats["rows"] += len(r)
        stats["budget_hit"] += sum(1 for v in ref.tracker.metric.samples.values() if len(v) >= kw["nn_budget"])
        out[f"tlwh_{t}"], out[f"feat_{t}"], out[f"cls_{t}"] = tl, ft.astype(np.float16), cl
        out[f"out_{t}"], out[f"table_{t}"] = r, tab
        out[f"mean_{t}"] = st["mean"]
    out["n_frames"] = np.int32(len(frames))
    out["params"] = np.array([kw["max_dist"], kw["max_iou_distance"], kw["max_age"], kw["n_init"], kw["nn_budget"]], np.float64)
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print(f"  {name}: {stats}")
    return stats


def gen_assoc():
    from .synth import Scenario
    _assoc_sequence("assoc_seq.npz", Scenario(n=40, seed=3, p_miss=0.08, p_new=0.03, p_leave=0.02), 24,
                    dict(max_dist=0.3, max_iou_distance=0.7, max_age=30, n_init=3, nn_budget=30))
    # second sequence: a tiny budget and a short life, so that the FIFO truncation of the galleries (nn_matching.py:153-154), the
    # deletion by age (track.py:146-152) and re-identification after misses are all inside what the unmodified reference pins
    st = _assoc_sequence("assoc_seq2.npz", Scenario(n=24, seed=9, p_miss=0.25, p_new=0.06, p_leave=0.04), 44,
                         dict(max_dist=0.3, max_iou_distance=0.7, max_age=3, n_init=2, nn_budget=4))
    assert st["deleted"] > 10 and st["budget_hit"] > 50, st


def gen_tiny():
    from yolo3.models import Darknet
    from yolo3.utils.model_build import soft_non_max_suppression
    from . import darknet_ref as D
    from .synth import make_frame, darknet_weights, frame_to_input
    cfg = os.path.join(CFG_DIR, "yolov3-tiny.cfg")
    blocks = D.parse_cfg(cfg)
    frames = [make_frame(416, 416, seed=s) for s in (0, 1)]
    ws, info = darknet_weights(blocks, frames, seed=0, target=50)
    print("  tiny416 calibration:", info)
    with tempfile.TemporaryDirectory() as td:
        wpath = os.path.join(td, "tiny.weights")
        D.write_weights(wpath, blocks, ws)
        model = Darknet(cfg, img_size=(416, 416))
        model.load_darknet_weights(wpath)
        model.eval()
        _, ws2 = D.read_weights(wpath, blocks)
    for a, b in zip(ws, ws2):
        _eq(a["w"], b["w"], "weights roundtrip") if a is ws[0] else None
    x = frame_to_input(frames[0])
    with torch.no_grad():
        pred = model(x)
        dets = soft_non_max_suppression(pred.clone(), 0.5, 0.4)[0]
    op = D.forward(blocks, ws2, x)
    _eq(op, pred, "Darknet.forward yolov3-tiny 416")
    od = D.postprocess(op[0].numpy(), 0.5, 0.4)
    _eq(od, dets.numpy(), f"soft_non_max_suppression ({len(od)} detections)")
    sel = np.argsort(-pred[0, :, 4].numpy(), kind="stable")[:256]
    np.savez_compressed(os.path.join(GOLD, "tiny416.npz"), frame_seed=np.int32(0), weight_seed=np.int32(0),
                        pred_top_idx=sel.astype(np.int32), pred_top=pred[0, sel].numpy(), dets=dets.numpy(),
                        pred_sha256=np.frombuffer(hashlib.sha256(pred.numpy().tobytes()).digest(), np.uint8))


def gen_other_cfgs():
    """yolov4-tiny 416 (grouped routes) and yolov4 416 (Mish, SPP 5/9/13 max-pools, shortcuts, PAN routes): forward digest and
    soft_non_max_suppression output of the unmodified reference, on seeded weights with calibrated heads."""
    from yolo3.models import Darknet
    from yolo3.utils.model_build import soft_non_max_suppression
    from . import darknet_ref as D
    from .synth import make_frame, darknet_weights, frame_to_input
    for name in ("yolov4-tiny", "yolov4"):
        cfg = os.path.join(CFG_DIR, name + ".cfg")
        blocks = D.parse_cfg(cfg)
        frames = [make_frame(416, 416, seed=s) for s in (6, 7)]
        ws, info = darknet_weights(blocks, frames, seed=2, target=40)
        with tempfile.TemporaryDirectory() as td:
            wpath = os.path.join(td, "w.weights")
            D.write_weights(wpath, blocks, ws)
            model = Darknet(cfg, img_size=(416, 416))
            model.load_darknet_weights(wpath)
            model.eval()
        x = frame_to_input(frames[0])
        with torch.no_grad():
            pred = model(x)
            dets = soft_non_max_suppression(pred.clone(), 0.5, 0.4)[0]
        op = D.forward(blocks, ws, x)
        _eq(op, pred, f"Darknet.forward {name} 416")
        od = D.postprocess(op[0].numpy(), 0.5, 0.4)
        _eq(od, dets.numpy(), f"soft_non_max_suppression {name} ({len(od)} detections)")
        sel = np.argsort(-pred[0, :, 4].numpy(), kind="stable")[:256]
        np.savez_compressed(os.path.join(GOLD, name.replace("-", "_") + "_416.npz"), frame_seeds=np.asarray([6, 7], np.int32),
                            weight_seed=np.int32(2), pred_top_idx=sel.astype(np.int32), pred_top=pred[0, sel].numpy(), dets=dets.numpy(),
                            pred_sha256=np.frombuffer(hashlib.sha256(pred.numpy().tobytes()).digest(), np.uint8))


def video_fixture(td):
    """Files of the VideoDetector fixture: a 9-frame FFV1 clip of one held 416x416 scene, yolov3-tiny weights calibrated on it, a ReID
    checkpoint and a names file.  Returns (cfg, blocks, ws, sd, paths, clip)."""
    import cv2
    from . import darknet_ref as D
    from .synth import darknet_weights, make_frame, reid_state_dict
    cfg = os.path.join(CFG_DIR, "yolov3-tiny.cfg")
    blocks = D.parse_cfg(cfg)
    scenes = [make_frame(416, 416, seed=s) for s in (0, 1)]
    ws, _ = darknet_weights(blocks, scenes, seed=0, target=50)
    sd = reid_state_dict(seed=0)
    paths = {k: os.path.join(td, v) for k, v in (("weights", "tiny.weights"), ("ckpt", "ckpt.t7"), ("names", "coco.names"), ("video", "clip.avi"))}
    D.write_weights(paths["weights"], blocks, ws)
    torch.save({"net_dict": sd, "acc": 0.0, "epoch": 0}, paths["ckpt"])
    with open(paths["names"], "w") as fh:
        fh.write("\n".join(f"c{i}" for i in range(80)) + "\n")
    clip = [scenes[0]] * 9
    wr = cv2.VideoWriter(paths["video"], cv2.VideoWriter_fourcc(*"FFV1"), 25, (416, 416))
    assert wr.isOpened()
    for f in clip:
        wr.write(cv2.cvtColor(f, cv2.COLOR_RGB2BGR))
    wr.release()
    return cfg, blocks, ws, sd, paths, clip


def gen_video():
    """The reference's own VideoDetector loop (yolo3/detect/video_detect.py:78-208) with video_deepsort.py's keyword arguments
    (skip_frames=2 among them) on a lossless clip: per frame the held rows and the digest of the yielded image (overlay drawn)."""
    from deep_sort import DeepSort
    from yolo3.detect.video_detect import VideoDetector
    from yolo3.models import Darknet
    from . import darknet_ref as D, reid_ref as R, sort_ref as S
    with tempfile.TemporaryDirectory() as td:
        cfg, blocks, ws, sd, paths, clip = video_fixture(td)
        model = Darknet(cfg, img_size=(416, 416))
        model.load_darknet_weights(paths["weights"])
        tracker = DeepSort(paths["ckpt"], min_confidence=1, use_cuda=False, nn_budget=30, n_init=3, max_iou_distance=0.7, max_dist=0.3, max_age=30)
        vd = VideoDetector(model, paths["names"], thickness=2, skip_frames=2, thres=0.5, class_mask=[0, 2, 4], nms_thres=0.4, tracker=tracker,
                           half=False)
        out = {}
        orc = S.DeepSortRef(lambda fr, tl: R.extract(sd, fr, tl), max_dist=0.3, max_iou_distance=0.7, max_age=30, n_init=3, nn_budget=30)
        held = None
        n = 0
        for t, (image, rows, actions) in enumerate(vd.detect(paths["video"], real_show=False, skip_secs=0, show_fps=False)):
            rows = np.asarray(rows, np.int32).reshape(-1, 6)
            if t % 2 == 0:                                # the oracle steps on the frames the reference loop detects on
                det = D.detect(blocks, ws, clip[t], (416, 416), 0.5, 0.4)
                tlwh, conf, cls = D.to_tracker_inputs(det, [0, 2, 4])
                held = np.asarray(orc.update(tlwh, conf, clip[t], torch.from_numpy(cls)), np.int32).reshape(-1, 6)
            _eq(held, rows, f"VideoDetector frame {t} ({len(rows)} rows held)")
            out[f"rows_{t}"] = rows
            out[f"image_sha256_{t}"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(image).tobytes()).digest(), np.uint8)
            n += 1
        assert n == len(clip)
        out["n_frames"] = np.int32(n)
        np.savez_compressed(os.path.join(GOLD, "video_detector.npz"), **out)


def gen_reid():
    from deep_sort.deep.feature_extractor import Extractor
    from . import reid_ref as R
    from .synth import make_frame, reid_state_dict
    from .cv_resize_ref import crops_to_batch
    sd = reid_state_dict(seed=0)
    frame = make_frame(608, 608, seed=5)
    rng = np.random.default_rng(11)
    m = 12
    tlwh = np.stack([rng.uniform(-5, 540, m), rng.uniform(-5, 470, m), rng.uniform(20, 90, m), rng.uniform(40, 160, m)], 1).astype(np.float32)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "ckpt.t7")
        torch.save({"net_dict": sd}, p)
        ex = Extractor(p, use_cuda=False)
    H, W = frame.shape[:2]
    crops = []
    for b in torch.from_numpy(tlwh):                  # deep_sort/deep_sort.py:116-122,138-141
        x, y, w, h = b
        x1, x2, y1, y2 = max(int(x), 0), min(int(x + w), W - 1), max(int(y), 0), min(int(y + h), H - 1)
        crops.append(frame[y1:y2, x1:x2])
    feats = ex(crops)
    of = R.extract(sd, frame, tlwh)
    # crop + fixed-point resize + normalisation are bit-exact; the 20-conv fp32 stack is compared with a
    # tolerance because oneDNN's fp32 conv result depends on buffer placement (observed 2e-7).
    _eq(crops_to_batch(frame, tlwh), ex._preprocess(crops).numpy(), "Extractor._preprocess (crop + cv2.resize + normalise)")
    _close(of, feats, 1e-6, "Extractor features (Net forward)")
    np.savez_compressed(os.path.join(GOLD, "reid.npz"), frame_seed=np.int32(5), weight_seed=np.int32(0), tlwh=tlwh,
                        feats=feats.numpy())


def window_image_and_weights():
    """The sliding-window fixture: a 700x1000 synthetic image and yolov3-tiny weights whose heads are calibrated on its six
    416x416(+15 %) windows (so that every window yields ~12 detections)."""
    import cv2
    from . import darknet_ref as D
    from .synth import darknet_ref as _d, frame_to_input, make_frame, shape_heads, calibrate_heads
    cfg = os.path.join(CFG_DIR, "yolov3-tiny.cfg")
    blocks = D.parse_cfg(cfg)
    img = make_frame(700, 1000, seed=21, n_rect=90)
    win, ov = (416, 416), 0.15
    tiles = []
    for x in range(0, 1000, win[0]):
        for y in range(0, 700, win[1]):
            sub = img[y:y + win[1] + int(win[1] * ov), x:x + win[0] + int(win[0] * ov)]
            tiles.append(cv2.resize(sub, (416, 416), interpolation=cv2.INTER_LINEAR))
    ws = D.init_weights(blocks, 5)
    D.forward(blocks, ws, torch.cat([frame_to_input(t) for t in tiles], 0), calibrate_bn=True)
    shape_heads(ws)
    ws, info = calibrate_heads(blocks, ws, tiles, want_dets=12)
    return cfg, blocks, ws, img, win, ov, info


def merge_cases():
    """Hand-built corner-box predictions for soft_non_max_suppression(merge=True, is_p1p2=True): (a) nothing suppressed (k == n:
    every kept row becomes the one weighted mean box), (b) one cluster (k == 1: the intended merge), (c) two clusters (1 < k < n:
    the merge line raises inside the reference's try/except and nothing is merged)."""
    def mk(boxes, scores):
        p = np.zeros((len(boxes), 85), np.float32)
        p[:, :4] = np.asarray(boxes, np.float32)
        p[:, 4] = np.asarray(scores, np.float32)
        p[:, 5] = 0.95
        return p
    a = mk([[50 + 160 * i, 100, 150 + 160 * i, 220] for i in range(4)], [0.9, 0.7, 0.8, 0.6])
    b = mk([[50 + 3 * i, 100 + 2 * i, 150 + 3 * i, 220 + 2 * i] for i in range(5)], [0.7, 0.9, 0.8, 0.65, 0.75])
    c = mk([[50 + 3 * i, 100, 150 + 3 * i, 220] for i in range(3)] + [[400 + 3 * i, 300, 520 + 3 * i, 420] for i in range(3)],
           [0.7, 0.9, 0.8, 0.65, 0.75, 0.85])
    return {"all_kept": a, "one_cluster": b, "two_clusters": c}


def gen_window():
    import contextlib
    import io
    from yolo3.detect.img_detect import ImageDetector
    from yolo3.models import Darknet
    from yolo3.utils.model_build import soft_non_max_suppression
    from . import darknet_ref as D
    cfg, blocks, ws, img, win, ov, info = window_image_and_weights()
    print("  window calibration:", info)
    with tempfile.TemporaryDirectory() as td:
        wpath, names = os.path.join(td, "tiny.weights"), os.path.join(td, "coco.names")
        D.write_weights(wpath, blocks, ws)
        with open(names, "w") as fh:
            fh.write("\n".join(f"c{i}" for i in range(80)) + "\n")
        model = Darknet(cfg, img_size=(416, 416))
        model.load_darknet_weights(wpath)
        det = ImageDetector(model, names, thres=0.5, nms_thres=0.4, win_size=win, overlap=ov, half=False)
        with contextlib.redirect_stdout(io.StringIO()):           # the merge block prints its tensors when it raises
            ref = det.detect(img)
    od, _ = D.detect_windows(blocks, ws, img, (416, 416), win, ov, 0.5, 0.4)
    _eq(od, ref.numpy(), f"ImageDetector.detect, sliding-window mode ({len(od)} detections)")
    out = {"dets": ref.numpy(), "win": np.asarray(win, np.int32), "overlap": np.float64(ov)}
    for name, p in merge_cases().items():
        with contextlib.redirect_stdout(io.StringIO()):
            r = soft_non_max_suppression(torch.from_numpy(p)[None].clone(), 0.5, 0.4, merge=True, is_p1p2=True)[0]
        o = D.postprocess(p, 0.5, 0.4, merge=True, is_p1p2=True)
        _eq(o, r.numpy(), f"soft_non_max_suppression(merge=True) case {name} ({len(o)} rows)")
        out["merge_" + name] = r.numpy()
    np.savez_compressed(os.path.join(GOLD, "window.npz"), **out)


def overlay_inputs():
    """Frames and rows for the overlay fixture: tracker rows [x1,y1,x2,y2,id,cls] (int32, DeepSort.update's output) and detector
    rows [x1,y1,x2,y2,conf,cls] (float32), on a 240x320 and a 608x608 synthetic frame."""
    from .synth import make_frame
    rng = np.random.default_rng(33)
    out = {}
    for name, (h, w) in (("small", (240, 320)), ("full", (608, 608))):
        n = 14
        x1 = rng.integers(-5, w - 40, n); y1 = rng.integers(-5, h - 60, n)
        rows = np.stack([x1, y1, x1 + rng.integers(20, 90, n), y1 + rng.integers(30, 120, n), rng.integers(1, 400, n),
                         rng.choice([0, 2, 4, 7], n)], 1).astype(np.int32)
        dets = np.concatenate([rows[:, :4].astype(np.float32) + rng.uniform(0, 1, (n, 4)).astype(np.float32),
                               rng.uniform(0.5, 1.0, (n, 1)).astype(np.float32), rows[:, 5:6].astype(np.float32)], 1)
        out[name] = (make_frame(h, w, seed=40 + len(out)), rows, dets)
    return out


def gen_overlay():
    from yolo3.utils.label_draw import LabelDrawer
    classes = [f"c{i}" for i in range(80)]
    res = {}
    for name, (frame, rows, dets) in overlay_inputs().items():
        ld = LabelDrawer(classes, None, 10, 2, img_size=frame.shape[:2])
        a, _, _ = ld.draw_labels_by_trackers(frame.copy(), rows, only_rect=False)
        b, _, _ = ld.draw_labels(frame.copy(), torch.from_numpy(dets), only_rect=False)
        c, _, _ = ld.draw_labels_by_trackers(frame.copy(), rows, only_rect=True)
        if name == "small":
            res["small_tracks"], res["small_dets"], res["small_rects"] = a, b, c
        for k, im in (("tracks", a), ("dets", b), ("rects", c)):
            res[f"{name}_{k}_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(im).tobytes()).digest(), np.uint8)
        res[name + "_colors"] = np.asarray(ld.colors, np.int32)
    np.savez_compressed(os.path.join(GOLD, "overlay.npz"), **res)
    print("  overlay fixture written (reference LabelDrawer on 2 frames x 3 modes)")


from .action_ref import ACTION_RULES, action_sequence  # noqa: E402  (shared with the tests: no reference needed to rebuild the inputs)


def gen_action():
    """action.npz: the reference ActionIdentify.update (action/action_Identify.py:15-47) on action_sequence()."""
    import time as _time
    from action import actions as RA
    from action.action_Identify import ActionIdentify as RefAI
    frames, stamps = action_sequence()
    rules = [getattr(RA, name)(cid, prm) for name, cid, prm in ACTION_RULES]
    ai = RefAI(rules, max_age=6, max_size=4)
    real = _time.time
    out = {"n_frames": np.int32(len(frames)), "stamps": np.asarray(stamps, np.float64)}
    total = 0
    try:
        for f, (rows, ts) in enumerate(zip(frames, stamps)):
            _time.time = lambda ts=ts: ts
            res = ai.update(rows if len(rows) else [])
            names = [r.name for r in rules]
            trip = np.asarray([(tid, cid, [i for i, r in enumerate(rules) if r.name == nm and r.class_id == cid][0]) for tid, cid, nm in res],
                              np.int32).reshape(-1, 3)
            out[f"rows_{f}"] = rows
            out[f"actions_{f}"] = trip
            total += len(trip)
    finally:
        _time.time = real
    assert total > 40, total
    kinds = sorted({int(t[2]) for f in range(len(frames)) for t in out[f"actions_{f}"]})
    assert kinds == list(range(len(ACTION_RULES))), kinds       # every rule fires somewhere
    np.savez_compressed(os.path.join(GOLD, "action.npz"), **out)
    print(f"  action fixture written ({len(frames)} frames, {total} (track, class, rule) triples, rules seen {kinds})")


def main():
    ref_shims.install()
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    for fn in (gen_kalman, gen_assoc, gen_tiny, gen_other_cfgs, gen_reid, gen_window, gen_overlay, gen_video, gen_action):
        print(fn.__name__)
        fn()
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    sys.exit(main())
