"""Oracle run of a held-scene clip (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py): the per-frame flow of
VideoDetector.detect (yolo3/detect/video_detect.py:134-149) with the detector and ReID results memoised per distinct scene
(the clip holds each scene for several frames), the tracker running frame by frame.  Used by the margin tests, the GPU
end-to-end parity tests, bench.py's CPU legs and oracle/gen_bench_calib.py (schedule search)."""
import numpy as np
import torch


class ClipOracle:
    def __init__(self, blocks, ws, reid_sd, scenes, thres, nms_thres, class_mask, tracker_kw, half_detector=False, noise=None):
        """half_detector=True restates the reference's half=True mode (yolo3/detect/img_detect.py:48-50,79-82: fp16 weights and
        activations) with the rounding points of the fused CUDA epilogue; False is the fp32 CPU path.  The ReID net is fp32 in
        the reference on every device (deep_sort/deep/feature_extractor.py never halves).
        noise=(seed, box_px, feat_rel): margin probe -- every frame the width and height of the detector boxes handed to the tracker are jittered by
        U(-box_px, box_px) about their centres (the ReID crops stay those of the unjittered boxes: the calibrated detector puts every corner on
        the half-pixel lattice, so a sub-half-pixel error never changes a crop) and the features by relative Gaussian noise
        of norm feat_rel; a clip whose ids survive that is insensitive to rounding noise of that size."""
        from . import sort_ref as S
        self.blocks, self.ws, self.sd, self.scenes = blocks, ws, reid_sd, scenes
        self.thres, self.nms_thres, self.mask, self.half = thres, nms_thres, class_mask, half_detector
        self._det, self._feat = {}, {}
        kw = {k: v for k, v in tracker_kw.items() if k != "min_confidence"}
        self.tracker = S.DeepSortRef(self._features, **kw)
        self._cur = None
        self.noise = noise
        self._rng = np.random.default_rng(noise[0]) if noise else None

    def detect(self, si):
        """(det (n,6) | None, tlwh, conf, cls) of scene `si`."""
        if si not in self._det:
            from . import darknet_ref as D
            f = self.scenes[si]
            x = torch.from_numpy(np.ascontiguousarray(f)).permute(2, 0, 1) / 255.
            pred = D.forward(self.blocks, self.ws, x.unsqueeze(0), half_storage=self.half)
            det = D.postprocess(pred[0].numpy(), self.thres, self.nms_thres)
            self._det[si] = (det,) + (D.to_tracker_inputs(det, self.mask) if det is not None else (None, None, None))
        return self._det[si]

    def features(self, si):
        if si not in self._feat:
            from . import reid_ref as R
            self._feat[si] = R.extract(self.sd, self.scenes[si], self.detect(si)[1])
        return self._feat[si]

    def _features(self, frame, tlwh):
        f = self.features(self._cur)
        if self.noise:
            n = torch.from_numpy(self._rng.standard_normal(tuple(f.shape)).astype(np.float32))
            f = f + self.noise[2] * f.norm(dim=1, keepdim=True) * n / n.norm(dim=1, keepdim=True)
        return f

    def step(self, si):
        """One frame showing scene `si`: returns (rows int32 (K,6) | None when nothing was detected, det)."""
        det, tlwh, conf, cls = self.detect(si)
        if det is None:
            return None, None
        self._cur = si
        if self.noise:
            # the calibrated detector saturates the box centres (exact in every arithmetic): only the sizes carry rounding noise
            d = self._rng.uniform(-self.noise[1], self.noise[1], (len(tlwh), 2)).astype(np.float32)
            tlwh = np.concatenate([tlwh[:, :2] - d / 2, tlwh[:, 2:] + d], 1).astype(np.float32)
        out = self.tracker.update(tlwh, conf, self.scenes[si], torch.from_numpy(cls))
        return np.asarray(out, np.int32).reshape(-1, 6), det


def clip_from_schedule(schedule):
    return [int(s) for s, n in schedule for _ in range(int(n))]


def ids_of_run(make, clip, **kw):
    """Per frame: the (track id, class) columns of the tracker rows of one oracle run over `clip` (scene index per frame)."""
    o = make(**kw)
    out = []
    for si in clip:
        rows, _ = o.step(si)
        out.append(None if rows is None else rows[:, 4:].copy())
    return out


def robust(make, clip, probes=((1, 0.3, 5e-3), (2, 0.3, 5e-3))):
    """True if the ids of the fp32 run, the half-detector run and the noise-probed runs of both agree on every frame."""
    ref = ids_of_run(make, clip, half_detector=False)
    runs = [dict(half_detector=True)] + [dict(half_detector=h, noise=p) for p in probes for h in (False, True)]
    for kw in runs:
        got = ids_of_run(make, clip, **kw)
        for t, (a, b) in enumerate(zip(ref, got)):
            if (a is None) != (b is None) or (a is not None and (a.shape != b.shape or not np.array_equal(a, b))):
                return False, t, kw
    return True, None, None


def find_schedule(make, n_scenes, n_frames=64, cycles=2, tries=60, seed=0, verbose=False):
    """A (scene, hold) schedule of n_frames frames that uses every scene, revisits some within max_age, and on which
    the track ids are robust (robust()) over `cycles` repetitions.  Held scenes + rejection: SURVEY 7 ("build e2e synthetic data
    with margins around thresholds").  Association decisions (gate 5.99, max_dist, max_iou_distance, LSAP near-ties) cannot be
    given margins by construction the way the detector's can, so schedules are drawn until one passes the noise probes."""
    for k in range(tries):
        rng = np.random.default_rng(seed * 1000 + k)
        sched, total, prev = [], 0, -1
        unused = list(range(n_scenes))
        while total < n_frames:
            if unused and rng.random() < 0.7:
                s = unused.pop(int(rng.integers(len(unused))))
            else:
                s = int(rng.integers(n_scenes))
                if s == prev:
                    continue
                if s in unused:
                    unused.remove(s)
            hold = int(rng.choice([2, 3, 5, 6, 8, 8, 8, 10]))
            hold = min(hold, n_frames - total)
            sched.append((s, hold)); total += hold; prev = s
        if unused or sched[0][0] == sched[-1][0]:
            continue
        ok, t, kw = robust(make, clip_from_schedule(sched) * cycles)
        if verbose:
            print("  schedule try %d: %s -> %s" % (k, sched, "robust" if ok else f"ids differ at frame {t} under {kw}"))
        if ok:
            return sched
    raise RuntimeError("no robust schedule found")


class ParityCheck:
    """Stage-by-stage parity of a CUDA run of a held-scene clip against the oracle, on the REAL data flow of the run:

      detector    per frame, against the fp32 oracle AND the half-storage oracle: same number of detections, same classes, box
                  centres bit-exact (the calibrated heads saturate them), sizes within a few tenths of a pixel, the integer crop
                  rectangles of the ReID stage identical -- and the same ORDER; frames on which the same set comes out in another
                  order are counted in `frames_with_other_detection_order` (paired by the exact centres), so the caller decides;
      ReID        the features the CUDA association consumed against the fp32 oracle's features of the same crops;
      association the oracle tracker is fed, frame by frame, exactly the (boxes, features, class ids) the CUDA tracker was fed
                  (ydst_pipeline_last_inputs) and must return bit-identical (K,6) rows -- ids, classes AND int32 boxes -- on
                  every frame: the association stage is bit-exact on the inputs it actually sees;
      end to end  the free-running oracles (from pixels) are tracked as well; `e2e_first_id_mismatch` is the first frame on which
                  their ids differ from the CUDA run's (None = never).  Association decisions at scene changes (Mahalanobis gate,
                  IoU >= 0.3, LSAP near-ties among 50 x 50 weakly discriminative appearance costs) are not stable against the
                  0.3 px / 1e-3 differences between ANY two arithmetic pipelines -- the fp32 and half-storage oracles part ways
                  with each other the same way -- so that number documents the reference's own sensitivity."""

    def __init__(self, blocks, ws, reid_sd, scenes, thres, nms_thres, class_mask, tracker_kw):
        from . import sort_ref as S
        mk = lambda half: ClipOracle(blocks, ws, reid_sd, scenes, thres, nms_thres, class_mask, tracker_kw, half)
        self.o32, self.o16 = mk(False), mk(True)
        kw = {k: v for k, v in tracker_kw.items() if k != "min_confidence"}
        self._feat = None
        self.forced = S.DeepSortRef(lambda fr, tl: self._feat, **kw)
        self.scenes = scenes
        self.frames = 0
        self.det_equal = True
        self.assoc_equal = True
        self.rows = 0
        self.max_size_px = 0.0
        self.max_centre_px = 0.0
        self.max_score = 0.0
        self.max_feat_rel = 0.0
        self.max_box_rel = 0.0
        self.e2e_first = {"fp32": None, "half": None}
        self.e2e_oracles_part = None
        self.problems = []
        self.order_swaps = {"fp32": 0, "half": 0}

    def frame(self, si, dets, rows, inputs):
        """si: scene shown; dets (n,6) float32 and rows (K,6) int32 | None from the CUDA run; inputs = (tlwh, feats, cls)."""
        t = self.frames
        self.frames += 1
        crop = lambda d: np.stack([np.maximum(d[:, 0].astype(np.int64), 0), np.maximum(d[:, 1].astype(np.int64), 0),
                                   (d[:, 0] + (d[:, 2] - d[:, 0])).astype(np.int64), (d[:, 1] + (d[:, 3] - d[:, 1])).astype(np.int64)], 1)
        for name, o in (("fp32", self.o32), ("half", self.o16)):
            ref = o.detect(si)[0]
            if ref is None or dets is None or ref.shape != dets.shape:
                self.det_equal = False
                self.problems.append(f"frame {t}: {0 if dets is None else len(dets)} detections vs {0 if ref is None else len(ref)} in the {name} oracle")
                continue
            cg = 0.5 * (dets[:, :2] + dets[:, 2:4]); cr = 0.5 * (ref[:, :2] + ref[:, 2:4])
            perm = np.arange(len(ref))
            if np.abs(cg - cr).max(initial=0) > 1e-2:
                # not the same order: the centres are exact in every arithmetic, so they pair the two sets unambiguously
                d = np.abs(cg[None, :, :] - cr[:, None, :]).max(-1)
                perm = d.argmin(1)
                if d.min(1).max() > 1e-2 or len(set(perm.tolist())) != len(perm):
                    self.det_equal = False
                    self.problems.append(f"frame {t}: the detections are not the {name} oracle's")
                    continue
                self.order_swaps[name] += 1
                if len(self.problems) < 50:
                    self.problems.append(f"frame {t}: same detections as the {name} oracle, {int((perm != np.arange(len(perm))).sum())} rows in another order")
            g = dets[perm]
            sg = g[:, 2:4] - g[:, :2]; sr = ref[:, 2:4] - ref[:, :2]
            dc, dsz = float(np.abs(cg[perm] - cr).max(initial=0)), float(np.abs(sg - sr).max(initial=0))
            if not np.array_equal(g[:, 5], ref[:, 5]) or dsz > 4.0 or not np.array_equal(crop(g), crop(ref)):
                self.det_equal = False
                self.problems.append(f"frame {t}: classes / boxes / crop rectangles differ from the {name} oracle (size {dsz:.3f} px)")
                continue
            self.max_centre_px, self.max_size_px = max(self.max_centre_px, dc), max(self.max_size_px, dsz)
            self.max_score = max(self.max_score, float(np.abs(g[:, 4] - ref[:, 4]).max(initial=0)))
            self.max_box_rel = max(self.max_box_rel, float((np.abs(g[:, :4] - ref[:, :4]).max(1) / np.maximum(sr.min(1), 1.0)).max(initial=0)))
        tl, ft, cl = inputs
        # ReID stage: the CUDA features against the oracle's features for the same crops (valid when the crops are the oracle's)
        f32 = self.o32.features(si).numpy()
        if f32.shape == ft.shape and self.order_swaps["fp32"] == 0:
            self.max_feat_rel = max(self.max_feat_rel, float((np.linalg.norm(ft - f32, axis=1) / np.linalg.norm(f32, axis=1)).max(initial=0)))
        # association stage, teacher-forced with the CUDA path's own inputs
        if dets is not None and len(dets):
            self._feat = torch.from_numpy(np.ascontiguousarray(ft))
            want = np.asarray(self.forced.update(tl, None, self.scenes[si], torch.from_numpy(cl.astype(np.float32))), np.int32).reshape(-1, 6)
            got = np.asarray(rows if rows is not None else [], np.int32).reshape(-1, 6)
            self.rows += len(got)
            if want.shape != got.shape or not np.array_equal(want, got):
                self.assoc_equal = False
                self.problems.append(f"frame {t}: track rows differ from the oracle tracker fed with the same inputs ({len(got)} vs {len(want)} rows)")
        # free-running oracles (informative)
        r32, _ = self.o32.step(si)
        r16, _ = self.o16.step(si)
        got = np.asarray(rows if rows is not None else [], np.int32).reshape(-1, 6)
        for name, r in (("fp32", r32), ("half", r16)):
            r = np.zeros((0, 6), np.int32) if r is None else r
            if self.e2e_first[name] is None and (r.shape != got.shape or not np.array_equal(r[:, 4:], got[:, 4:])):
                self.e2e_first[name] = t
        a, b = (np.zeros((0, 6), np.int32) if r is None else r for r in (r32, r16))
        if self.e2e_oracles_part is None and (a.shape != b.shape or not np.array_equal(a[:, 4:], b[:, 4:])):
            self.e2e_oracles_part = t

    def summary(self):
        return {"frames": self.frames, "track_rows": self.rows, "detections_equal": self.det_equal, "ids_equal": self.assoc_equal,
                "frames_with_other_detection_order": dict(self.order_swaps),
                "max_centre_px": round(self.max_centre_px, 6), "max_size_px": round(self.max_size_px, 4),
                "max_box_rel": round(self.max_box_rel, 6), "max_score_err": round(self.max_score, 5),
                "max_feature_rel": round(self.max_feat_rel, 6),
                "e2e_first_id_mismatch": dict(self.e2e_first), "oracles_part_ways_at": self.e2e_oracles_part,
                "note": "ids_equal: oracle association fed the CUDA path's own per-frame inputs returns bit-identical (K,6) rows; "
                        "detections_equal: count, classes, order, crop rectangles vs the fp32 and half-storage oracles from pixels"}
