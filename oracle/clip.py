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
