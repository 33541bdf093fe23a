"""Oracle: DeepSORT association core + DeepSort.update facade (CPU; torch fp32 for the floating
point stages, plain Python/numpy for index and lifecycle work).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, with a struct-of-arrays tracker instead of per-track objects, but the SAME fp32 ATen ops in
the same order (so on CPU it is bit-identical to the reference; checked by oracle/gen_golden.py):
  kf_*                 deep_sort/sort/kalman_filter.py:54-87 (initiate), :89-123 (predict),
                       :125-159 (project), :161-204 (update; torch.solve == LU solve), :206-256 (gating)
  cosine_cost          deep_sort/sort/nn_matching.py:30-53, :77-100, :158-187
  gate                 deep_sort/sort/linear_assignment.py:147-203
  iou_cost             deep_sort/sort/iou_matching.py:5-91, deep_sort/sort/track.py:81-94
  min_cost_matching    deep_sort/sort/linear_assignment.py:8-73 (and :139-142: the cascade is one solve)
  TrackerRef           deep_sort/sort/tracker.py:49-176, deep_sort/sort/track.py:63-152,
                       deep_sort/sort/nn_matching.py:139-156
  DeepSortRef.update   deep_sort/deep_sort.py:46-88
"""
import numpy as np
import torch

from .lsap_ref import lsap_ref

INFTY_COST = 1e+5
CHI2INV95_2 = 5.9915
TENTATIVE, CONFIRMED, DELETED = 1, 2, 3

_M = torch.eye(8, 8, dtype=torch.float32)
for _i in range(4):
    _M[_i, 4 + _i] = 1
_MOTION = _M.t()                                   # kalman_filter.py:27-31 (stored transposed)
_STD_POS = torch.tensor([[1. / 20, 1. / 20, 0, 1. / 20]])
_STD_VEL = torch.tensor([[1. / 160, 1. / 160, 0, 1. / 160]])


def kf_initiate(xyah):
    """xyah: (4,) tensor -> mean (1,8), cov (1,8,8)."""
    m = torch.as_tensor(xyah, dtype=torch.float32)
    mean = torch.cat([m, torch.zeros_like(m)], dim=-1).view(1, -1)
    h = m[3]
    std = torch.tensor([[2 * (1. / 20) * h, 2 * (1. / 20) * h, 1e-2, 2 * (1. / 20) * h,
                         10 * (1. / 160) * h, 10 * (1. / 160) * h, 1e-5, 10 * (1. / 160) * h]])
    return mean, torch.diag_embed(torch.pow(std, 2))


def kf_predict(mean, cov):
    std_pos = mean[:, 3:4] * _STD_POS
    std_vel = mean[:, 3:4] * _STD_VEL
    std_pos[:, 2] = 1e-2
    std_vel[:, 2] = 1e-5
    q = torch.diag_embed(torch.pow(torch.cat([std_pos, std_vel], dim=-1), 2))
    mean = torch.matmul(mean, _MOTION)
    cov = torch.matmul(torch.matmul(cov.permute(0, 2, 1), _MOTION).permute(0, 2, 1), _MOTION)
    return mean, cov + q


def kf_project(mean, cov):
    std = mean[:, 3:4] * _STD_POS
    std[:, 2] = 1e-1
    return mean[:, :4].clone(), cov[:, :4, :4].clone() + torch.diag_embed(torch.pow(std, 2))


def kf_update(mean, cov, meas):
    pm, pc = kf_project(mean, cov)
    upd = torch.eye(8, 4, dtype=torch.float32)
    gain = torch.linalg.solve(pc, torch.matmul(cov, upd).permute(0, 2, 1)).permute(0, 2, 1)
    innov = meas.view(-1, 4) - pm
    gt = gain.permute(0, 2, 1)
    new_mean = mean + torch.bmm(innov.unsqueeze(1), gt).view(-1, 8)
    new_cov = cov - torch.matmul(torch.matmul(pc.permute(0, 2, 1), gt).permute(0, 2, 1), gt)
    return new_mean, new_cov


def kf_gating_position(mean, cov, meas_xyah, row_chunk=64):
    """gating_distance(only_position=True).  The reference builds an N x M x M temporary
    (kalman_filter.py:253); each batch element of a bmm is independent, so evaluating it in row
    chunks is the same arithmetic with bounded memory."""
    pm, pc = kf_project(mean, cov)
    out = []
    for s in range(0, pm.shape[0], row_chunk):
        m, c = pm[s:s + row_chunk, None, :2], pc[s:s + row_chunk, :2, :2]
        d = -m + meas_xyah[None, :, :2]
        sq = torch.bmm(torch.bmm(d, torch.inverse(c)), d.permute(0, 2, 1))
        out.append(torch.diagonal(sq, dim1=-2, dim2=-1))
    return torch.cat(out, 0)


def tlwh_to_xyah(tlwh):
    m = tlwh.clone()
    m[:, :2] += m[:, 2:] / 2
    m[:, 2] /= m[:, 3]
    return m


def cosine_cost(gallery_rows, bp, feats):
    """_nn_cosine_distance: normalise both sides, 1 - A.B^T, min over each track's rows."""
    a = gallery_rows / torch.norm(gallery_rows, dim=-1, keepdim=True)
    b = feats / torch.norm(feats, dim=-1, keepdim=True)
    d = 1. - torch.mm(a, b.t())
    return torch.stack([d[bp[i]:bp[i + 1]].min(dim=0)[0] for i in range(len(bp) - 1)], dim=0)


def iou_cost(track_tlwh, det_tlwh):
    b, c = track_tlwh.unsqueeze(1), det_tlwh.unsqueeze(0)
    inter_wh = torch.clamp(torch.min(b[..., :2] + b[..., 2:], c[..., 2:] + c[..., :2])
                           - torch.max(b[..., :2], c[..., :2]) + 1, min=0)
    inter = inter_wh[..., 0] * inter_wh[..., 1]
    return 1. - inter / (b[..., 2] * b[..., 3] + c[..., 2] * c[..., 3] - inter)


def split_assignment(cost, max_distance, track_indices, detection_indices, rows, cols):
    """The list-building half of min_cost_matching (linear_assignment.py:58-72); the ORDER of
    unmatched_detections decides new track ids."""
    rows, cols = [int(r) for r in rows], [int(c) for c in cols]
    rs, cs = set(rows), set(cols)
    matches = []
    um_d = [d for c, d in enumerate(detection_indices) if c not in cs]
    um_t = [t for r, t in enumerate(track_indices) if r not in rs]
    for r, c in zip(rows, cols):
        if cost[r, c] > max_distance:
            um_t.append(track_indices[r]); um_d.append(detection_indices[c])
        else:
            matches.append((track_indices[r], detection_indices[c]))
    return matches, um_t, um_d


def min_cost_matching(cost_fn, max_distance, track_indices, detection_indices, lsap=lsap_ref):
    if len(detection_indices) == 0 or len(track_indices) == 0:
        return [], list(track_indices), list(detection_indices)
    cost = cost_fn(track_indices, detection_indices)
    cost[cost > max_distance] = max_distance + 1e-5
    cost = cost.numpy()
    rows, cols = lsap(cost)
    return split_assignment(cost, max_distance, track_indices, detection_indices, rows, cols)


class TrackerRef:
    """Tracker + Track + NearestNeighborDistanceMetric, struct-of-arrays."""

    def __init__(self, max_dist=0.2, max_iou_distance=0.7, max_age=70, n_init=3, budget=100, lsap=lsap_ref):
        self.max_dist, self.max_iou, self.max_age, self.n_init, self.budget = max_dist, max_iou_distance, max_age, n_init, budget
        self.lsap = lsap
        self.mean, self.cov = [], []                      # per track (1,8), (1,8,8)
        self.ids, self.hits, self.age, self.tsu, self.state, self.payload = [], [], [], [], [], []
        self.pending = []                                 # Track.features (list per track)
        self.samples = {}                                 # metric.samples: id -> list of (512,)
        self.next_id = 1
        self.debug = {}

    def __len__(self):
        return len(self.ids)

    def predict(self):
        if not len(self):
            return
        m, c = kf_predict(torch.cat(self.mean, 0), torch.cat(self.cov, 0))
        for i in range(len(self)):
            self.mean[i], self.cov[i] = m[i].unsqueeze(0), c[i].unsqueeze(0)
            self.age[i] += 1
            self.tsu[i] += 1

    def _track_tlwh(self, i):
        r = self.mean[i].flatten()[:4].clone()
        r[2] *= r[3]
        r[:2] -= r[2:] / 2
        return r

    def _match(self, det_tlwh, det_feat):
        ndet = det_tlwh.shape[0]

        def gated_metric(tis, dis):
            feats = torch.stack([det_feat[i] for i in dis], 0)
            bp, rows = [0], []
            for t in tis:
                s = self.samples[self.ids[t]]
                rows += s
                bp.append(bp[-1] + len(s))
            cost = cosine_cost(torch.stack(rows, 0), bp, feats)
            meas = tlwh_to_xyah(torch.stack([det_tlwh[i] for i in dis], 0))
            g = kf_gating_position(torch.cat([self.mean[t] for t in tis], 0),
                                   torch.cat([self.cov[t] for t in tis], 0), meas)
            cost[g > CHI2INV95_2] = INFTY_COST
            return cost

        def iou_metric(tis, dis):
            cost = iou_cost(torch.stack([self._track_tlwh(t) for t in tis], 0),
                            torch.stack([det_tlwh[i] for i in dis], 0))
            for r, t in enumerate(tis):
                if self.tsu[t] > 1:
                    cost[r, :] = INFTY_COST
            return cost

        confirmed = [i for i in range(len(self)) if self.state[i] == CONFIRMED]
        unconfirmed = [i for i in range(len(self)) if self.state[i] != CONFIRMED]
        m_a, um_t_a, um_d = min_cost_matching(gated_metric, self.max_dist, confirmed, list(range(ndet)), self.lsap)
        iou_cand = unconfirmed + [k for k in um_t_a if self.tsu[k] == 1]
        um_t_a = [k for k in um_t_a if self.tsu[k] != 1]
        m_b, um_t_b, um_d = min_cost_matching(iou_metric, self.max_iou, iou_cand, um_d, self.lsap)
        self.debug = dict(matches_a=m_a, matches_b=m_b)
        return m_a + m_b, list(set(um_t_a + um_t_b)), um_d

    def update(self, det_tlwh, det_feat, det_payload):
        """det_tlwh (m,4) f32 tensor, det_feat (m,512) f32 tensor, det_payload: sequence of m class ids."""
        matches, um_t, um_d = self._match(det_tlwh, det_feat)
        if matches:
            meas = tlwh_to_xyah(torch.stack([det_tlwh[d] for _, d in matches], 0))
            nm, ncov = kf_update(torch.cat([self.mean[t] for t, _ in matches], 0),
                                 torch.cat([self.cov[t] for t, _ in matches], 0), meas)
            for i, (t, d) in enumerate(matches):
                self.mean[t], self.cov[t] = nm[i].unsqueeze(0), ncov[i].unsqueeze(0)
                self.pending[t].append(det_feat[d])
                self.hits[t] += 1
                self.tsu[t] = 0
                if self.state[t] == TENTATIVE and self.hits[t] >= self.n_init:
                    self.state[t] = CONFIRMED
                self.payload[t] = det_payload[d]
        for t in um_t:                                   # Track.mark_missed
            if self.state[t] == TENTATIVE or self.tsu[t] > self.max_age:
                self.state[t] = DELETED
        for d in um_d:                                   # _initiate_track, in list order
            m, c = kf_initiate(tlwh_to_xyah(det_tlwh[d:d + 1])[0])
            self.mean.append(m); self.cov.append(c)
            self.ids.append(self.next_id); self.hits.append(1); self.age.append(1); self.tsu.append(0)
            self.state.append(TENTATIVE); self.payload.append(det_payload[d]); self.pending.append([det_feat[d]])
            self.next_id += 1
        keep = [i for i in range(len(self)) if self.state[i] != DELETED]
        for name in ("mean", "cov", "ids", "hits", "age", "tsu", "state", "payload", "pending"):
            setattr(self, name, [getattr(self, name)[i] for i in keep])
        active = [self.ids[i] for i in range(len(self)) if self.state[i] == CONFIRMED]
        for i in range(len(self)):                       # partial_fit
            if self.state[i] != CONFIRMED:
                continue
            for f in self.pending[i]:
                self.samples.setdefault(self.ids[i], []).append(f)
                if self.budget is not None:
                    self.samples[self.ids[i]] = self.samples[self.ids[i]][-self.budget:]
            self.pending[i] = []
        self.samples = {k: self.samples[k] for k in active}
        self.debug.update(unmatched_tracks=sorted(um_t), unmatched_dets=list(um_d))

    def outputs(self):
        """DeepSort.update output block (deep_sort/deep_sort.py:63-88): np.int32 (K,6) or []."""
        idx = [i for i in range(len(self)) if self.state[i] == CONFIRMED and self.tsu[i] <= 1]
        if not idx:
            return []
        b = torch.cat([self.mean[i] for i in idx], 0)[:, :4]
        b[:, 2] *= b[:, 3]
        b[:, :2] -= b[:, 2:] / 2
        b[:, 2:] += b[:, :2]
        b[:, :2] = torch.clamp(b[:, :2], min=0)
        rows = [[b[k, 0], b[k, 1], b[k, 2], b[k, 3], self.ids[i], self.payload[i]] for k, i in enumerate(idx)]
        return np.array(rows, dtype=np.int32)

    def state_arrays(self):
        """Snapshot for parity checks: ids, hits, age, tsu, state (int), mean (n,8), cov (n,8,8)."""
        n = len(self)
        return dict(ids=np.asarray(self.ids, np.int32), hits=np.asarray(self.hits, np.int32),
                    age=np.asarray(self.age, np.int32), tsu=np.asarray(self.tsu, np.int32),
                    state=np.asarray(self.state, np.int32),
                    mean=torch.cat(self.mean, 0).numpy().copy() if n else np.zeros((0, 8), np.float32),
                    cov=torch.cat(self.cov, 0).numpy().copy() if n else np.zeros((0, 8, 8), np.float32))


class DeepSortRef:
    """DeepSort facade (deep_sort/deep_sort.py:16-88) with an injectable feature function
    ``features(frame, tlwh) -> (m,512) tensor`` (the reference's extractor injection point, :28-31)."""

    def __init__(self, features, max_dist=0.2, max_iou_distance=0.7, max_age=70, n_init=3, nn_budget=100, lsap=lsap_ref):
        self.features = features
        self.tracker = TrackerRef(max_dist, max_iou_distance, max_age, n_init, nn_budget, lsap)

    def update(self, bbox_tlwh, confidences, frame, payload):
        tlwh = torch.as_tensor(np.asarray(bbox_tlwh, np.float32))
        feats = self.features(frame, tlwh.numpy()) if len(tlwh) else torch.zeros((0, 512))
        self.tracker.predict()
        self.tracker.update(tlwh, torch.as_tensor(feats), list(payload))
        return self.tracker.outputs()
