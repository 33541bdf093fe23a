"""Parity of the association kernels (Kalman, gating, cost matrices, LSAP, full tracker) against the oracle and the
golden vectors produced by the unmodified reference.  Integer/index results (assignments, track ids, hit counters,
output rows) must be bit-exact; Kalman state is compared at 1e-5 relative (fp32, different but equivalent operation
order inside the 4x4 solve), Mahalanobis/cosine costs at 1e-5 absolute."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from util import DEV
from yolo_deepsort_b200._lib import check, lib, ptr, stream_ptr

pytestmark = pytest.mark.gpu


_KEEP = []          # device tensors must outlive the asynchronous kernels that read their raw pointers


def dev(a, dtype=torch.float32):
    t = torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(DEV).contiguous()
    _KEEP.append(t)
    return t


@pytest.fixture(autouse=True)
def _release_device_tensors():
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def xyah_to_tlwh(z):
    z = np.asarray(z, np.float64)
    w = z[:, 2] * z[:, 3]
    return np.stack([z[:, 0] - w / 2, z[:, 1] - z[:, 3] / 2, w, z[:, 3]], 1).astype(np.float32)


def test_kalman_known_answer():
    """The reference's only known-answer snippet (deep_sort/sort/kalman_filter.py:259-273)."""
    g = np.load(os.path.join(GOLDEN, "kalman_demo.npz"))
    tl = dev(xyah_to_tlwh([[10, 15, 0.5, 10]]))
    mean = torch.zeros((1, 8), device=DEV)
    cov = torch.zeros((1, 8, 8), device=DEV)
    check(lib().ydst_kf_initiate(ptr(tl), 1, ptr(mean), ptr(cov), stream_ptr()))
    np.testing.assert_array_equal(torch.diagonal(cov[0]).cpu().numpy(), g["diag_init"])
    check(lib().ydst_kf_predict(ptr(mean), ptr(cov), 1, stream_ptr()))
    np.testing.assert_array_equal(torch.diagonal(cov[0]).cpu().numpy(), g["diag_pred"])
    z = dev(xyah_to_tlwh([[12, 20, 0.6, 11]]))
    check(lib().ydst_kf_update(ptr(mean), ptr(cov), ptr(z), 1, stream_ptr()))
    np.testing.assert_allclose(mean.cpu().numpy(), g["mean_upd"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(cov.cpu().numpy(), g["cov_upd"], rtol=1e-4, atol=1e-7)


def test_kalman_batch_golden():
    g = np.load(os.path.join(GOLDEN, "kalman_batch.npz"))
    n = g["mean0"].shape[0]
    mean, cov = dev(g["mean0"]), dev(g["cov0"])
    check(lib().ydst_kf_predict(ptr(mean), ptr(cov), n, stream_ptr()))
    # predict is sums of at most two fp32 terms per element in a fixed association: bit-exact
    np.testing.assert_array_equal(mean.cpu().numpy(), g["mean1"])
    np.testing.assert_array_equal(cov.cpu().numpy(), g["cov1"])
    # update: measurement given as tlwh whose xyah conversion is exact enough -> feed xyah through a tlwh that maps back
    z = g["z"]
    tl = xyah_to_tlwh(z)
    mean, cov = dev(g["mean1"]), dev(g["cov1"])
    check(lib().ydst_kf_update(ptr(mean), ptr(cov), ptr(dev(tl)), n, stream_ptr()))
    np.testing.assert_allclose(mean.cpu().numpy(), g["mean2"], rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(cov.cpu().numpy(), g["cov2"], rtol=1e-3, atol=1e-5)
    # gating distance
    dets = g["dets_xyah"]
    m = dets.shape[0]
    maha = torch.zeros((n, m), device=DEV)
    check(lib().ydst_gate_position(ptr(dev(g["mean3"])), ptr(dev(g["cov3"])), n, ptr(dev(xyah_to_tlwh(dets))), m, ptr(maha), stream_ptr()))
    ref = g["gate2"]
    np.testing.assert_allclose(maha.cpu().numpy(), ref, rtol=2e-4, atol=1e-3)


def test_kalman_vs_oracle_large():
    """N = 2000 tracks (config 5 size): predict bit-exact, update within tolerance; ragged n (not a multiple of 4)."""
    from oracle import sort_ref as S
    rng = np.random.default_rng(0)
    for n in (1, 3, 2000, 2001):
        tl = np.stack([rng.uniform(0, 500, n), rng.uniform(0, 500, n), rng.uniform(20, 80, n), rng.uniform(40, 160, n)], 1).astype(np.float32)
        mean, cov = torch.zeros((n, 8), device=DEV), torch.zeros((n, 8, 8), device=DEV)
        check(lib().ydst_kf_initiate(ptr(dev(tl)), n, ptr(mean), ptr(cov), stream_ptr()))
        om, oc = zip(*[S.kf_initiate(S.tlwh_to_xyah(torch.from_numpy(tl[i:i + 1]))[0]) for i in range(n)])
        om, oc = torch.cat(om, 0), torch.cat(oc, 0)
        np.testing.assert_array_equal(mean.cpu().numpy(), om.numpy())
        np.testing.assert_array_equal(cov.cpu().numpy(), oc.numpy())
        for _ in range(3):
            check(lib().ydst_kf_predict(ptr(mean), ptr(cov), n, stream_ptr()))
            om, oc = S.kf_predict(om, oc)
        np.testing.assert_array_equal(mean.cpu().numpy(), om.numpy())
        np.testing.assert_array_equal(cov.cpu().numpy(), oc.numpy())
        z = (tl + rng.normal(0, 2, tl.shape)).astype(np.float32)
        check(lib().ydst_kf_update(ptr(mean), ptr(cov), ptr(dev(z)), n, stream_ptr()))
        om, oc = S.kf_update(om, oc, S.tlwh_to_xyah(torch.from_numpy(z)))
        np.testing.assert_allclose(mean.cpu().numpy(), om.numpy(), rtol=2e-5, atol=2e-4)
        np.testing.assert_allclose(cov.cpu().numpy(), oc.numpy(), rtol=1e-3, atol=1e-5)


def lsap_abi(cost, max_dist=0.3):
    nr, nc = cost.shape
    k = min(nr, nc)
    rows, cols, over = np.zeros(max(k, 1), np.int32), np.zeros(max(k, 1), np.int32), np.zeros(max(k, 1), np.int32)
    c = dev(cost)
    check(lib().ydst_lsap(ptr(c), nr, nc, float(max_dist), rows.ctypes.data, cols.ctypes.data, over.ctypes.data, stream_ptr()))
    return rows[:k], cols[:k], over[:k]


@pytest.mark.parametrize("kind", ["random", "integer_ties", "clamped", "gated"])
def test_lsap_matches_oracle(kind):
    """Exact assignment incl. scipy's tie-breaking, rectangular shapes both ways, all three kernel widths."""
    from oracle.lsap_ref import lsap_ref
    rng = np.random.default_rng({"random": 1, "integer_ties": 2, "clamped": 3, "gated": 4}[kind])
    shapes = [(1, 1), (1, 7), (7, 1), (5, 5), (17, 40), (40, 17), (50, 50), (96, 96), (97, 130), (200, 120), (300, 300), (64, 1100)]
    for nr, nc in shapes:
        c = rng.random((nr, nc))
        if kind == "integer_ties":
            c = rng.integers(0, 4, (nr, nc)).astype(float)
        elif kind == "clamped":
            c[c > 0.3] = 0.30001
        elif kind == "gated":
            c[rng.random((nr, nc)) > 0.1] = 1e5
            c[c > 0.3] = 0.30001
        c = c.astype(np.float32)
        r0, c0 = lsap_ref(c)
        r1, c1, over = lsap_abi(c)
        np.testing.assert_array_equal(r1, r0, err_msg=f"{kind} {nr}x{nc} rows")
        np.testing.assert_array_equal(c1, c0, err_msg=f"{kind} {nr}x{nc} cols")
        np.testing.assert_array_equal(over, (c[r0, c0] > np.float32(0.3)).astype(np.int32))


def test_lsap_2000():
    """Config 5 size: 2000 x 2000, 2 %-valid gated cost matrix; bit-exact indices and a checksum of the optimum."""
    from oracle.lsap_ref import lsap_ref
    rng = np.random.default_rng(5)
    c = rng.random((2000, 2000)).astype(np.float32)
    c[rng.random((2000, 2000)) > 0.02] = 1e5
    c[c > 0.3] = np.float32(0.3 + 1e-5)
    r0, c0 = lsap_ref(c)
    r1, c1, _ = lsap_abi(c)
    np.testing.assert_array_equal(c1, c0)
    assert sorted(c1.tolist()) == list(range(2000))          # a permutation
    assert abs(float(c[r1, c1].astype(np.float64).sum()) - float(c[r0, c0].astype(np.float64).sum())) == 0.0


def test_costs_vs_oracle():
    from oracle import sort_ref as S
    from oracle.synth import unit_rows
    rng = np.random.default_rng(9)
    n, m, budget = 37, 53, 30
    tl = np.stack([rng.uniform(0, 500, n), rng.uniform(0, 500, n), rng.uniform(20, 80, n), rng.uniform(40, 160, n)], 1).astype(np.float32)
    om, oc = zip(*[S.kf_initiate(S.tlwh_to_xyah(torch.from_numpy(tl[i:i + 1]))[0]) for i in range(n)])
    om, oc = S.kf_predict(torch.cat(om, 0), torch.cat(oc, 0))
    counts = rng.integers(1, budget + 1, n)
    seg = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    gal = unit_rows(rng, int(seg[-1])) * rng.uniform(0.9, 1.1, (int(seg[-1]), 1)).astype(np.float32)
    k = min(n, m - 10)
    dets = np.concatenate([tl[:k] + rng.normal(0, 3, (k, 4)), np.stack([rng.uniform(0, 500, m - k), rng.uniform(0, 500, m - k),
                          rng.uniform(20, 80, m - k), rng.uniform(40, 160, m - k)], 1)], 0).astype(np.float32)[:m]
    feats = unit_rows(rng, m)
    feats[:20] = gal[seg[:20]] / np.linalg.norm(gal[seg[:20]], axis=1, keepdims=True) + 0.02 * rng.standard_normal((20, 512)).astype(np.float32)
    # oracle
    cost = S.cosine_cost(torch.from_numpy(gal), seg.tolist(), torch.from_numpy(feats))
    gate = S.kf_gating_position(om, oc, S.tlwh_to_xyah(torch.from_numpy(dets)))
    raw = cost.clone().numpy()
    cost[gate > S.CHI2INV95_2] = S.INFTY_COST
    cost[cost > 0.3] = 0.3 + 1e-5
    out = torch.zeros((n, m), device=DEV)
    check(lib().ydst_appearance_cost(ptr(dev(gal)), seg.ctypes.data, n, ptr(dev(feats)), m, ptr(dev(om.numpy())), ptr(dev(oc.numpy())),
                                     ptr(dev(dets)), 0.3, ptr(out), stream_ptr()))
    got, ref = out.cpu().numpy(), cost.numpy()
    # entries far from both thresholds must agree to 1e-5; entries within 1e-4 of a threshold may fall on either side
    near = (np.abs(raw - 0.3) < 1e-4) | (np.abs(gate.numpy() - S.CHI2INV95_2) < 1e-2)
    assert near.mean() < 0.01
    np.testing.assert_allclose(got[~near], ref[~near], rtol=0, atol=1e-5)
    assert (got <= np.float32(0.3 + 1e-5)).all() and (ref[ref > 0.3] == np.float32(0.3 + 1e-5)).all()
    # IoU cost: same fp32 operations -> bit-exact
    tsu = np.ones(n, np.int32); tsu[::7] = 2
    tr_tlwh = torch.stack([torch.cat([om[i, :2] - torch.stack([om[i, 2] * om[i, 3], om[i, 3]]) / 2, torch.stack([om[i, 2] * om[i, 3], om[i, 3]])]) for i in range(n)], 0)
    ic = S.iou_cost(tr_tlwh, torch.from_numpy(dets))
    ic[torch.from_numpy(tsu) > 1] = S.INFTY_COST
    ic[ic > 0.7] = 0.7 + 1e-5
    out2 = torch.zeros((n, m), device=DEV)
    check(lib().ydst_iou_cost(ptr(dev(om.numpy())), ptr(dev(tsu, torch.int32)), n, ptr(dev(dets)), m, 0.7, ptr(out2), stream_ptr()))
    np.testing.assert_array_equal(out2.cpu().numpy(), ic.numpy())


def run_tracker_sequence(frames, params, expect_out, expect_tab, expect_mean=None):
    from yolo_deepsort_b200.deepsort import TrackerHandle
    max_dist, max_iou, max_age, n_init, budget = params
    trk = TrackerHandle(max_dist, max_iou, int(max_age), int(n_init), int(budget), 4096, 2048, DEV)
    for t, (tl, ft, cl) in enumerate(frames):
        out = trk.update(dev(tl), dev(ft), cl.astype(np.int32))
        np.testing.assert_array_equal(out.reshape(-1, 6), expect_out[t].reshape(-1, 6), err_msg=f"frame {t}: output rows (ids / boxes)")
        tab, mean = trk.table()
        np.testing.assert_array_equal(tab, expect_tab[t], err_msg=f"frame {t}: track table")
        if expect_mean is not None and len(mean):
            np.testing.assert_allclose(mean, expect_mean[t], rtol=1e-4, atol=1e-3, err_msg=f"frame {t}: track means")


@pytest.mark.parametrize("name", ["assoc_seq.npz", "assoc_seq2.npz"])
def test_tracker_golden_sequence(name):
    """DeepSort.update sequences produced by the UNMODIFIED reference: identical (K,6) int32 rows (track ids, class ids,
    truncated boxes) and identical [id,hits,age,tsu,state] tables every frame.  assoc_seq: 24 frames, demo parameters;
    assoc_seq2: 44 frames with nn_budget=4, max_age=3, n_init=2 -- gallery FIFO truncation (nn_matching.py:153-154), deletion by
    age (track.py:146-152) and re-identification after misses."""
    g = np.load(os.path.join(GOLDEN, name))
    T = int(g["n_frames"])
    frames = [(g[f"tlwh_{t}"], g[f"feat_{t}"].astype(np.float32), g[f"cls_{t}"]) for t in range(T)]
    run_tracker_sequence(frames, g["params"], [g[f"out_{t}"] for t in range(T)], [g[f"table_{t}"] for t in range(T)],
                         [g[f"mean_{t}"] for t in range(T)])


@pytest.mark.parametrize("n,frames", [(300, 12), (2000, 4)])
def test_tracker_vs_oracle_stress(n, frames):
    """Larger scenes against the oracle (config 5: 2000 tracks x ~2000 detections): bit-exact ids and lifecycle."""
    from oracle import sort_ref as S
    from oracle.synth import Scenario
    sc = Scenario(n=n, frame_hw=(2160, 3840), seed=21, p_miss=0.05, p_new=0.02, p_leave=0.01)
    seq = [sc.step() for _ in range(frames)]
    cur = {}
    orc = S.DeepSortRef(lambda fr, tl: torch.from_numpy(cur["f"]), max_dist=0.3, max_iou_distance=0.7, max_age=30, n_init=3, nn_budget=30)
    outs, tabs = [], []
    img = np.zeros((4, 4, 3), np.uint8)
    for tl, ft, cl in seq:
        cur["f"] = ft
        o = orc.update(tl.copy(), None, img, torch.from_numpy(cl))
        outs.append(np.asarray(o, np.int32).reshape(-1, 6))
        st = orc.tracker.state_arrays()
        tabs.append(np.stack([st["ids"], st["hits"], st["age"], st["tsu"], st["state"]], 1).reshape(-1, 5))
    run_tracker_sequence(seq, (0.3, 0.7, 30, 3, 30), outs, tabs)


def test_tracker_edge_cases():
    """Empty detection lists, a single detection, more detections than tracks, everything disappearing."""
    from oracle import sort_ref as S
    rng = np.random.default_rng(2)
    from oracle.synth import unit_rows
    def mk(m):
        tl = np.stack([rng.uniform(0, 500, m), rng.uniform(0, 500, m), rng.uniform(20, 80, m), rng.uniform(40, 160, m)], 1).astype(np.float32).reshape(-1, 4)
        return tl, unit_rows(rng, m).reshape(-1, 512), rng.choice([0, 2, 4], m).astype(np.float32)
    base = mk(6)
    seq = [mk(0), base, base, base, mk(0), base, mk(1), (base[0][:2], base[1][:2], base[2][:2]), mk(0), mk(0), mk(0), mk(9)]
    cur = {}
    orc = S.DeepSortRef(lambda fr, tl: torch.from_numpy(cur["f"]), max_dist=0.3, max_iou_distance=0.7, max_age=2, n_init=2, nn_budget=3)
    outs, tabs = [], []
    for tl, ft, cl in seq:
        cur["f"] = ft
        o = orc.update(tl.copy(), None, np.zeros((4, 4, 3), np.uint8), torch.from_numpy(cl))
        outs.append(np.asarray(o, np.int32).reshape(-1, 6))
        st = orc.tracker.state_arrays()
        tabs.append(np.stack([st["ids"], st["hits"], st["age"], st["tsu"], st["state"]], 1).reshape(-1, 5))
    run_tracker_sequence(seq, (0.3, 0.7, 2, 2, 3), outs, tabs)
