"""CPU restatement of the reference's rule-based action recognition.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows action/action_Identify.py:15-47 (cache update, ageing, deletion, rule evaluation in dict order), action/orbit.py:5-26
(a deque(maxlen) of bottom-centre points (x1 + (x2 - x1) / 2, y2) and time stamps) and action/actions.py:23-150.  The rules'
`is_x` flag logic reduces to: the pair condition holds for EVERY consecutive pair of the deque and there is at least one pair.
Pinned by tests/golden/action.npz, written by the unmodified reference (oracle/gen_golden.py gen_action).
"""
from collections import OrderedDict

import numpy as np

TAKEOFF, LANDING, GLIDE, FAST_CROSSING, BREAK_INTO = range(5)
KIND = {"TakeOff": TAKEOFF, "Landing": LANDING, "Glide": GLIDE, "FastCrossing": FAST_CROSSING, "BreakInto": BREAK_INTO}


def pair_condition(kind, prm, a, b, ta, tb):
    """condition of one consecutive pair (older point a, newer point b), float64 (actions.py:37,61,84,109)."""
    if kind == TAKEOFF:
        return a[1] - b[1] > prm[1] and abs(a[0] - b[0]) > prm[0]
    if kind == LANDING:
        return b[1] - a[1] > prm[1] and abs(a[0] - b[0]) > prm[0]
    if kind == GLIDE:
        return abs(b[1] - a[1]) < prm[1] and abs(b[0] - a[0]) > prm[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        return bool(np.float64(abs(b[0] - a[0])) / np.float64((tb - ta) * 1000) > prm)


class ActionIdentifyRef:
    """rules: [(kind name, class_id, parameter)] with parameter = (dx, dy) | speed | timeout."""

    def __init__(self, rules, max_age=30, max_size=4):
        self.rules = [(KIND[n], c, p) for n, c, p in rules]
        self.max_age, self.max_size = max_age, max_size
        self.cache = OrderedDict()                     # track id -> {cls, age, pts, ts}; insertion order is the output order

    def update(self, rows, now):
        """rows (K,6) int [x1,y1,x2,y2,id,cls] -> [(track id, class id, rule index)]"""
        rows = np.asarray(rows, np.int64).reshape(-1, 6)
        seen = set()
        for x1, y1, x2, y2, tid, cls in rows:
            seen.add(int(tid))
            o = self.cache.get(int(tid))
            if o is None:                              # a new orbit starts EMPTY (action_Identify.py:24)
                self.cache[int(tid)] = {"cls": int(cls), "age": 0, "pts": [], "ts": []}
                continue
            o["age"] = 0
            o["pts"] = (o["pts"] + [(x1 + (x2 - x1) / 2, float(y2))])[-self.max_size:]
            o["ts"] = (o["ts"] + [now])[-self.max_size:]
        for tid in [t for t in self.cache if t not in seen]:
            self.cache[tid]["age"] += 1
            if self.cache[tid]["age"] >= self.max_age:
                del self.cache[tid]
        out = []
        for tid, o in self.cache.items():
            if o["age"] != 0:
                continue
            n = len(o["pts"])
            for r, (kind, cid, prm) in enumerate(self.rules):
                if n == 0 or o["cls"] != cid:
                    continue
                if kind == BREAK_INTO:
                    ok = n > prm
                else:
                    ok = n >= 2 and all(pair_condition(kind, prm, o["pts"][k - 1], o["pts"][k], o["ts"][k - 1], o["ts"][k]) for k in range(1, n))
                if ok:
                    out.append((tid, o["cls"], r))
        return out


def action_sequence(seed=0, n_frames=60):
    """A synthetic (K,6) track-row sequence that exercises every rule of action/actions.py: tracks of two classes drifting up /
    down / sideways by seeded steps around the rule thresholds, tracks that disappear for a while (ageing, deletion at max_age and
    re-insertion at the END of the cache) and frames without any row.  Time stamps are a seeded increasing sequence (the reference
    reads time.time(); the generator patches it)."""
    rng = np.random.default_rng(seed)
    n_tr = 14
    pos = rng.uniform(100, 500, (n_tr, 2))
    vel = rng.uniform(-14, 14, (n_tr, 2))
    vel[:3, 1] = -rng.uniform(6, 12, 3); vel[:3, 0] = rng.uniform(5, 9, 3)        # take-off like
    vel[3:6, 1] = rng.uniform(6, 12, 3); vel[3:6, 0] = -rng.uniform(5, 9, 3)       # landing like
    vel[6:9, 1] = rng.uniform(-1.5, 1.5, 3); vel[6:9, 0] = rng.uniform(6, 30, 3)   # glide / fast crossing
    size = rng.uniform(20, 80, (n_tr, 2))
    cls = rng.integers(0, 2, n_tr) * 4                                            # classes 0 and 4
    frames, stamps, t = [], [], 1000.0
    for f in range(n_frames):
        t += float(rng.uniform(0.01, 0.06))
        stamps.append(t)
        if f in (17, 18, 41):
            frames.append(np.zeros((0, 6), np.int32))
            continue
        rows = []
        for k in range(n_tr):
            # track k is absent during a window of its own (some longer than max_age = 6 used by the test)
            gap0 = 8 + 3 * k
            if gap0 <= f < gap0 + (3 if k % 2 else 9):
                continue
            if f < k // 3:
                continue
            pos[k] += vel[k] + rng.uniform(-2, 2, 2)
            x1, y1 = pos[k] - size[k] / 2
            x2, y2 = pos[k] + size[k] / 2
            rows.append([int(x1), int(y1), int(x2), int(y2), k + 1, int(cls[k])])
        order = rng.permutation(len(rows))
        frames.append(np.asarray(rows, np.int32).reshape(-1, 6)[order])
    return frames, stamps


ACTION_RULES = (("TakeOff", 4, (4.0, 5.0)), ("Landing", 4, (4.0, 5.0)), ("Glide", 0, (5.0, 3.0)), ("FastCrossing", 0, 0.4),
                ("BreakInto", 4, 2), ("Glide", 4, (5.0, 3.0)))
