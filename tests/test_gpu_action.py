"""Device ActionIdentify against the golden written by the reference's own action/ package
(tests/golden/action.npz) and against the oracle restatement on a second seeded sequence.  Exact: integer rows, float64 rules."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _rules():
    from oracle.action_ref import ACTION_RULES
    from yolo_deepsort_b200 import action as A
    return [getattr(A, name)(cid, prm) for name, cid, prm in ACTION_RULES], ACTION_RULES


def test_action_identify_matches_reference_golden():
    from yolo_deepsort_b200.action import ActionIdentify
    g = np.load(os.path.join(GOLDEN, "action.npz"))
    rules, spec = _rules()
    clock = {"t": 0.0}
    ai = ActionIdentify(rules, max_age=6, max_size=4, capacity=64, clock=lambda: clock["t"])
    for f in range(int(g["n_frames"])):
        clock["t"] = float(g["stamps"][f])
        got = ai.update(g[f"rows_{f}"])
        ref = [(int(t), int(c), rules[int(r)].name) for t, c, r in g[f"actions_{f}"]]
        assert got == ref, f"frame {f}: {got} vs {ref}"
    assert ai.update(None) is None                      # action_Identify.py:16-17


def test_action_identify_matches_oracle_second_sequence():
    from oracle.action_ref import ActionIdentifyRef
    from oracle.action_ref import action_sequence
    from yolo_deepsort_b200.action import ActionIdentify
    rules, spec = _rules()
    frames, stamps = action_sequence(seed=7, n_frames=90)
    clock = {"t": 0.0}
    ai = ActionIdentify(rules, max_age=5, max_size=3, capacity=32, clock=lambda: clock["t"])
    twin = ai.clone()                                   # a clone starts from an empty cache (action_Identify.py:12-13)
    ref = ActionIdentifyRef(spec, max_age=5, max_size=3)
    n = 0
    for rows, ts in zip(frames, stamps):
        clock["t"] = ts
        got = ai.update(rows)
        want = [(t, c, rules[r].name) for t, c, r in ref.update(rows, ts)]
        assert got == want
        n += len(got)
    assert n > 50
    clock["t"] = stamps[0]
    assert twin.update(frames[0]) == []                 # first sight of every id: orbits are created empty, no rule can fire


def test_action_identify_accepts_reference_style_rule_objects():
    """Rule objects of the reference's own package are recognised by class name and read for their parameters."""
    from yolo_deepsort_b200.action import ActionIdentify

    class Glide:                                        # shaped like action.actions.Glide
        def __init__(self, class_id, delta):
            self.name, self.class_id, self.delta = "glide", class_id, delta

    ai = ActionIdentify([Glide(0, (5.0, 3.0))], max_age=30, max_size=4, capacity=8, clock=lambda: 1.0)
    rows = lambda x: np.asarray([[x, 10, x + 20, 50, 1, 0]], np.int32)
    assert ai.update(rows(0)) == [] and ai.update(rows(10)) == [] and ai.update(rows(20)) == [(1, 0, "glide")]
    with pytest.raises(TypeError):
        ActionIdentify([object()], capacity=8)
