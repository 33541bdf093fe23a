"""Mirror of the reference's `action` package on the device.

`ActionIdentify(actions, max_age=30, max_size=4)` with `.update(detections)` and `.clone()` as in action/action_Identify.py:4-47;
the rule objects `TakeOff / Landing / Glide / FastCrossing / BreakInto` take the constructor arguments of action/actions.py:23-150.
The orbit cache (action/orbit.py:5-26: a bounded deque of bottom-centre points and time stamps per track id) and the rules are
evaluated by libydst (`ydst_action_update`, csrc/action.cu) -- there is no host implementation of `confirm`.  Rule objects of the
reference's own package are accepted as well: they are recognised by class name and read for their parameters.
"""
import ctypes
import time

import numpy as np

from ._lib import check, lib, require_cuda


class Action:
    """A named rule (action/actions.py:6-10).  `confirm` lives on the device."""
    kind = -1

    def __init__(self, name):
        self.name = name

    def params(self):
        raise NotImplementedError


class _DeltaRule(Action):
    def __init__(self, name, class_id, delta):
        self.delta = delta
        self.class_id = class_id
        super().__init__(name)

    def params(self):
        return self.class_id, float(self.delta[0]), float(self.delta[1])


class TakeOff(_DeltaRule):
    kind = 0

    def __init__(self, class_id, delta):
        super().__init__("takeoff", class_id, delta)


class Landing(_DeltaRule):
    kind = 1

    def __init__(self, class_id, delta):
        super().__init__("landing", class_id, delta)


class Glide(_DeltaRule):
    kind = 2

    def __init__(self, class_id, delta):
        super().__init__("glide", class_id, delta)


class FastCrossing(Action):
    kind = 3

    def __init__(self, class_id, speed):
        super().__init__("fast_crossing")
        self.class_id = class_id
        self.speed = speed

    def params(self):
        return self.class_id, float(self.speed), 0.0


class BreakInto(Action):
    kind = 4

    def __init__(self, class_id, timeout):
        super().__init__("break_into")
        self.class_id = class_id
        self.timeout = timeout

    def params(self):
        return self.class_id, float(self.timeout), 0.0


_KIND_BY_NAME = {"TakeOff": 0, "Landing": 1, "Glide": 2, "FastCrossing": 3, "BreakInto": 4}


def _rule_of(action):
    """(kind, class_id, p0, p1) of one of our rule objects or of the reference's (action/actions.py), recognised by class name."""
    if isinstance(action, Action):
        return (action.kind,) + tuple(action.params())
    kind = _KIND_BY_NAME.get(type(action).__name__)
    if kind is None:
        raise TypeError(f"ActionIdentify: no device rule for {type(action).__name__}")
    if kind <= 2:
        return kind, action.class_id, float(action.delta[0]), float(action.delta[1])
    return kind, action.class_id, float(action.speed if kind == 3 else action.timeout), 0.0


class ActionIdentify:
    """Drop-in for action.action_Identify.ActionIdentify.  `capacity` bounds the live orbits and the rows per update."""

    def __init__(self, actions, max_age=30, max_size=4, capacity=4096, clock=time.time):
        require_cuda()
        self.actions, self.max_age, self.max_size, self.capacity, self._clock = list(actions), max_age, max_size, int(capacity), clock
        rules = [_rule_of(a) for a in self.actions]
        n = len(rules)
        kinds = (ctypes.c_int * max(n, 1))(*[r[0] for r in rules])
        cls = (ctypes.c_int * max(n, 1))(*[int(r[1]) for r in rules])
        p0 = (ctypes.c_double * max(n, 1))(*[r[2] for r in rules])
        p1 = (ctypes.c_double * max(n, 1))(*[r[3] for r in rules])
        self._h = ctypes.c_void_p()
        check(lib().ydst_action_create(int(max_age), int(max_size), kinds, cls, p0, p1, n, self.capacity, ctypes.byref(self._h)))
        self._out = np.zeros((self.capacity * max(n, 1), 3), np.int32)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().ydst_action_destroy(h)
            except Exception:
                pass

    def clone(self):
        return ActionIdentify(self.actions, self.max_age, self.max_size, self.capacity, self._clock)

    def update(self, detections):
        """-> [(track_id, class_id, action name), ...] in the reference's order; None when `detections` is None (no ageing then:
        action_Identify.py:16-17)."""
        if detections is None:
            return None
        rows = np.ascontiguousarray(np.asarray(detections, np.int32).reshape(-1, 6))
        n = ctypes.c_int()
        check(lib().ydst_action_update(self._h, rows.ctypes.data_as(ctypes.c_void_p), int(rows.shape[0]), float(self._clock()),
                                       self._out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(n), None))
        return [(int(t), int(c), self.actions[int(r)].name) for t, c, r in self._out[:n.value]]
