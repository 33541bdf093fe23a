"""Oracle: linear_sum_assignment wrapper around oracle/lsap.c (plus a pure-Python twin for tiny cases).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Algorithm notes and the scipy citation are in
oracle/lsap.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_ref", "liblsap_ref.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _LIB = ctypes.CDLL(path)
        _LIB.lsap_ref_solve.restype = ctypes.c_int
        _LIB.lsap_ref_solve.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


def lsap_ref(cost):
    """cost: (nr,nc) float32.  Returns (row_ind, col_ind) int64 arrays like scipy."""
    cost = np.ascontiguousarray(cost, dtype=np.float32)
    nr, nc = cost.shape
    k = min(nr, nc)
    rows = np.zeros(k, np.int32)
    cols = np.zeros(k, np.int32)
    if k:
        got = _lib().lsap_ref_solve(nr, nc, cost.ctypes.data, rows.ctypes.data, cols.ctypes.data)
        if got < 0:
            raise ValueError("cost matrix is infeasible")
    return rows.astype(np.int64), cols.astype(np.int64)


def lsap_py(cost):
    """Pure-Python statement of the same algorithm (small cases only; used to cross-check lsap.c)."""
    cost = np.asarray(cost, dtype=np.float64)
    transpose = cost.shape[1] < cost.shape[0]
    if transpose:
        cost = cost.T
    nr, nc = cost.shape
    u, v = [0.0] * nr, [0.0] * nc
    path, col4row, row4col = [-1] * nc, [-1] * nr, [-1] * nc
    inf = float("inf")
    for cur in range(nr):
        min_val, i = 0.0, cur
        remaining = list(range(nc - 1, -1, -1))
        spc = [inf] * nc
        SR, SC = [False] * nr, [False] * nc
        sink = -1
        while sink == -1:
            index, lowest = -1, inf
            SR[i] = True
            for it, j in enumerate(remaining):
                r = min_val + cost[i, j] - u[i] - v[j]
                if r < spc[j]:
                    path[j], spc[j] = i, r
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest, index = spc[j], it
            min_val = lowest
            if min_val == inf:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            remaining[index] = remaining[-1]
            remaining.pop()
        u[cur] += min_val
        for i in range(nr):
            if SR[i] and i != cur:
                u[i] += min_val - spc[col4row[i]]
        for j in range(nc):
            if SC[j]:
                v[j] -= min_val - spc[j]
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    if not transpose:
        return np.arange(nr, dtype=np.int64), np.asarray(col4row, np.int64)
    pairs = sorted((col4row[i], i) for i in range(nr))
    return np.asarray([p[0] for p in pairs], np.int64), np.asarray([p[1] for p in pairs], np.int64)
