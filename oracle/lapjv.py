"""Row A6/A7 (assignment): ``lap.lapjv(cost, extend_cost=True, cost_limit=L)``.

Third-party (gatagat/lap 0.4.x; requirements/runtime.txt:2, unpinned), used at
mmtrack/models/trackers/ocsort_tracker_disparity.py:260-261 and :312-313.
``lapjv`` below calls the C restatement in lapjv.c (built by oracle/Makefile);
``lapjv_py`` is the same algorithm in pure Python for small cross-checks.
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
        so = os.path.join(_HERE, '_build', 'liboracle.so')
        if not os.path.exists(so):
            subprocess.check_call(['make', '-C', _HERE, '-s'])
        lib = ctypes.CDLL(so)
        lib.oracle_lapjv_extended.argtypes = [
            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_double,
            ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_lapjv_extended.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def lapjv(cost, cost_limit):
    """Returns (x [R] i32 row->col, y [C] i32 col->row), -1 = unmatched."""
    c = np.ascontiguousarray(cost, dtype=np.float64)
    R, C = c.shape
    x = np.full(R, -1, np.int32)
    y = np.full(C, -1, np.int32)
    if c.size > 0:
        _lib().oracle_lapjv_extended(R, C, c.ctypes.data, float(cost_limit),
                                     x.ctypes.data, y.ctypes.data)
    return x, y


def extend(cost, cost_limit):
    c = np.ascontiguousarray(cost, dtype=np.float64)
    R, C = c.shape
    ext = np.full((R + C, R + C), cost_limit / 2.0)
    ext[R:, C:] = 0
    ext[:R, :C] = c
    return ext


def _solve_square_py(c):
    n = c.shape[0]
    BIG = 1.0e300
    x = [-1] * n
    y = [0] * n
    v = [BIG] * n
    for i in range(n):
        for j in range(n):
            if c[i, j] < v[j]:
                v[j] = c[i, j]
                y[j] = i
    unique = [True] * n
    for j in range(n - 1, -1, -1):
        i = y[j]
        if x[i] < 0:
            x[i] = j
        else:
            unique[i] = False
            y[j] = -1
    free = []
    for i in range(n):
        if x[i] < 0:
            free.append(i)
        elif unique[i]:
            j1 = x[i]
            m = BIG
            for j2 in range(n):
                if j2 != j1:
                    m = min(m, c[i, j2] - v[j2])
            v[j1] -= m
    for _ in range(2):
        if not free:
            break
        nfree = len(free)
        free = free + [0] * (n - nfree)
        cur = nxt = visits = 0
        while cur < nfree:
            visits += 1
            fr = free[cur]
            cur += 1
            j1, u1, j2, u2 = 0, c[fr, 0] - v[0], -1, BIG
            for j in range(1, n):
                r = c[fr, j] - v[j]
                if r < u2:
                    if r >= u1:
                        u2, j2 = r, j
                    else:
                        u2, u1, j2, j1 = u1, r, j1, j
            i0 = y[j1]
            v1n = v[j1] - (u2 - u1)
            lowers = v1n < v[j1]
            if visits < cur * n:
                if lowers:
                    v[j1] = v1n
                elif i0 >= 0 and j2 >= 0:
                    j1 = j2
                    i0 = y[j2]
                if i0 >= 0:
                    if lowers:
                        cur -= 1
                        free[cur] = i0
                    else:
                        free[nxt] = i0
                        nxt += 1
            elif i0 >= 0:
                free[nxt] = i0
                nxt += 1
            x[fr] = j1
            y[j1] = fr
        free = free[:nxt]
    for start in free:
        cols = list(range(n))
        pred = [start] * n
        d = [c[start, j] - v[j] for j in range(n)]
        lo = hi = n_ready = 0
        final_j = -1
        while final_j == -1:
            if lo == hi:
                n_ready = lo
                hi = lo + 1
                mind = d[cols[lo]]
                for k in range(hi, n):
                    j = cols[k]
                    if d[j] <= mind:
                        if d[j] < mind:
                            hi = lo
                            mind = d[j]
                        cols[k] = cols[hi]
                        cols[hi] = j
                        hi += 1
                for k in range(lo, hi):
                    if y[cols[k]] < 0:
                        final_j = cols[k]
            if final_j == -1:
                scan_start = lo
                while lo != hi and final_j == -1:
                    j = cols[lo]
                    lo += 1
                    i = y[j]
                    mind = d[j]
                    h = c[i, j] - v[j] - mind
                    for k in range(hi, n):
                        j = cols[k]
                        r = c[i, j] - v[j] - h
                        if r < d[j]:
                            d[j] = r
                            pred[j] = i
                            if r == mind:
                                if y[j] < 0:
                                    final_j = j
                                    break
                                cols[k] = cols[hi]
                                cols[hi] = j
                                hi += 1
                if final_j != -1:
                    # an early return leaves the caller's lo at the value it
                    # had when the scan started
                    lo = scan_start
        mind = d[cols[lo]]
        for k in range(n_ready):
            j = cols[k]
            v[j] += d[j] - mind
        i, j = -1, final_j
        while i != start:
            i = pred[j]
            y[j] = i
            j, x[i] = x[i], j
    return x, y


def lapjv_py(cost, cost_limit):
    c = np.ascontiguousarray(cost, dtype=np.float64)
    R, C = c.shape
    if c.size == 0:
        return np.full(R, -1, np.int32), np.full(C, -1, np.int32)
    xe, ye = _solve_square_py(extend(c, cost_limit))
    x = np.asarray(xe[:R], np.int32)
    y = np.asarray(ye[:C], np.int32)
    x[x >= C] = -1
    y[y >= R] = -1
    return x, y
