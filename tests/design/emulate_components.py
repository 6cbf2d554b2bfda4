"""Round-2 design study for K4's assignment (CPU restatement, not product code).

The association problem `lap.lapjv(cost, extend_cost=True, cost_limit=L)` minimises
    sum over matched pairs (cost[i,j] - L)            (+ a constant)
over partial matchings, so only ADMISSIBLE edges (cost < L) can be used, and rows in different
connected components of the admissible graph do not interact.  On the bench workload 98.8 % of
the problems have only components of at most 4 x 4 (profiles/r01_assign_components_c2.txt), which
can be solved by enumerating every partial matching of the component - at most 209 for 4 x 4 -
one candidate per lane, and the same enumeration certifies uniqueness (the runner-up is worse by
more than `eps`), which is what makes the result equal to ANY exact solver's, the reference's JV
included.  Components are independent, so a CTA can give each to another warp.

`component_assign` returns (row2col, col2row, ok); ok == False means "use the existing path"
(a component larger than `max_side`, a near tie inside a component, or a cost within `eps` of L).
"""
import itertools

import numpy as np


def label_components(adm):
    """Min-label propagation as the CTA would run it on the admissibility bitmasks:
    rows start with their own index, columns take the smallest adjacent row label and
    hand it back, until nothing changes.  Returns (row_label, col_label, rounds); -1 = no edge."""
    R, C = adm.shape
    big = R + C
    row = np.where(adm.any(1), np.arange(R), big)
    col = np.full(C, big)
    rounds = 0
    while True:
        rounds += 1
        new_col = np.where(adm.any(0), np.where(adm, row[:, None], big).min(0), big)
        new_row = np.where(adm.any(1), np.minimum(row, np.where(adm, new_col[None, :], big).min(1)), big)
        if np.array_equal(new_col, col) and np.array_equal(new_row, row):
            break
        row, col = new_row, new_col
    return np.where(row == big, -1, row), np.where(col == big, -1, col), rounds


def solve_component(cost, rows, cols, limit, eps):
    """Enumerate every partial matching of the component on admissible edges.
    Returns (assignment dict row->col, unique)."""
    best, second, arg = 0.0, np.inf, ()              # the empty matching
    options = [[None] + [j for j in cols if float(cost[i, j]) < limit] for i in rows]
    first = True
    for pick in itertools.product(*options):
        used = [j for j in pick if j is not None]
        if len(used) != len(set(used)):
            continue
        total = sum(float(cost[i, j]) - limit for i, j in zip(rows, pick) if j is not None)
        if first:
            best, arg, first = total, pick, False
            continue
        if total < best:
            best, second, arg = total, best, pick
        elif total < second:
            second = total
    return {i: j for i, j in zip(rows, arg) if j is not None}, second - best > eps


def component_assign(cost, limit, max_side=4, eps=1e-9):
    cost = np.asarray(cost, np.float32)
    R, C = cost.shape
    x, y = np.full(R, -1, np.int32), np.full(C, -1, np.int32)
    if R == 0 or C == 0:
        return x, y, True
    if (np.abs(cost.astype(np.float64) - limit) <= 2 * eps).any():
        return x, y, False
    adm = cost.astype(np.float64) < limit              # float64 like the kernel (NumPy 2 would compare in float32)
    row_label, col_label, _ = label_components(adm)
    for root in np.unique(row_label[row_label >= 0]):
        rows = np.nonzero(row_label == root)[0].tolist()
        cols = np.nonzero(col_label == root)[0].tolist()
        if len(rows) > max_side or len(cols) > max_side:
            return x, y, False
        match, unique = solve_component(cost, rows, cols, limit, eps)
        if not unique:
            return x, y, False
        for i, j in match.items():
            x[i], y[j] = j, i
    return x, y, True


def component_assign_lanes(cost, limit, max_side=4, eps=1e-9):
    """The same algorithm in the form one warp would run it (the transcription target for
    lapjv.cuh): lane t owns the t-th ACTIVE row (rows that are neither edgeless nor an isolated
    pair) and its admissibility bitmask; closure of the masks by shuffles; one leader per
    component; the component's partial matchings enumerated as base-(c+1) codes, code
    `lane + 32 k` on lane `lane`; best / runner-up reduced over the warp."""
    cost = np.asarray(cost, np.float32)
    R, C = cost.shape
    x, y = np.full(R, -1, np.int32), np.full(C, -1, np.int32)
    if R == 0 or C == 0:
        return x, y, True
    c64 = cost.astype(np.float64)
    if np.isnan(c64).any() or (np.abs(c64 - limit) <= 2 * eps).any():
        return x, y, False
    adm = c64 < limit                                  # float64 like the kernel
    row_cnt, col_cnt = adm.sum(1), adm.sum(0)
    active = []
    for i in range(R):
        if row_cnt[i] == 1 and col_cnt[np.nonzero(adm[i])[0][-1]] == 1:
            j = int(np.nonzero(adm[i])[0][-1])
            x[i], y[j] = j, i
        elif row_cnt[i] > 0:
            active.append(i)
    n = len(active)
    if n > 32:
        return x, y, False
    mask = [sum(1 << int(j) for j in np.nonzero(adm[i])[0]) for i in active]
    while True:                                   # closure: in-place, lane k broadcasts in turn
        changed = False
        for k in range(n):
            mk = mask[k]
            for t in range(n):
                if mask[t] & mk and mask[t] | mk != mask[t]:
                    mask[t] |= mk
                    changed = True
        if not changed:
            break
    done = [False] * n
    while not all(done):
        leader = done.index(False)
        lanes = [t for t in range(n) if mask[t] == mask[leader]]
        rows = [active[t] for t in lanes]
        cols = [j for j in range(C) if mask[leader] >> j & 1]
        for t in lanes:
            done[t] = True
        r, c = len(rows), len(cols)
        if r > max_side or c > max_side:
            return x, y, False
        base, n_codes = c + 1, (c + 1) ** r
        lane_best = [(np.inf, -1)] * 32
        lane_second = [np.inf] * 32
        for code in range(n_codes):
            lane = code % 32
            used, total, ok, rest = 0, 0.0, True, code
            for k in range(r):
                d, rest = rest % base, rest // base
                if d == 0:
                    continue
                j = cols[d - 1]
                if not adm[rows[k], j] or used >> j & 1:
                    ok = False
                    break
                used |= 1 << j
                total += c64[rows[k], j] - limit
            if not ok:
                continue
            if total < lane_best[lane][0]:
                lane_second[lane] = lane_best[lane][0]
                lane_best[lane] = (total, code)
            elif total < lane_second[lane]:
                lane_second[lane] = total
        win = min(range(32), key=lambda q: (lane_best[q][0], q))
        best, code = lane_best[win]
        second = min([lane_second[q] for q in range(32)] + [lane_best[q][0] for q in range(32) if q != win])
        if not second - best > eps:
            return x, y, False
        rest = code
        for k in range(r):
            d, rest = rest % base, rest // base
            if d:
                x[rows[k]], y[cols[d - 1]] = cols[d - 1], rows[k]
    return x, y, True
