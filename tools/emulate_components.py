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
    options = [[None] + [j for j in cols if cost[i, j] < limit] for i in rows]
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
    adm = cost < limit
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
