"""Why the literal JV emulation runs on the WHOLE stage problem when the certified paths refuse: dropping the
rows and columns that have no admissible entry (they can never be matched) changes which of several equal-cost
optima the reference's solver returns.  Random tie-rich problems, the JV oracle on the full matrix against the
same oracle on the reduced one (order of the remaining rows / columns kept).  Analysis tool, not a test.

    python tests/design/jv_reduction_counterexample.py        # ~2.5 % of the tie-rich problems differ
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.lapjv import lapjv  # noqa: E402

L = 0.9


def reduced(cost):
    adm = cost.astype(np.float64) < L
    rr, cc = np.nonzero(adm.any(1))[0], np.nonzero(adm.any(0))[0]
    x, y = np.full(cost.shape[0], -1), np.full(cost.shape[1], -1)
    if len(rr) and len(cc):
        xs, _ = lapjv(np.ascontiguousarray(cost[np.ix_(rr, cc)]), L)
        for i, j in enumerate(xs):
            if j >= 0:
                x[rr[i]], y[cc[j]] = cc[j], rr[i]
    return x, y


def main(trials=3000):
    rng = np.random.default_rng(0)
    differ = 0
    for trial in range(trials):
        R, C = int(rng.integers(2, 25)), int(rng.integers(2, 25))
        kind = trial % 3
        cost = rng.uniform(0.95, 1.3, (R, C)).astype(np.float32)
        if kind == 0:          # costs on a coarse grid
            cost = (rng.integers(0, 8, (R, C)) * 0.15).astype(np.float32)
        elif kind == 1:        # the stress config's artefact: many entries exactly float32(0.9), just under the limit
            cost[rng.random((R, C)) < 0.2] = np.float32(0.9)
            m = rng.random((R, C)) < 0.1
            cost[m] = rng.uniform(0.1, 0.8, m.sum()).astype(np.float32)
        else:                  # duplicate tracks / detections
            cost = rng.uniform(0.0, 1.3, (R, C)).astype(np.float32)
            cost[rng.random((R, C)) < 0.5] = 1.0
            cost[1], cost[:, 1] = cost[0], cost[:, 0]
        x, y = lapjv(cost, L)
        xr, yr = reduced(cost)
        differ += not (np.array_equal(x, xr) and np.array_equal(y, yr))
    print(f'{trials} tie-rich problems: {differ} are resolved differently once the unmatched-for-sure rows / columns are dropped')


if __name__ == '__main__':
    main()
