"""CPU analysis for the round-2 plan of K4's assignment: on the bench workload (C2), how do the
rows of the stage-1 cost matrix split into connected components of the ADMISSIBLE graph
(edges with cost < cost_limit)?  Rows in different components cannot influence each other's
assignment, so each component is an independent small LAP that a separate warp could solve.
Runs the oracle (test infrastructure) - this is an analysis tool, not part of the product.

    python tests/design/assign_components.py [n_sequences] [n_frames] [C2|C3|C4]
"""
import collections
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import tracker as otracker  # noqa: E402
from oracle.model import StereoTrackOracle  # noqa: E402
from stereotracking_b200.synthetic import CONFIGS, SyntheticStereo  # noqa: E402


def components(adm):
    """Connected components of the bipartite graph rows x cols with edge matrix `adm`:
    list of (rows, cols)."""
    R, C = adm.shape
    seen_r, seen_c = np.zeros(R, bool), np.zeros(C, bool)
    out = []
    for r0 in range(R):
        if seen_r[r0] or not adm[r0].any():
            continue
        rows, cols, stack = [], [], [('r', r0)]
        seen_r[r0] = True
        while stack:
            kind, i = stack.pop()
            if kind == 'r':
                rows.append(i)
                for j in np.nonzero(adm[i] & ~seen_c)[0]:
                    seen_c[j] = True
                    stack.append(('c', j))
            else:
                cols.append(i)
                for k in np.nonzero(adm[:, i] & ~seen_r)[0]:
                    seen_r[k] = True
                    stack.append(('r', k))
        out.append((rows, cols))
    return out


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    F = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    cfg_name = sys.argv[3] if len(sys.argv) > 3 else 'C2'
    gen = SyntheticStereo(CONFIGS[cfg_name], S, F)
    fr = gen.frames(0, F)
    stats = collections.Counter()
    sizes = collections.Counter()
    worst = []
    orig = otracker.OCSortDisparityOracle._assign

    def spy(self, dists):
        R, C = dists.shape
        if dists.size:
            limit = 1 - self.match_iou_thr
            adm = dists.astype(np.float64) < limit
            comps = components(adm)
            nontrivial = [c for c in comps if not (len(c[0]) == 1 and len(c[1]) == 1)]
            stats['problems'] += 1
            stats['rows'] += R
            stats['rows_with_edges'] += int(adm.any(1).sum())
            stats['isolated_pairs'] += len(comps) - len(nontrivial)
            stats['nontrivial_components'] += len(nontrivial)
            stats['rows_in_nontrivial'] += sum(len(c[0]) for c in nontrivial)
            for c in nontrivial:
                sizes[(len(c[0]), len(c[1]))] += 1
            stats['problems_all_components_le_4x4'] += all(len(c[0]) <= 4 and len(c[1]) <= 4 for c in nontrivial)
            if nontrivial:
                worst.append((sum(len(c[0]) for c in nontrivial), max(len(c[0]) for c in nontrivial), len(nontrivial), R, C))
        return orig(self, dists)

    otracker.OCSortDisparityOracle._assign = spy
    for s in range(S):
        o = StereoTrackOracle()
        for f in range(F):
            o.step(f, fr['raw'][f, s].numpy(), [c[f, s].numpy() for c in fr['cls']],
                   [c[f, s].numpy() for c in fr['bbox']], [c[f, s].numpy() for c in fr['obj']])
    print(dict(stats))
    print('component sizes (rows, cols) -> count:', sorted(sizes.items(), key=lambda kv: -kv[1])[:12])
    worst.sort(reverse=True)
    print('worst problems (rows in non-trivial comps, largest comp rows, #comps, R, C):', worst[:8])
    if worst:
        serial = np.asarray([w[0] for w in worst], float)
        parallel = np.asarray([w[1] for w in worst], float)
        print('rows handled serially by one warp today: mean %.1f max %d; with a warp per component the critical '
              'path is the largest component: mean %.1f max %d' % (serial.mean(), serial.max(), parallel.mean(), parallel.max()))


if __name__ == '__main__':
    main()
