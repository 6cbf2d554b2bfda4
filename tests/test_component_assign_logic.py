"""Soundness of the component-wise assignment planned for K4 (tests/design/emulate_components.py, a CPU
restatement; the kernel of this round does not use it yet): whenever it certifies a result, the
result equals the oracle's lap.lapjv restatement - on random sparse problems, on problems with
duplicated rows/columns (exact ties must be refused, not guessed), and on the stage-1/2/3 cost
matrices of a synthetic C2 sequence."""
import os
import sys

import numpy as np
from hypothesis import given, settings, strategies as st

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'design'))
from emulate_components import component_assign, component_assign_lanes, label_components  # noqa: E402
from oracle.lapjv import lapjv  # noqa: E402


def _sparse_cost(rng, R, C, density):
    cost = rng.uniform(0.9, 1.3, (R, C)).astype(np.float32)            # inadmissible background
    hit = rng.random((R, C)) < density
    cost[hit] = rng.uniform(0.0, 0.88, int(hit.sum())).astype(np.float32)
    return cost


@settings(max_examples=150, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 30), st.integers(1, 30), st.sampled_from([0.03, 0.08, 0.15]))
def test_certified_results_equal_the_jv_restatement(seed, R, C, density):
    rng = np.random.default_rng(seed)
    cost = _sparse_cost(rng, R, C, density)
    x, y, ok = component_assign(cost, 0.9)
    if ok:
        rx, ry = lapjv(cost, 0.9)
        assert np.array_equal(x, rx) and np.array_equal(y, ry)


@settings(max_examples=150, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 40), st.integers(1, 40), st.sampled_from([0.02, 0.05, 0.1]))
def test_lane_formulation_agrees_with_the_set_formulation(seed, R, C, density):
    """The warp-shaped restatement (bitmask closure, base-(c+1) codes, per-lane best / runner-up)
    certifies exactly the problems the straightforward one certifies (unless more than 32 rows are
    active), with the same result."""
    rng = np.random.default_rng(seed)
    cost = _sparse_cost(rng, R, C, density)
    if seed % 3 == 0 and R > 1:
        cost[1] = cost[0]                               # sprinkle exact ties
    x, y, ok = component_assign(cost, 0.9)
    lx, ly, lok = component_assign_lanes(cost, 0.9)
    if lok:
        assert ok and np.array_equal(x, lx) and np.array_equal(y, ly)
        rx, ry = lapjv(cost, 0.9)
        assert np.array_equal(lx, rx) and np.array_equal(ly, ry)
    elif ok:
        adm = cost < 0.9
        iso = [(adm[i].sum() == 1 and adm[:, np.nonzero(adm[i])[0][-1]].sum() == 1) for i in range(R)]
        assert sum(adm[i].any() and not iso[i] for i in range(R)) > 32


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(2, 12), st.integers(2, 12))
def test_exact_ties_are_refused(seed, R, C):
    rng = np.random.default_rng(seed)
    cost = _sparse_cost(rng, R, C, 0.3)
    cost[0] = np.float32(1.0)
    cost[0, 0] = np.float32(0.4)
    cost[1] = cost[0]                                   # two identical tracks that see one detection ...
    cost[2:, 0] = np.float32(1.0)                       # ... nobody else can take: either may get it
    x, y, ok = component_assign(cost, 0.9)
    assert not ok
    cost[1, 0] = np.float32(0.41)                       # tie broken: row 0 wins, certified
    x, y, ok = component_assign(cost, 0.9)               # (other components may still be large or tied)
    small = cost[:2, :1]
    assert component_assign(small, 0.9)[2] and component_assign(small, 0.9)[0].tolist() == [0, -1]
    if ok:
        assert x[0] == 0 and x[1] == -1
        rx, ry = lapjv(cost, 0.9)
        assert np.array_equal(x, rx) and np.array_equal(y, ry)


def test_labels_are_connected_components():
    rng = np.random.default_rng(4)
    for _ in range(50):
        adm = rng.random((12, 14)) < 0.12
        row, col, rounds = label_components(adm)
        assert rounds <= 12 + 14
        for i, j in zip(*np.nonzero(adm)):
            assert row[i] == col[j] >= 0                # an edge joins equal labels
        assert ((row >= 0) == adm.any(1)).all() and ((col >= 0) == adm.any(0)).all()
        for lab in np.unique(row[row >= 0]):
            assert row[lab] == lab                      # the label is the smallest row of the component


def test_bench_like_problems_are_mostly_certified():
    from oracle import tracker as otracker
    from oracle.model import StereoTrackOracle
    from stereotracking_b200.synthetic import SceneConfig, SyntheticStereo
    cfg = SceneConfig(360, 640, 10, 30, n_clutter=10)
    F = 25
    fr = SyntheticStereo(cfg, 1, F, seed=3).frames(0, F)
    seen = dict(n=0, ok=0)
    orig = otracker.OCSortDisparityOracle._assign

    def spy(self, dists):
        rx, ry = orig(self, dists)
        if dists.size:
            x, y, ok = component_assign(dists, 1 - self.match_iou_thr)
            seen['n'] += 1
            seen['ok'] += ok
            if ok:
                assert np.array_equal(x, rx) and np.array_equal(y, ry)
        return rx, ry

    otracker.OCSortDisparityOracle._assign = spy
    try:
        o = StereoTrackOracle()
        for f in range(F):
            o.step(f, fr['raw'][f, 0].numpy(), [c[f, 0].numpy() for c in fr['cls']],
                   [c[f, 0].numpy() for c in fr['bbox']], [c[f, 0].numpy() for c in fr['obj']])
    finally:
        otracker.OCSortDisparityOracle._assign = orig
    assert seen['n'] > 30 and seen['ok'] >= 0.75 * seen['n']      # a crowded scene: 30 objects in 640 x 360
