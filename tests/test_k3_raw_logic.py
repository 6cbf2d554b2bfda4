"""Soundness of the round-2 plan to compute K3's per-box depth statistic from the raw u16 codes
(tests/design/emulate_k3_raw.py, a CPU restatement; not used by this round's kernels): identical -1 /
NaN cases and the same d and scale as the oracle's extract_depth on the float32 depth map, for
boxes inside, across and outside the frame and its /32 pad."""
import os
import sys

import numpy as np
from hypothesis import given, settings, strategies as st

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'design'))
from emulate_k3_raw import box_depth_from_raw, depth_table  # noqa: E402
from oracle import boxdepth, disparity  # noqa: E402

TABLE = depth_table()


@settings(max_examples=120, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(5, 70), st.integers(5, 70), st.sampled_from([4, 64, 4000]))
def test_statistic_from_raw_codes_equals_statistic_from_depth_map(seed, H, W, n_levels):
    rng = np.random.default_rng(seed)
    raw = (rng.integers(1, n_levels + 1, (H, W)) * (60000 // n_levels)).astype(np.uint16)   # few or many distinct codes
    raw[rng.random((H, W)) < 0.25] = 65535
    raw[rng.random((H, W)) < 0.05] = rng.integers(0, 18)                                     # codes deeper than 150 m
    depth = disparity.preprocess(raw)['depth']
    Hp, Wp = depth.shape
    x1, y1 = rng.uniform(-4, Wp), rng.uniform(-4, Hp)
    box = np.asarray([x1, y1, x1 + rng.uniform(0.2, Wp), y1 + rng.uniform(0.2, Hp)], np.float32)
    with np.errstate(all='ignore'):
        rd, rs = boxdepth.box_depth(depth, box)
        d, s = box_depth_from_raw(raw, box, Hp, Wp, TABLE)
    if isinstance(rd, int) or rd == -1:
        assert d == -1 and s == 1.0
    elif np.isnan(rd):
        assert np.isnan(d) and np.isnan(s) == np.isnan(rs)
    else:
        np.testing.assert_allclose(np.float32(d), np.float32(rd), rtol=1e-6, atol=0)
        np.testing.assert_allclose(np.float32(s), np.float32(rs), rtol=2e-6, atol=0)


def test_reference_fixture_from_raw_codes():
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, 'ref_boxdepth.npz'))
    Hp, Wp = g['depth_map'].shape
    for box, rd, rs in zip(g['boxes'], g['d'], g['scales']):
        with np.errstate(all='ignore'):
            d, s = box_depth_from_raw(g['raw'], box, Hp, Wp, TABLE)
        if np.isnan(rd):
            assert np.isnan(np.float32(d))
        else:
            np.testing.assert_allclose(np.float32(d), rd, rtol=1e-6, atol=0)
            np.testing.assert_allclose(np.float32(s), rs, rtol=2e-6, atol=0)
