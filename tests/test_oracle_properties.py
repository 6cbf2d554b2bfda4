"""Property tests (Hypothesis, CPU) of the oracle pieces whose third-party originals are not in
the reference tree (NMS, the extended LAP, nearest resize) and of the host-side sharding: the
invariants any correct restatement must satisfy, independent of golden vectors."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import disparity, head, lapjv
from stereotracking_b200 import shard


def _boxes(rng, n):
    xy = rng.uniform(0, 200, (n, 2)).astype(np.float32)
    wh = rng.uniform(2, 80, (n, 2)).astype(np.float32)
    return np.concatenate([xy, xy + wh], 1)


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(0, 120), st.sampled_from([0.3, 0.5, 0.7]))
def test_greedy_nms_invariants(seed, n, thr):
    rng = np.random.default_rng(seed)
    b = _boxes(rng, n)                                   # already "sorted by score": position = rank
    keep = head.greedy_nms(b, thr)
    assert list(keep) == sorted(set(keep.tolist()))      # ascending positions = descending score
    kept = set(keep.tolist())
    for i in range(n):
        hit = [k for k in keep if k < i and head.iou_gt(b[k], b[i], thr)]
        assert (i in kept) == (len(hit) == 0)            # suppressed <=> overlapped by an earlier KEPT box
    again = head.greedy_nms(b[keep], thr)                # idempotent on its own output
    assert len(again) == len(keep)


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 12), st.integers(1, 12), st.sampled_from([0.5, 0.7, 0.9]))
def test_extended_lap_invariants(seed, R, C, limit):
    rng = np.random.default_rng(seed)
    cost = rng.uniform(0, 1.2, (R, C)).astype(np.float32)
    x, y = lapjv.lapjv(cost, limit)
    xp, yp = lapjv.lapjv_py(cost, limit)
    assert np.array_equal(x, xp) and np.array_equal(y, yp)             # C restatement == Python restatement
    for i, j in enumerate(x):
        if j >= 0:
            assert y[j] == i and cost[i, j] < limit                       # a matching of admissible pairs
    assert sum(j >= 0 for j in x) == sum(i >= 0 for i in y)
    # optimal: no matched pair can be dropped, no unmatched admissible pair added, at a gain
    total = sum(float(cost[i, j]) - limit for i, j in enumerate(x) if j >= 0)
    import scipy.optimize
    ext = lapjv.extend(cost, limit)
    r, c = scipy.optimize.linear_sum_assignment(ext)
    best = ext[r, c].sum() - limit / 2.0 * (R + C)                         # same objective, shifted
    assert abs(total - best) < 1e-9


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 400), st.integers(1, 400))
def test_nearest_index_is_monotone_and_in_range(dst_n, src_n):
    idx = disparity.nearest_index(dst_n, src_n)
    assert len(idx) == dst_n and idx.min() >= 0 and idx.max() <= src_n - 1
    assert np.all(np.diff(idx) >= 0)
    if dst_n == src_n:
        assert np.array_equal(idx, np.arange(src_n))


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 3000), st.integers(1, 16))
def test_shards_partition_the_sequences(n, world):
    if n < world:
        return
    blocks = [shard.shard_sequences(n, world, r) for r in range(world)]
    assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
    for (a, ca), (b, _) in zip(blocks, blocks[1:]):
        assert a + ca == b                                                # contiguous, disjoint, in rank order
    sizes = [c for _, c in blocks]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
