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


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(8, 40), st.integers(8, 40))
def test_depth_reducers_are_order_statistics_of_the_valid_patch(seed, H, W):
    """extract_depth and its four ablation reducers: -1 exactly when the patch has no valid depth
    (or the box is wider than 800 px); otherwise every statistic but `center` lies between the
    smallest and largest valid depth of the patch, min <= median <= max, and scale in [1, 3]."""
    from oracle import boxdepth
    rng = np.random.default_rng(seed)
    depth = rng.uniform(0.5, 220, (H, W)).astype(np.float32)
    depth[rng.random((H, W)) < 0.2] = np.float32(1.6e8)          # invalid disparity -> huge depth
    x1, y1 = rng.uniform(-3, W - 2), rng.uniform(-3, H - 2)
    box = np.asarray([x1, y1, x1 + rng.uniform(0.5, W), y1 + rng.uniform(0.5, H)], np.float32)
    b = boxdepth.trunc_box(box)
    patch = depth[b[1]:b[3], b[0]:b[2]]
    valid = patch[(patch < 150) & (patch > 0)]
    with np.errstate(all='ignore'):
        results = {'window': boxdepth.box_depth(depth, box)}
        for mode in boxdepth.ABLATION_MODES:
            try:
                results[mode] = boxdepth.box_depth_ablation(depth, box, mode)
            except IndexError:
                assert mode == 'center'
    for mode, (d, s) in results.items():
        if valid.size == 0:
            assert d == -1 and s == 1.0
            continue
        assert d != -1
        if np.isnan(np.float32(d)):
            assert valid.size == 1 and mode in ('window', 'truncated_mean')       # sorted[:-1] of one element
            continue
        assert 1.0 <= s <= 3.0
        if mode != 'center':
            assert valid.min() <= d <= valid.max()
    if valid.size:
        assert results['median'][0] == np.float32(np.median(valid))


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 70), st.integers(1, 70), st.booleans())
def test_pack_image_layout(seed, H, W, swap):
    """K1b's oracle: padded to /32, the image region is the (optionally channel-swapped) transpose,
    the pad is the pad value, normalisation is affine per channel."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
    out = disparity.pack_image(img, pad_val=(7, 8, 9), swap_rb=swap)
    assert out.dtype == np.float32 and out.shape[1] % 32 == 0 and out.shape[2] % 32 == 0
    assert out.shape[1] - H < 32 and out.shape[2] - W < 32
    src = img[..., ::-1] if swap else img
    assert np.array_equal(out[:, :H, :W], src.transpose(2, 0, 1).astype(np.float32))
    pad = (9, 8, 7) if swap else (7, 8, 9)
    for c in range(3):
        assert (out[c, H:, :] == pad[c]).all() and (out[c, :, W:] == pad[c]).all()
    mean, std = (10.0, 20.0, 30.0), (2.0, 4.0, 8.0)
    norm = disparity.pack_image(img, pad_val=(7, 8, 9), mean=mean, std=std, swap_rb=swap)
    for c in range(3):
        assert np.array_equal(norm[c], (out[c] - np.float32(mean[c])) / np.float32(std[c]))


def test_depth_is_strictly_decreasing_in_the_raw_code():
    """Basis of the round-2 plan to take K3's order statistics on the 16-bit raw codes: over all
    valid codes the float32 depth Z = f*B / (raw/16 + 1e-6) is strictly decreasing, so sorting by
    depth is sorting by code in reverse, and equal depths are equal codes."""
    raw = np.arange(1, 65535, dtype=np.uint16)[None, :]
    depth = disparity.preprocess(raw, size_divisor=1)['depth'][0, :65534]
    assert depth.dtype == np.float32 and np.all(np.diff(depth) < 0)
    valid = (depth < 150) & (depth > 0)
    assert valid.sum() == 65534 - 17 and not valid[:17].any()       # raw codes 1..17 are deeper than 150 m
