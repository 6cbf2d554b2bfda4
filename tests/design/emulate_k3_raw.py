"""Round-2 design study for K3 (CPU restatement, not product code): the per-box depth statistic
of OCSORT_Disparity.extract_depth (ocsort_disparity.py:136-175) computed from the RAW u16
disparity codes of the box instead of from the float32 depth map.

Depth is a strictly decreasing function of the code (tests/test_oracle_properties.py), so
  * the sorted list of valid depths is the multiset of valid codes in DESCENDING code order,
  * every order statistic is a rank in a 65 536-bin histogram of codes (two 8-bit radix
    passes on a GPU, or a small table of distinct codes),
  * the window mean is sum(depth(code) * multiplicity inside the window) / length,
and the depth map need not exist: pixels outside the H x W frame (the /32 pad) and code 65535
behave as disparity 0, i.e. the pad depth f*B/1e-6, which the `< 150` filter drops and the
2 x 2 corner means see as they see it in the map.
"""
import numpy as np

from oracle import boxdepth, disparity


def depth_table(baseline=0.25, focal_length=640):
    """float32 depth of every raw code with K1's arithmetic; 65535 -> disparity 0."""
    codes = np.arange(65536, dtype=np.uint16)[None, :]
    return disparity.preprocess(codes, size_divisor=1, baseline=baseline, focal_length=focal_length)['depth'][0]


def box_depth_from_raw(raw, box, Hp, Wp, table):
    """(d, scale) of one box like oracle.boxdepth.box_depth, from raw codes."""
    H, W = raw.shape
    bi = boxdepth.trunc_box(box)
    with np.errstate(over='ignore'):
        wbox = int(bi[2] - bi[0])
        e = [int(bi[0] + np.int64(2)), int(bi[1] + np.int64(2)), int(bi[2] - np.int64(2)), int(bi[3] - np.int64(2))]
    b = [int(v) for v in bi]

    def codes(ys, xs):
        """codes of the padded map restricted to Python slices ys, xs (pad = 65535)"""
        full = np.full((Hp, Wp), 65535, np.uint16)
        full[:H, :W] = raw
        return full[ys, xs]

    patch = codes(slice(b[1], b[3]), slice(b[0], b[2]))
    hist = np.bincount(patch.reshape(-1), minlength=65536)
    valid_code = (table < boxdepth.MAX_DEPTH) & (table > 0)
    valid_code[65535] = False
    hist = np.where(valid_code, hist, 0)
    n = int(hist.sum())
    if n < 1 or wbox > boxdepth.MAX_BOX_W:
        return -1, 1.0
    # ascending depth = descending code: below[c] = number of valid pixels with a LARGER code
    desc = hist[::-1]
    below_desc = np.concatenate([[0], np.cumsum(desc)[:-1]])

    def code_at(rank):
        k = int(np.searchsorted(below_desc + desc, rank, side='right'))
        return 65535 - k

    d_mid = table[code_at(n // 2)]

    def corner(ys, xs):
        v = table[codes(ys, xs)].reshape(-1)
        if v.size == 0:
            return np.float32(np.nan)
        acc = v[0]
        for t in v[1:]:
            acc = np.float32(acc + t)
        return np.float32(acc / np.float32(v.size))

    corners = np.asarray([corner(slice(b[1], e[1]), slice(b[0], e[0])), corner(slice(b[1], e[1]), slice(e[2], b[2])),
                          corner(slice(e[3], b[3]), slice(b[0], e[0])), corner(slice(e[3], b[3]), slice(e[2], b[2]))],
                         np.float32)
    with np.errstate(invalid='ignore'):
        k = int(np.count_nonzero(corners > d_mid))
    w_start = min(1 - k / 4, 0.4) * n
    w_end = w_start + 0.6 * n
    lo, hi, _ = slice(int(w_start), int(w_end)).indices(n)
    if hi <= lo:
        lo, hi, _ = slice(None, -1).indices(n)
    m = hi - lo
    if m <= 0:
        d = np.float32(np.nan)
    else:
        first = below_desc                                   # rank range of each code: [first, first + desc)
        inside = np.clip(np.minimum(first + desc, hi) - np.maximum(first, lo), 0, None)
        total = float(np.sum(table[::-1].astype(np.float64) * inside))
        d = np.float32(total / m)
    dd = float(np.float32(d * d))
    scale = max(min(dd, 3.0), 1.0)
    return d, scale
