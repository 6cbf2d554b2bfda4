"""Rows A3/A4: per-box depth statistic and depth-driven box scaling.

Follows mmtrack/models/mot/ocsort_disparity.py:113-130 (bbox_postp_depth),
:136-175 (extract_depth) and mmtrack/models/trackers/utils.py:58-73
(scale_bbox).  Scalar typing follows the reference's pinned NumPy 1.x: the
window bounds are Python floats (float64), d is float32, d*d is float32.
"""
import math

import numpy as np

MAX_DEPTH = 150      # ocsort_disparity.py:148
MAX_BOX_W = 800      # ocsort_disparity.py:150


def trunc_box(box):
    """ocsort_disparity.py:141 ``box.astype(np.int)``: C truncation toward
    zero; non-finite / out-of-range values give INT64_MIN as on x86-64."""
    b = np.asarray(box, np.float32)
    out = np.empty(4, np.int64)
    for k in range(4):
        v = float(b[k])
        if not math.isfinite(v) or abs(v) >= 2.0 ** 63:
            out[k] = np.iinfo(np.int64).min
        else:
            out[k] = int(v)
    return out


def _mean_f32(a):
    """np.mean of a float32 array (float32 accumulation, NaN when empty)."""
    a = np.asarray(a, np.float32)
    if a.size == 0:
        return np.float32(np.nan)
    return np.float32(np.mean(a))


def box_depth(depth, box):
    """One box of extract_depth (:140-173).  Returns (d, scale): d is a
    float32 (possibly NaN) or the int -1; scale a Python float."""
    bi = trunc_box(box)                       # np.int64: arithmetic wraps like the reference's
    with np.errstate(over='ignore'):
        w = int(bi[2] - bi[0])
        e = [int(bi[0] + np.int64(2)), int(bi[1] + np.int64(2)),
             int(bi[2] - np.int64(2)), int(bi[3] - np.int64(2))]
    b = [int(v) for v in bi]
    patch = depth[b[1]:b[3], b[0]:b[2]]
    vals = patch[(patch < MAX_DEPTH) & (patch > 0)]
    n = int(vals.size)
    if n < 1 or w > MAX_BOX_W:
        return -1, 1.0
    srt = np.sort(vals, axis=None)
    d_mid = srt[n // 2]
    corners = np.asarray([
        _mean_f32(depth[b[1]:e[1], b[0]:e[0]]),
        _mean_f32(depth[b[1]:e[1], e[2]:b[2]]),
        _mean_f32(depth[e[3]:b[3], b[0]:e[0]]),
        _mean_f32(depth[e[3]:b[3], e[2]:b[2]])], np.float32)
    with np.errstate(invalid='ignore'):
        k = int(np.count_nonzero(corners > d_mid))
    w_start = min(1 - k / 4, 0.4) * n
    w_end = w_start + 0.6 * n
    seg = srt[int(w_start):int(w_end)]
    if seg.size == 0:
        seg = srt[:-1]
    d = _mean_f32(seg)
    dd = float(np.float32(d * d))
    # Python min/max keep the first argument when a comparison is False, so a
    # NaN d gives a NaN scale (:171-172).
    scale = min(dd, 3.0)
    scale = max(scale, 1.0)
    return d, scale


def extract_depth(depth, bboxes):
    """ocsort_disparity.py:136-175.  depth [H',W'] f32, bboxes [D,4] f32 xyxy.
    Returns (d_values f32 [D] with -1 / NaN as in the reference, scales f32)."""
    depth = np.asarray(depth, np.float32)
    d_vals, scales = [], []
    for box in np.asarray(bboxes, np.float32).reshape(-1, 4):
        d, s = box_depth(depth, box)
        d_vals.append(d)
        scales.append(s)
    return (np.asarray(d_vals, np.float32).reshape(-1),
            np.asarray(scales, np.float32).reshape(-1))


# --- ablation reducers (mmtrack/models/mot/depth_extraction_comparison.py) -----------------
ABLATION_MODES = ('truncated_mean', 'mean', 'median', 'center')
ABLATION_SCALE_DIV = 400     # depth_extraction_comparison.py:41,:75,:110,:145


def box_depth_ablation(depth, box, mode):
    """One box of the decorators truncated_mean (:5-46), mean (:49-80), median (:83-115)
    and center (:117-150).  Same patch, validity filter and ``-1`` rule as extract_depth;
    the statistic differs and the scale is ``clip(d*d/400, 1, 3)``.  ``center`` reads the
    unfiltered map at the integer box centre with NumPy indexing (negative indices wrap,
    out of range raises IndexError as in the reference)."""
    bi = trunc_box(box)
    with np.errstate(over='ignore'):
        w = int(bi[2] - bi[0])
        cx = int((bi[0] + bi[2]) // 2)
        cy = int((bi[1] + bi[3]) // 2)
    b = [int(v) for v in bi]
    patch = depth[b[1]:b[3], b[0]:b[2]]
    vals = patch[(patch < MAX_DEPTH) & (patch > 0)]
    n = int(vals.size)
    if n < 1 or w > MAX_BOX_W:
        return -1, 1.0
    if mode == 'truncated_mean':
        srt = np.sort(vals, axis=None)
        seg = srt[int(0.1 * n):int((1 - 0.1) * n)]
        if seg.size == 0:
            seg = srt[:-1]
        d = _mean_f32(seg)
    elif mode == 'mean':
        d = _mean_f32(vals)
    elif mode == 'median':
        d = np.float32(np.median(vals))
    elif mode == 'center':
        d = depth[cy, cx]
    else:
        raise ValueError(f'unknown depth reducer {mode!r}')
    with np.errstate(over='ignore'):
        dd = float(np.float32(d * d)) / ABLATION_SCALE_DIV
    scale = min(dd, 3.0)
    scale = max(scale, 1.0)
    return d, scale


def extract_depth_ablation(depth, bboxes, mode):
    """The decorated ``extract_depth`` over all boxes: (d_values f32, scales f32)."""
    depth = np.asarray(depth, np.float32)
    d_vals, scales = [], []
    for box in np.asarray(bboxes, np.float32).reshape(-1, 4):
        d, s = box_depth_ablation(depth, box, mode)
        d_vals.append(d)
        scales.append(s)
    return (np.asarray(d_vals, np.float32).reshape(-1),
            np.asarray(scales, np.float32).reshape(-1))


def scale_bbox(bboxes, scales):
    """trackers/utils.py:58-73 in float32."""
    b = np.asarray(bboxes, np.float32).reshape(-1, 4)
    s = np.asarray(scales, np.float32)
    two = np.float32(2)
    cx = (b[:, 0] + b[:, 2]) / two
    cy = (b[:, 1] + b[:, 3]) / two
    w = (b[:, 2] - b[:, 0]) * s
    h = (b[:, 3] - b[:, 1]) * s
    return np.stack([cx - w / two, cy - h / two, cx + w / two, cy + h / two],
                    -1).astype(np.float32)


def bbox_postp_depth(depth, bboxes):
    """ocsort_disparity.py:113-130: returns (inflated boxes, scales, d)."""
    d, s = extract_depth(depth, bboxes)
    return scale_bbox(bboxes, s), s, d
