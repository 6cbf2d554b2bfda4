"""Cross-checks of the oracle pieces whose reference code is third-party
(SURVEY.md section 8c): LAPJV vs SciPy, NMS vs torchvision, IoU vs
torchvision, nearest resize vs OpenCV, transcendental rounding vs torch."""
import numpy as np
import pytest
import scipy.optimize
import torch

from oracle import disparity, head
from oracle.lapjv import extend, lapjv, lapjv_py
from oracle.tracker import bbox_overlaps


def _lsa_cols(cost, limit):
    R, C = cost.shape
    ext = extend(cost, limit)
    r, c = scipy.optimize.linear_sum_assignment(ext)
    y = np.empty(R + C, np.int64)
    y[c] = r
    y = y[:C]
    y[y >= R] = -1
    return y


@pytest.mark.parametrize('seed', range(6))
def test_lapjv_matches_scipy_on_tie_free_costs(seed):
    rng = np.random.default_rng(seed)
    for _ in range(150):
        R, C = rng.integers(0, 14, 2)
        cost = rng.uniform(0.0, 1.3, (R, C)).astype(np.float32)
        x, y = lapjv(cost, 0.9)
        if R and C:
            np.testing.assert_array_equal(y, _lsa_cols(cost, 0.9))
        for j, i in enumerate(y):
            if i >= 0:
                assert x[i] == j and cost[i, j] < 0.9
        assert (x >= 0).sum() == (y >= 0).sum()


def test_lapjv_c_equals_python_including_ties():
    rng = np.random.default_rng(7)
    for _ in range(200):
        R, C = rng.integers(1, 9, 2)
        cost = rng.integers(0, 4, (R, C)).astype(np.float64) * 0.25   # many exact ties
        xc, yc = lapjv(cost, 0.9)
        xp, yp = lapjv_py(cost, 0.9)
        np.testing.assert_array_equal(xc, xp)
        np.testing.assert_array_equal(yc, yp)


def test_lapjv_large_and_empty():
    rng = np.random.default_rng(3)
    cost = rng.uniform(0, 1.2, (120, 90)).astype(np.float32)
    _, y = lapjv(cost, 0.9)
    np.testing.assert_array_equal(y, _lsa_cols(cost, 0.9))
    x, y = lapjv(np.zeros((0, 5)), 0.9)
    assert len(x) == 0 and (y == -1).all()
    x, y = lapjv(np.zeros((4, 0)), 0.9)
    assert len(y) == 0 and (x == -1).all()


def test_nms_matches_torchvision():
    tv = pytest.importorskip('torchvision')
    rng = np.random.default_rng(0)
    for n in [0, 1, 5, 200, 1000]:
        xy = rng.uniform(0, 300, (n, 2))
        wh = rng.uniform(5, 80, (n, 2))
        boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        scores = rng.permutation(n).astype(np.float32) / max(n, 1)
        order = np.argsort(-scores, kind='stable')
        keep = head.greedy_nms(boxes[order], 0.5)
        ref = tv.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), 0.5).numpy()
        np.testing.assert_array_equal(order[keep], ref)


def test_iou_matches_torchvision():
    tv = pytest.importorskip('torchvision')
    rng = np.random.default_rng(1)
    a = rng.uniform(0, 200, (40, 2)); b = rng.uniform(0, 200, (30, 2))
    A = np.concatenate([a, a + rng.uniform(1, 90, (40, 2))], 1).astype(np.float32)
    B = np.concatenate([b, b + rng.uniform(1, 90, (30, 2))], 1).astype(np.float32)
    ref = tv.ops.box_iou(torch.from_numpy(A), torch.from_numpy(B)).numpy()
    np.testing.assert_array_equal(bbox_overlaps(A, B), ref)


def test_nearest_resize_matches_opencv():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(2)
    for (h, w), scale in [((720, 1280), (720, 1280)), ((1080, 1920), (720, 1280)),
                          ((480, 640), (720, 1280)), ((375, 1242), (576, 1600)), ((97, 131), (50, 77))]:
        img = rng.uniform(0, 200, (h, w)).astype(np.float32)
        nh, nw = disparity.rescale_size(h, w, scale)
        ref = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_NEAREST)
        np.testing.assert_array_equal(disparity.nearest_resize(img, nh, nw), ref)


def test_bilinear_resize_matches_opencv():
    """bilinear_resize_u8 = cv2.resize(INTER_LINEAR) on 8-bit images, any size pair (including the exact
    2 x downscale, where OpenCV switches to the box average, and 1-pixel sides)."""
    cv2 = pytest.importorskip('cv2')
    rng = np.random.default_rng(5)
    shapes = [((1080, 1920), (720, 1280)), ((720, 1280), (1080, 1920)), ((480, 640), (240, 320)), ((100, 100), (50, 100)),
              ((9, 7), (1, 1)), ((1, 1), (4, 4)), ((2, 2), (1, 1)), ((300, 400), (100, 133)), ((64, 64), (63, 65))]
    shapes += [((int(rng.integers(1, 150)), int(rng.integers(1, 150))), (int(rng.integers(1, 200)), int(rng.integers(1, 200))))
               for _ in range(25)]
    for (h, w), (nh, nw) in shapes:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        np.testing.assert_array_equal(disparity.bilinear_resize_u8(img, nh, nw),
                                      cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR), err_msg=str(((h, w), (nh, nw))))


def test_preprocess_shapes_and_invalid():
    raw = np.full((720, 1280), 160, np.uint16)
    raw[:10] = 65535
    pre = disparity.preprocess(raw, img_scale=(720, 1280))
    assert pre['depth'].shape == (736, 1280)
    assert (pre['mask'][:10] == 0).all() and (pre['mask'][10:720] == 1).all()
    assert (pre['disp'][720:] == 0).all() and (pre['depth'][720:] > 1e8).all()
    assert abs(pre['depth'][100, 5] - 16.0) < 1e-4


def test_transcendentals_within_one_ulp_of_torch():
    rng = np.random.default_rng(4)
    x = rng.normal(0, 4, 200000).astype(np.float32)
    for ours, theirs in [(head.sigmoid_f32(x), torch.sigmoid(torch.from_numpy(x)).numpy()),
                         (head.exp_f32(x), torch.exp(torch.from_numpy(x)).numpy())]:
        ulp = np.spacing(np.abs(theirs))
        assert np.all(np.abs(ours.astype(np.float64) - theirs) <= 2 * ulp)


def test_head_postprocess_orders_and_clamps():
    rng = np.random.default_rng(9)
    lv = [(8, 10), (4, 5), (2, 3)]
    cls = [rng.normal(-2, 1, (1, h, w)).astype(np.float32) for h, w in lv]
    obj = [rng.normal(-1, 2, (1, h, w)).astype(np.float32) for h, w in lv]
    box = [rng.normal(0, 1, (4, h, w)).astype(np.float32) for h, w in lv]
    r = head.postprocess(cls, box, obj, (60, 75))
    assert (np.diff(r['scores']) <= 0).all() and (r['scores'] > 0.01).all()
    assert r['bboxes'][:, 0::2].max() <= 75 and r['bboxes'][:, 1::2].max() <= 60 and r['bboxes'].min() >= 0
    assert (np.diff(r['keep']) > 0).all()
