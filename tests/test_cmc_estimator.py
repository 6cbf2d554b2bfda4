"""Host-side camera-motion estimator behind cmc=dict(method='glme_affine') (CPU, OpenCV):
the part of the CMC row that stays on the host (trackers/utils.py:6-55); its application to
the Kalman states is the tracker kernel's (tests/test_gpu_parity.py, fixture 'cmc')."""
import numpy as np
import torch


def test_mesh_affine_estimator_recovers_a_known_shift():
    import cv2
    from stereotracking_b200.mmtrack_api import _mesh_affine
    rng = np.random.default_rng(3)
    base = rng.integers(0, 255, (60, 80, 3)).astype(np.uint8)
    big = cv2.resize(base, (640, 480), interpolation=cv2.INTER_CUBIC).astype(np.float32)
    prev = big[20:420, 20:580]
    cur = big[24:424, 14:574]                      # content moves by dx = +6, dy = -4
    to_t = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1)))[None]
    w = _mesh_affine(to_t(cur), to_t(prev), dict(img_shape=(400, 560)))
    assert w is not None and w.shape == (2, 3) and w.dtype == np.float32
    np.testing.assert_allclose(w[:, :2], np.eye(2), atol=0.03)
    np.testing.assert_allclose(w[:, 2], [6.0, -4.0], atol=1.5)


def test_mesh_affine_returns_none_without_texture():
    from stereotracking_b200.mmtrack_api import _mesh_affine
    flat = torch.zeros(1, 3, 64, 64)
    w = _mesh_affine(flat, flat, dict(img_shape=(64, 64)))
    assert w is None or np.allclose(w, [[1, 0, 0], [0, 1, 0]], atol=1e-3)
