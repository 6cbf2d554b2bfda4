"""Row A5: DeepSORT-style constant-velocity Kalman filter, float64.

Follows mmtrack/models/motion/kalman_filter.py:38-58 (model matrices),
:60-85 (initiate), :87-119 (predict), :121-152 (project), :154-189 (update).
State (cx, cy, a, h, vcx, vcy, va, vh).  All arithmetic in float64 as under
the reference's pinned NumPy 1.x (float32 measurements promote exactly).
"""
import numpy as np
import scipy.linalg

NDIM = 4
W_POS = 1.0 / 20
W_VEL = 1.0 / 160
F = np.eye(2 * NDIM)
F[:NDIM, NDIM:] = np.eye(NDIM)
H = np.eye(NDIM, 2 * NDIM)


def initiate(z):
    z = np.asarray(z, np.float64)
    h = z[3]
    mean = np.concatenate([z, np.zeros(NDIM)])
    std = np.array([2 * W_POS * h, 2 * W_POS * h, 1e-2, 2 * W_POS * h,
                    10 * W_VEL * h, 10 * W_VEL * h, 1e-5, 10 * W_VEL * h])
    return mean, np.diag(std * std)


def predict(mean, cov):
    h = mean[3]
    std = np.array([W_POS * h, W_POS * h, 1e-2, W_POS * h,
                    W_VEL * h, W_VEL * h, 1e-5, W_VEL * h])
    return F @ mean, np.linalg.multi_dot((F, cov, F.T)) + np.diag(std * std)


def project(mean, cov, bbox_score=0., use_nsa=False):
    h = mean[3]
    std = [W_POS * h, W_POS * h, 1e-1, W_POS * h]
    if use_nsa:                                   # kalman_filter.py:144-145
        std = [(1 - bbox_score) * x for x in std]
    std = np.array(std)
    return H @ mean, np.linalg.multi_dot((H, cov, H.T)) + np.diag(std * std)


def update(mean, cov, z, bbox_score=0., use_nsa=False):
    pm, pc = project(mean, cov, bbox_score, use_nsa)
    chol = scipy.linalg.cho_factor(pc, lower=True, check_finite=False)
    gain = scipy.linalg.cho_solve(chol, (cov @ H.T).T, check_finite=False).T
    innov = np.asarray(z, np.float64) - pm
    return mean + innov @ gain.T, cov - np.linalg.multi_dot((gain, pc, gain.T))


def gating_distance(mean, cov, measurements, only_position=False):
    """kalman_filter.py:190-234: squared Mahalanobis distance of each measurement (N,4)."""
    pm, pc = project(mean, cov)
    z = np.asarray(measurements, np.float64)
    if only_position:
        pm, pc, z = pm[:2], pc[:2, :2], z[:, :2]
    chol = np.linalg.cholesky(pc)
    y = scipy.linalg.solve_triangular(chol, (z - pm).T, lower=True, check_finite=False)
    return np.sum(y * y, axis=0)
