// Warp-cooperative constant-velocity Kalman filter, float64, one warp per track.
// Replaces KalmanFilter.initiate/predict/project/update of
// mmtrack/models/motion/kalman_filter.py:60-189 (NumPy + scipy cho_factor /
// cho_solve on the host in the reference).  State (cx, cy, a, h, vcx, vcy, va,
// vh); F = [I I; 0 I], H = [I 0].  Products follow np.linalg.multi_dot's
// association A(BC) for these shapes; F and H are applied structurally (their
// zero/one entries add exact zeros in the reference's dgemm).
#pragma once
#include "st_common.cuh"

namespace stkf {

constexpr double kWPos = 1.0 / 20;
constexpr double kWVel = 1.0 / 160;
constexpr int kScratchDoubles = 64 + 64 + 8 + 32 + 32 + 16;   // P, X, m, K, SKt, S

// mean[8], cov[64] in global memory; ws = this warp's shared scratch
__device__ inline void initiate(double *mean, double *cov, const float *z_cxcyah) {
    const int lane = threadIdx.x & 31;
    const double h = (double)z_cxcyah[3];
    if (lane < 8) mean[lane] = lane < 4 ? (double)z_cxcyah[lane] : 0.0;
    for (int e = lane; e < 64; e += 32) {
        const int i = e >> 3, j = e & 7;
        double v = 0.0;
        if (i == j) {
            double sd;
            if (i == 2) sd = 1e-2;
            else if (i == 6) sd = 1e-5;
            else if (i < 4) sd = 2 * kWPos * h;
            else sd = 10 * kWVel * h;
            v = sd * sd;
        }
        cov[e] = v;
    }
    __syncwarp();
}

__device__ inline void predict(double *mean, double *cov, double *ws) {
    const int lane = threadIdx.x & 31;
    double *P = ws, *X = ws + 64, *m = ws + 128;
    for (int e = lane; e < 64; e += 32) P[e] = cov[e];
    if (lane < 8) m[lane] = mean[lane];
    __syncwarp();
    const double h = m[3];
    // X = P F^T
    for (int e = lane; e < 64; e += 32) {
        const int i = e >> 3, j = e & 7;
        X[e] = j < 4 ? P[i * 8 + j] + P[i * 8 + j + 4] : P[e];
    }
    __syncwarp();
    // P' = F X + Q
    for (int e = lane; e < 64; e += 32) {
        const int i = e >> 3, j = e & 7;
        double v = i < 4 ? X[e] + X[(i + 4) * 8 + j] : X[e];
        if (i == j) {
            double sd;
            if (i == 2) sd = 1e-2;
            else if (i == 6) sd = 1e-5;
            else if (i < 4) sd = kWPos * h;
            else sd = kWVel * h;
            v += sd * sd;
        }
        cov[e] = v;
    }
    if (lane < 8) mean[lane] = lane < 4 ? m[lane] + m[lane + 4] : m[lane];
    __syncwarp();
}

// z = measurement (cx, cy, a, h) already promoted to float64
__device__ inline void update(double *mean, double *cov, const double *z, double *ws) {
    const int lane = threadIdx.x & 31;
    double *P = ws, *m = ws + 128, *K = ws + 136, *SKt = ws + 168, *S = ws + 200;
    for (int e = lane; e < 64; e += 32) P[e] = cov[e];
    if (lane < 8) m[lane] = mean[lane];
    __syncwarp();
    const double h = m[3];
    // S = H P H^T + R (project)
    if (lane < 16) {
        const int r = lane >> 2, q = lane & 3;
        double v = P[r * 8 + q];
        if (r == q) {
            const double sd = r == 2 ? 1e-1 : kWPos * h;
            v += sd * sd;
        }
        S[lane] = v;
    }
    __syncwarp();
    // Cholesky S = L L^T (lower), every lane redundantly in registers
    double L[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double acc = S[j * 4 + j];
#pragma unroll
        for (int k = 0; k < j; ++k) acc -= L[j][k] * L[j][k];
        L[j][j] = sqrt(acc);
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            double a2 = S[i * 4 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) a2 -= L[i][k] * L[j][k];
            L[i][j] = a2 / L[j][j];
        }
    }
    // K = (S^-1 (P H^T)^T)^T : lane c < 8 solves S x = P[c][0:4]
    if (lane < 8) {
        double y[4], x[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            double acc = P[lane * 8 + r];
#pragma unroll
            for (int k = 0; k < r; ++k) acc -= L[r][k] * y[k];
            y[r] = acc / L[r][r];
        }
#pragma unroll
        for (int r = 3; r >= 0; --r) {
            double acc = y[r];
#pragma unroll
            for (int k = r + 1; k < 4; ++k) acc -= L[k][r] * x[k];
            x[r] = acc / L[r][r];
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) K[lane * 4 + r] = x[r];
    }
    __syncwarp();
    // SKt = S K^T (4x8)
    {
        const int r = lane >> 3, c = lane & 7;
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc += S[r * 4 + q] * K[c * 4 + q];
        SKt[r * 8 + c] = acc;
    }
    __syncwarp();
    // mean' = mean + innovation K^T ; P' = P - K (S K^T)
    if (lane < 8) {
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) acc += (z[r] - m[r]) * K[lane * 4 + r];
        mean[lane] = m[lane] + acc;
    }
    for (int e = lane; e < 64; e += 32) {
        const int i = e >> 3, j = e & 7;
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) acc += K[i * 4 + r] * SKt[r * 8 + j];
        cov[e] = P[e] - acc;
    }
    __syncwarp();
}

// bbox_xyxy_to_cxcyah (mmtrack/structures/bbox/transforms.py:72-86), float32
__device__ __forceinline__ void xyxy_to_cxcyah(const float *b, float *o) {
    const float w = __fsub_rn(b[2], b[0]), h = __fsub_rn(b[3], b[1]);
    o[0] = __fmul_rn(__fadd_rn(b[2], b[0]), 0.5f);
    o[1] = __fmul_rn(__fadd_rn(b[3], b[1]), 0.5f);
    o[2] = __fdiv_rn(w, h);
    o[3] = h;
}

}  // namespace stkf
