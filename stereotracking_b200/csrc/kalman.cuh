// Constant-velocity Kalman filter, float64, four tracks per warp.
// Replaces KalmanFilter.initiate/predict/project/update of
// mmtrack/models/motion/kalman_filter.py:60-189 (NumPy + scipy cho_factor /
// cho_solve on the host in the reference).  State (cx, cy, a, h, vcx, vcy, va,
// vh); F = [I I; 0 I], H = [I 0], applied structurally (their zero/one entries add
// exact zeros in the reference's dgemm).
#pragma once
#include "st_common.cuh"

namespace stkf {

constexpr double kWPos = 1.0 / 20;
constexpr double kWVel = 1.0 / 160;
constexpr int kScratchDoubles = 64 + 64 + 8;   // per-warp scratch of the camera-motion step: P, X, m

// The filter runs on groups of 8 lanes with the state in registers: lane gl (0..7) of a
// group owns row gl of P and element gl of the mean, so one warp advances four tracks
// at once.  No shared memory; rows are exchanged with width-8 shuffles, which all 32
// lanes execute (a group without a track runs on an identity state and stores
// nothing).  (The first version used one warp per track with the 8x8 matrices staged
// in shared memory; a sequence's tracks then took four times as many rounds.)
// Products follow np.linalg.multi_dot's association for these shapes; the Cholesky
// pivots are inverted once and multiplied (LAPACK's dpotrf / dtrsm do the same): a few
// ulp of float64 from the reference, far inside the 1e-5 gate (tests pin 1e-10).
struct Row8 {
    double p[8];      // row gl of the covariance
    double m;         // element gl of the mean
};

__device__ __forceinline__ double shfl8(double v, int src) { return __shfl_sync(0xffffffffu, v, src, 8); }

__device__ __forceinline__ Row8 load8(const double *mean, const double *cov, int gl, bool active) {
    Row8 r;
#pragma unroll
    for (int j = 0; j < 8; ++j) r.p[j] = active ? cov[gl * 8 + j] : (j == gl ? 1.0 : 0.0);
    r.m = active ? mean[gl] : (gl == 3 ? 1.0 : 0.0);
    return r;
}

__device__ __forceinline__ void store8(const Row8 &r, double *mean, double *cov, int gl, bool active) {
    if (!active) return;
#pragma unroll
    for (int j = 0; j < 8; ++j) cov[gl * 8 + j] = r.p[j];
    mean[gl] = r.m;
}

// KalmanFilter.initiate by one 8-lane group: lane gl writes mean[gl] and row gl of P
__device__ __forceinline__ void initiate8(double *mean, double *cov, const float *z_cxcyah, int gl) {
    const double h = (double)z_cxcyah[3];
    mean[gl] = gl < 4 ? (double)z_cxcyah[gl] : 0.0;
    double sd;
    if (gl == 2) sd = 1e-2;
    else if (gl == 6) sd = 1e-5;
    else if (gl < 4) sd = 2 * kWPos * h;
    else sd = 10 * kWVel * h;
#pragma unroll
    for (int j = 0; j < 8; ++j) cov[gl * 8 + j] = j == gl ? sd * sd : 0.0;
}

__device__ __forceinline__ void predict8(Row8 &r, int gl) {
    const double h = shfl8(r.m, 3);
    double x[8];                              // X = P F^T, own row
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = j < 4 ? r.p[j] + r.p[j + 4] : r.p[j];
    // P' = F X + Q: rows 0..3 add row i+4
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double lower = shfl8(x[j], (gl + 4) & 7);
        double v = gl < 4 ? x[j] + lower : x[j];
        if (j == gl) {
            double sd;
            if (gl == 2) sd = 1e-2;
            else if (gl == 6) sd = 1e-5;
            else if (gl < 4) sd = kWPos * h;
            else sd = kWVel * h;
            v += sd * sd;
        }
        r.p[j] = v;
    }
    const double mv = shfl8(r.m, (gl + 4) & 7);
    r.m = gl < 4 ? r.m + mv : r.m;
}

// z = measurement (cx, cy, a, h) promoted to float64, the same in all lanes of the group.
// nsa = 1 - bbox_score when the filter was built with use_nsa=True (kalman_filter.py:144-145
// scales every measurement std by it), else 1 (x * 1.0 is exact, so the default path is unchanged).
__device__ __forceinline__ void update8(Row8 &r, const double *z, int gl, double nsa = 1.0) {
    const double h = shfl8(r.m, 3);
    // S = H P H^T + R, replicated in every lane
    double S[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double v = shfl8(r.p[q], a);
            if (a == q) {
                const double sd = nsa * (a == 2 ? 1e-1 : kWPos * h);
                v += sd * sd;
            }
            S[a][q] = v;
        }
    // Cholesky S = L L^T (lower); pivots inverted once
    double L[4][4], inv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double acc = S[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) acc -= L[j][k] * L[j][k];
        L[j][j] = sqrt(acc);
        inv[j] = 1.0 / L[j][j];
#pragma unroll
        for (int i = j + 1; i < 4; ++i) {
            double a2 = S[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) a2 -= L[i][k] * L[j][k];
            L[i][j] = a2 * inv[j];
        }
    }
    // K row of this lane: S x = P[gl][0:4]
    double y[4], K[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double acc = r.p[a];
#pragma unroll
        for (int k = 0; k < a; ++k) acc -= L[a][k] * y[k];
        y[a] = acc * inv[a];
    }
#pragma unroll
    for (int a = 3; a >= 0; --a) {
        double acc = y[a];
#pragma unroll
        for (int k = a + 1; k < 4; ++k) acc -= L[k][a] * K[k];
        K[a] = acc * inv[a];
    }
    // column gl of S K^T
    double skt[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc += S[a][q] * K[q];
        skt[a] = acc;
    }
    // mean' = mean + innovation K^T
    double macc = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) macc += (z[a] - shfl8(r.m, a)) * K[a];
    // P' = P - K (S K^T)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) acc += K[a] * shfl8(skt[a], j);
        r.p[j] = r.p[j] - acc;
    }
    r.m = r.m + macc;
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
