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

// ---------------------------------------------------------------------------------
// The same filter on groups of 8 lanes, state in registers: lane gl (0..7) of a group
// owns row gl of P and element gl of the mean, so one warp advances four tracks at
// once and a sequence's tracks are done in a quarter of the rounds.  No shared
// memory; rows are exchanged with width-8 shuffles, which all 32 lanes execute (a
// group without a track runs on zeros and stores nothing).  Same operation order
// as predict() / update() above, except that the Cholesky pivots are inverted once
// and multiplied (LAPACK's dpotrf / dtrsm do the same); the difference is a few
// ulp of float64, far inside the 1e-5 gate (tests pin 1e-10 against the reference).
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

// z = measurement (cx, cy, a, h) promoted to float64, the same in all lanes of the group
__device__ __forceinline__ void update8(Row8 &r, const double *z, int gl) {
    const double h = shfl8(r.m, 3);
    // S = H P H^T + R, replicated in every lane
    double S[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double v = shfl8(r.p[q], a);
            if (a == q) {
                const double sd = a == 2 ? 1e-1 : kWPos * h;
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
