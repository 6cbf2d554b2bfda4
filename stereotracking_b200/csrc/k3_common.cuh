// Shared pieces of the K3 kernels (per-box depth statistic, ocsort_disparity.py:136-175):
// Python slice / NumPy index semantics of the box, the corner patches, the result
// write-out, and the depth source the pixels are gathered from - the float32 depth
// map K1 wrote, or the raw u16 disparity frame converted on the fly with K1's
// arithmetic (bit-identical: st_code_depth is the expression K1 evaluates).
#pragma once
#include "st_common.cuh"

namespace stk3 {

struct Slice { int lo, hi; };            // hi <= lo means empty

// Python slice semantics of a[start:stop] on an axis of length n (step 1)
__device__ __forceinline__ Slice py_slice(long long start, long long stop, int n) {
    if (start < 0) { start += n; if (start < 0) start = 0; } else if (start > n) start = n;
    if (stop < 0) { stop += n; if (stop < 0) stop = 0; } else if (stop > n) stop = n;
    Slice s; s.lo = (int)start; s.hi = (int)stop;
    return s;
}

// astype(np.int64) of a float32 on x86-64: truncation toward zero, INT64_MIN
// for NaN / out-of-range
__device__ __forceinline__ long long trunc_i64(float v) {
    if (!(fabsf(v) < 9.2233720368547758e18f)) return (long long)0x8000000000000000ull;
    return (long long)v;
}

// int64 arithmetic that wraps like NumPy's (no signed-overflow UB)
__device__ __forceinline__ long long wrap_add(long long a, long long b) {
    return (long long)((unsigned long long)a + (unsigned long long)b);
}

__device__ __forceinline__ bool depth_valid(float d) { return d < 150.f && d > 0.f; }

// The four 2x2 corner patches of a box (:157-160): element e (0..3) of corner c
// (0 tl, 1 tr, 2 bl, 3 br) in row-major order of the (<=2)x(<=2) patch that Python
// slicing yields.  Returns false when the patch has no such element.
__device__ __forceinline__ bool corner_elem(int c, int e, long long b0, long long b1, long long b2, long long b3,
                                            int Hp, int Wp, int &y, int &x, int &cnt) {
    const Slice ys = (c < 2) ? py_slice(b1, wrap_add(b1, 2), Hp) : py_slice(wrap_add(b3, -2), b3, Hp);
    const Slice xs = (c & 1) ? py_slice(wrap_add(b2, -2), b2, Wp) : py_slice(b0, wrap_add(b0, 2), Wp);
    const int ny = max(ys.hi - ys.lo, 0), nx = max(xs.hi - xs.lo, 0);
    cnt = ny * nx;
    if (e >= cnt) return false;
    y = ys.lo + e / nx;
    x = xs.lo + e % nx;
    return true;
}

// np.mean of a corner patch from its prefetched elements v[0..cnt): sequential
// float32 sum in row-major order, one float32 division; empty -> NaN
__device__ __forceinline__ float corner_mean_of(const float *v, int cnt) {
    if (cnt == 0) return __int_as_float(0x7fc00000);
    float acc = v[0];
    for (int e = 1; e < cnt; ++e) acc = __fadd_rn(acc, v[e]);
    return __fdiv_rn(acc, (float)cnt);
}

// scale = max(min(d*d/1, 3.), 1.) with Python min/max NaN behaviour (:170-172)
__device__ __forceinline__ float scale_of(float d) {
    const float dd = __fmul_rn(d, d);
    return (dd != dd) ? dd : fmaxf(fminf(dd, 3.f), 1.f);
}

// scale_bbox (trackers/utils.py:58-73), float32, unfused
__device__ __forceinline__ float4 inflate(float4 box, float scale) {
    const float cx = __fmul_rn(__fadd_rn(box.x, box.z), 0.5f);
    const float cy = __fmul_rn(__fadd_rn(box.y, box.w), 0.5f);
    const float hw = __fmul_rn(__fmul_rn(__fsub_rn(box.z, box.x), scale), 0.5f);
    const float hh = __fmul_rn(__fmul_rn(__fsub_rn(box.w, box.y), scale), 0.5f);
    return make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}

__device__ __forceinline__ void write_box(long long o, float4 box, float d, float scale, float *d_out,
                                          float *scale_out, float *inflated) {
    d_out[o] = d;
    scale_out[o] = scale;
    if (inflated) reinterpret_cast<float4 *>(inflated)[o] = inflate(box, scale);
}

// window of the sorted valid depths (:162-166): [seg.lo, seg.hi) for n values and k corner votes
__device__ __forceinline__ Slice window_of(int n, int k) {
    // w_start = min(1 - k/4, 0.4) * len ; w_end = w_start + 0.6 * len   (float64)
    const double frac = fmin(1.0 - (double)k / 4.0, 0.4);
    const double w_start = __dmul_rn(frac, (double)n);
    const double w_end = __dadd_rn(w_start, __dmul_rn(0.6, (double)n));
    Slice seg = py_slice((long long)w_start, (long long)w_end, n);
    if (seg.hi <= seg.lo) seg = py_slice(0, -1, n);       // d_sorted[:-1]
    return seg;
}

// ---- depth sources ---------------------------------------------------------------------
// The padded depth map [Hp, Wp] of one sequence as K1 wrote it.
struct MapSrc {
    const float *dm;
    int Wp;
    __device__ __forceinline__ float at(int y, int x) const { return dm[(long long)y * Wp + x]; }
};

// The raw u16 frame [H, W] of one sequence seen as that padded map (resize factor 1): pixels
// of the pad and code 65535 are disparity 0 (loading_disparity.py:129-134, transforms_disparity.py
// :177-194), every other code c is depth(c) by K1's expression.
struct RawSrc {
    const uint16_t *raw;
    int H, W;
    float bf;
    __device__ __forceinline__ uint32_t code(int y, int x) const {
        return (y < H && x < W) ? (uint32_t)raw[(long long)y * W + x] : 65535u;
    }
    __device__ __forceinline__ float at(int y, int x) const { return st_code_depth(code(y, x), bf); }
};

}  // namespace stk3
