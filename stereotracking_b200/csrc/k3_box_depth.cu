// K3: per-box depth statistic + depth-driven box scale, one CTA per box.
// Replaces OCSORT_Disparity.extract_depth / bbox_postp_depth
// (mmtrack/models/mot/ocsort_disparity.py:113-130,:136-175) and scale_bbox
// (mmtrack/models/trackers/utils.py:58-73).
//
// The reference sorts the valid depths of the box and averages a 60 % window of
// the sorted list whose position is voted by the four box corners.  Here the
// three order statistics that define the result (median, first and last element
// of the window) are found with an MSB-first radix select over the float bit
// patterns using shared-memory histograms; the window sum is then one masked
// pass (values strictly between the two boundary values, plus the boundary
// values weighted by how many of their duplicates fall inside the window).
#include "st_common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kStage = 4096;             // valid depths staged in shared memory

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

// np.mean of depth[ys, xs] for a (<=2)x(<=2) patch: sequential float32 sum in
// row-major order, then one float32 division; empty -> NaN
__device__ float corner_mean(const float *__restrict__ dm, int Wp, Slice ys, Slice xs) {
    int cnt = 0;
    float acc = 0.f;
    for (int y = ys.lo; y < ys.hi; ++y)
        for (int x = xs.lo; x < xs.hi; ++x) {
            float v = dm[(long long)y * Wp + x];
            acc = cnt ? __fadd_rn(acc, v) : v;
            ++cnt;
        }
    if (cnt == 0) return __int_as_float(0x7fc00000);
    return __fdiv_rn(acc, (float)cnt);
}

struct BoxCtx {
    const float *dm;      // depth map of this sequence
    int Wp;
    Slice ys, xs;
    int bw, bh, npix;
    const float *staged;  // valid values in shared memory (when n_valid <= kStage)
    int n_valid;
    bool use_stage;
};

// visit every valid depth value of the box: f(value)
template <typename F>
__device__ __forceinline__ void for_each_valid(const BoxCtx &c, F f) {
    if (c.use_stage) {
        for (int i = threadIdx.x; i < c.n_valid; i += kThreads) f(c.staged[i]);
    } else {
        for (int i = threadIdx.x; i < c.npix; i += kThreads) {
            int y = c.ys.lo + i / c.bw, x = c.xs.lo + i % c.bw;
            float v = c.dm[(long long)y * c.Wp + x];
            if (depth_valid(v)) f(v);
        }
    }
}

// Radix select of up to two ranks at once over positive-float bit patterns.
// On return value[k] is the element of rank rank[k] in ascending order and
// below[k] the number of elements strictly smaller than it.
__device__ void radix_select2(const BoxCtx &c, int nsel, const int *rank_in, uint32_t *value,
                              int *below, int *hist /* [2*256] shared */, int *sh /* [8] shared */) {
    uint32_t prefix[2] = {0u, 0u};
    int rank[2] = {rank_in[0], nsel > 1 ? rank_in[1] : 0};
    int less[2] = {0, 0};
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 512; i += kThreads) hist[i] = 0;
        __syncthreads();
        const uint32_t himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
        const uint32_t p0 = prefix[0], p1 = prefix[1];
        for_each_valid(c, [&](float v) {
            uint32_t u = __float_as_uint(v);
            if ((u & himask) == p0) atomicAdd(&hist[(u >> shift) & 255u], 1);
            if (nsel > 1 && (u & himask) == p1) atomicAdd(&hist[256 + ((u >> shift) & 255u)], 1);
        });
        __syncthreads();
        if (threadIdx.x < nsel) {
            const int k = threadIdx.x;
            int acc = 0, b = 0;
            for (; b < 256; ++b) {
                int hcnt = hist[k * 256 + b];
                if (acc + hcnt > rank[k]) break;
                acc += hcnt;
            }
            sh[k * 2] = b;
            sh[k * 2 + 1] = acc;
        }
        __syncthreads();
        for (int k = 0; k < nsel; ++k) {
            prefix[k] |= (uint32_t)sh[k * 2] << shift;
            rank[k] -= sh[k * 2 + 1];
            less[k] += sh[k * 2 + 1];
        }
        __syncthreads();
    }
    for (int k = 0; k < nsel; ++k) { value[k] = prefix[k]; below[k] = less[k]; }
}

__device__ double block_sum(double v, double *red /* [kThreads/32] shared */) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    return t;
}

__device__ int block_count(int v, int *red) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    return t;
}

__global__ void __launch_bounds__(kThreads)
k3_box_depth(const float *__restrict__ depth, int Hp, int Wp, const float *__restrict__ boxes,
             const int32_t *__restrict__ counts, int cap, float *__restrict__ d_out,
             float *__restrict__ scale_out, float *__restrict__ inflated) {
    const int s = blockIdx.y, bi = blockIdx.x;
    if (bi >= counts[s]) return;
    __shared__ float staged[kStage];
    __shared__ int hist[512];
    __shared__ int sh[8];
    __shared__ int ired[kThreads / 32];
    __shared__ double dred[kThreads / 32];
    __shared__ int stage_fill;

    const long long o = (long long)s * cap + bi;
    const float4 box = reinterpret_cast<const float4 *>(boxes)[o];
    const long long b0 = trunc_i64(box.x), b1 = trunc_i64(box.y), b2 = trunc_i64(box.z), b3 = trunc_i64(box.w);
    BoxCtx c;
    c.dm = depth + (long long)s * Hp * Wp;
    c.Wp = Wp;
    c.ys = py_slice(b1, b3, Hp);
    c.xs = py_slice(b0, b2, Wp);
    c.bw = max(c.xs.hi - c.xs.lo, 0);
    c.bh = max(c.ys.hi - c.ys.lo, 0);
    c.npix = c.bw * c.bh;
    c.staged = staged;
    c.use_stage = false;
    // pass 0: count valid values (and stage them when they fit)
    int local = 0;
    for (int i = threadIdx.x; i < c.npix; i += kThreads) {
        int y = c.ys.lo + i / c.bw, x = c.xs.lo + i % c.bw;
        local += depth_valid(c.dm[(long long)y * Wp + x]);
    }
    const int n = block_count(local, ired);
    c.n_valid = n;
    // (box[2]-box[0]) > 800 in int64; wraps like NumPy for the INT64_MIN case
    const long long wbox = wrap_add(b2, (long long)(0ull - (unsigned long long)b0));

    float d = -1.f, scale = 1.f;
    if (!(n < 1 || wbox > 800)) {
        if (n <= kStage) {
            if (threadIdx.x == 0) stage_fill = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < c.npix; i += kThreads) {
                int y = c.ys.lo + i / c.bw, x = c.xs.lo + i % c.bw;
                float v = c.dm[(long long)y * Wp + x];
                if (depth_valid(v)) staged[atomicAdd(&stage_fill, 1)] = v;
            }
            __syncthreads();
            c.use_stage = true;
        }
        uint32_t val[2];
        int below[2];
        int rk[2] = {n / 2, 0};
        radix_select2(c, 1, rk, val, below, hist, sh);
        const float d_mid = __uint_as_float(val[0]);
        // corner votes on the unfiltered map (:157-160)
        int k = 0;
        {
            Slice ytop = py_slice(b1, wrap_add(b1, 2), Hp), ybot = py_slice(wrap_add(b3, -2), b3, Hp);
            Slice xl = py_slice(b0, wrap_add(b0, 2), Wp), xr = py_slice(wrap_add(b2, -2), b2, Wp);
            k += corner_mean(c.dm, Wp, ytop, xl) > d_mid;
            k += corner_mean(c.dm, Wp, ytop, xr) > d_mid;
            k += corner_mean(c.dm, Wp, ybot, xl) > d_mid;
            k += corner_mean(c.dm, Wp, ybot, xr) > d_mid;
        }
        // w_start = min(1 - k/4, 0.4) * len ; w_end = w_start + 0.6 * len   (float64)
        const double frac = fmin(1.0 - (double)k / 4.0, 0.4);
        const double w_start = __dmul_rn(frac, (double)n);
        const double w_end = __dadd_rn(w_start, __dmul_rn(0.6, (double)n));
        Slice seg = py_slice((long long)w_start, (long long)w_end, n);
        if (seg.hi <= seg.lo) seg = py_slice(0, -1, n);       // d_sorted[:-1]
        const int m = seg.hi - seg.lo;
        if (m <= 0) {
            d = __int_as_float(0x7fc00000);                    // mean of empty -> NaN
        } else {
            rk[0] = seg.lo; rk[1] = seg.hi - 1;
            radix_select2(c, 2, rk, val, below, hist, sh);
            const uint32_t ua = val[0], ub = val[1];
            double sum;
            if (ua == ub) {
                sum = (double)__uint_as_float(ua) * (double)m;
            } else {
                double part = 0.0;
                int eq_a = 0;
                for_each_valid(c, [&](float v) {
                    uint32_t u = __float_as_uint(v);
                    if (u > ua && u < ub) part += (double)v;
                    eq_a += (u == ua);
                });
                sum = block_sum(part, dred);
                const int n_eq_a = block_count(eq_a, ired);
                const int in_a = below[0] + n_eq_a - seg.lo;   // duplicates of s[lo] inside
                const int in_b = seg.hi - below[1];            // duplicates of s[hi-1] inside
                sum += (double)__uint_as_float(ua) * (double)in_a + (double)__uint_as_float(ub) * (double)in_b;
            }
            d = (float)(sum / (double)m);
        }
        // scale = max(min(d*d/1, 3.), 1.) with Python min/max NaN behaviour
        const float dd = __fmul_rn(d, d);
        scale = (dd != dd) ? dd : fmaxf(fminf(dd, 3.f), 1.f);
    }
    if (threadIdx.x == 0) {
        d_out[o] = d;
        scale_out[o] = scale;
        if (inflated) {
            // scale_bbox (trackers/utils.py:58-73), float32, unfused
            const float cx = __fmul_rn(__fadd_rn(box.x, box.z), 0.5f);
            const float cy = __fmul_rn(__fadd_rn(box.y, box.w), 0.5f);
            const float hw = __fmul_rn(__fmul_rn(__fsub_rn(box.z, box.x), scale), 0.5f);
            const float hh = __fmul_rn(__fmul_rn(__fsub_rn(box.w, box.y), scale), 0.5f);
            reinterpret_cast<float4 *>(inflated)[o] =
                make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
        }
    }
}

__global__ void k3_scale_boxes(const float *__restrict__ boxes, const float *__restrict__ scales,
                               const int32_t *__restrict__ counts, int cap, int invert,
                               float *__restrict__ out) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[s]) return;
    const long long o = (long long)s * cap + i;
    const float4 b = reinterpret_cast<const float4 *>(boxes)[o];
    float sc = scales[o];
    if (invert) sc = __frcp_rn(sc);           // torch `1 / scales`
    const float cx = __fmul_rn(__fadd_rn(b.x, b.z), 0.5f);
    const float cy = __fmul_rn(__fadd_rn(b.y, b.w), 0.5f);
    const float hw = __fmul_rn(__fmul_rn(__fsub_rn(b.z, b.x), sc), 0.5f);
    const float hh = __fmul_rn(__fmul_rn(__fsub_rn(b.w, b.y), sc), 0.5f);
    reinterpret_cast<float4 *>(out)[o] =
        make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
}

}  // namespace

extern "C" int st_box_depth(const float *depth, int S, int Hp, int Wp, const float *boxes,
                            const int32_t *counts, int cap, float *d, float *scale,
                            float *inflated, void *stream) {
    if (!depth || !boxes || !counts || !d || !scale || S < 0 || cap < 1 || Hp < 1 || Wp < 1)
        return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
    k3_box_depth<<<dim3(cap, S), kThreads, 0, (cudaStream_t)stream>>>(depth, Hp, Wp, boxes, counts, cap, d,
                                                                      scale, inflated);
    return st_last_launch();
}

extern "C" int st_scale_boxes(const float *boxes, const float *scales, const int32_t *counts, int S,
                              int cap, int invert, float *out, void *stream) {
    if (!boxes || !scales || !counts || !out || S < 0 || cap < 1) return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
    k3_scale_boxes<<<dim3((cap + 127) / 128, S), 128, 0, (cudaStream_t)stream>>>(boxes, scales, counts, cap,
                                                                                invert, out);
    return st_last_launch();
}
