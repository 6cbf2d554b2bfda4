// K3, general path: one CTA per box, exact order statistics by a staged radix select over
// the float32 bit patterns (8 bits per pass, shared-memory histograms).  Handles any box
// and any value distribution, so it is the overflow path of the faster kernels (the
// distinct-value table of k3_box_depth.cu, the code histogram of k3_box_depth_raw.cu) and the
// engine of the ablation reducers.  Templated on the depth source (k3_common.cuh).
#pragma once
#include "k3_common.cuh"

namespace stk3 {

constexpr int kHugeThreads = 512, kHugeStage = 45056;      // 176 KB stage, one CTA per SM
constexpr int kMaxCtaWarps = 16;

template <typename Src>
struct BoxCtx {
    Src src;              // depth source of this sequence
    Slice ys, xs;
    unsigned bw, npix;
    const float *staged;  // valid values in shared memory (when n_valid fits the stage)
    int n_valid;
    bool use_stage;
};

// visit every valid depth value of the box: f(value)
template <typename Src, typename F>
__device__ __forceinline__ void for_each_valid_cta(const BoxCtx<Src> &c, F f) {
    if (c.use_stage) {
        for (int i = threadIdx.x; i < c.n_valid; i += blockDim.x) f(c.staged[i]);
    } else {
        for (unsigned i0 = threadIdx.x; i0 < c.npix; i0 += 4 * blockDim.x) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {           // four independent loads in flight
                const unsigned i = i0 + u * blockDim.x;
                v[u] = -1.f;
                if (i < c.npix) {
                    const unsigned ry = i / c.bw, rx = i - ry * c.bw;
                    v[u] = c.src.at(c.ys.lo + (int)ry, c.xs.lo + (int)rx);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (depth_valid(v[u])) f(v[u]);
        }
    }
}

// Radix select of up to two ranks at once over positive-float bit patterns.
// On return value[k] is the element of rank rank[k] in ascending order and
// below[k] the number of elements strictly smaller than it.
template <typename Src>
__device__ void radix_select2_cta(const BoxCtx<Src> &c, int nsel, const int *rank_in, uint32_t *value,
                                  int *below, int *hist /* [2*256] shared */, int *sh /* [8] shared */) {
    uint32_t prefix[2] = {0u, 0u};
    int rank[2] = {rank_in[0], nsel > 1 ? rank_in[1] : 0};
    int less[2] = {0, 0};
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 512; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const uint32_t himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
        const uint32_t p0 = prefix[0], p1 = prefix[1];
        for_each_valid_cta(c, [&](float v) {
            // one shared-memory atomic per distinct bin among the active lanes of a
            // warp: object pixels share one depth and would serialise otherwise
            const uint32_t u = __float_as_uint(v);
            const int lane_ = threadIdx.x & 31;
            const int bin0 = ((u & himask) == p0) ? (int)((u >> shift) & 255u) : -1;
            const unsigned act = __activemask();
            const unsigned peers0 = __match_any_sync(act, bin0);
            if (bin0 >= 0 && lane_ == __ffs(peers0) - 1) atomicAdd(&hist[bin0], __popc(peers0));
            if (nsel > 1) {
                const int bin1 = ((u & himask) == p1) ? (int)((u >> shift) & 255u) : -1;
                const unsigned peers1 = __match_any_sync(act, bin1);
                if (bin1 >= 0 && lane_ == __ffs(peers1) - 1) atomicAdd(&hist[256 + bin1], __popc(peers1));
            }
        });
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0 locates the bin holding each rank: 8 bins per lane + warp scan
            const int lane_ = threadIdx.x;
            for (int k = 0; k < nsel; ++k) {
                int local[8], tot = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) { local[b] = hist[k * 256 + lane_ * 8 + b]; tot += local[b]; }
                int inc = tot;
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane_ >= o) inc += t;
                }
                int acc = inc - tot, found_bin = -1, found_acc = 0;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    if (found_bin < 0 && acc + local[b] > rank[k]) { found_bin = lane_ * 8 + b; found_acc = acc; }
                    acc += local[b];
                }
                const unsigned who = __ballot_sync(0xffffffffu, found_bin >= 0);
                if (who && lane_ == __ffs(who) - 1) { sh[k * 2] = found_bin; sh[k * 2 + 1] = found_acc; }
            }
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

__device__ inline double block_sum(double v, double *red /* [kMaxCtaWarps] shared */) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
}

__device__ inline int block_count(int v, int *red) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
}

// MakeSrc: functor s -> Src of sequence s
template <int kThreadsT, int kStageT, typename MakeSrc>
__global__ void __launch_bounds__(kThreadsT)
k3_box_depth_cta(MakeSrc make_src, int Hp, int Wp, const float *__restrict__ boxes,
                 int cap, float *__restrict__ d_out, float *__restrict__ scale_out,
                 float *__restrict__ inflated, const int32_t *__restrict__ n_work_ptr,
                 const int32_t *__restrict__ worklist) {
  const int n_work = *n_work_ptr;
  for (int item = blockIdx.x; item < n_work; item += gridDim.x) {
    __syncthreads();                                   // shared memory is reused across items
    const int s = worklist[2 * item], bi = worklist[2 * item + 1];
    const int lane = threadIdx.x & 31;
    extern __shared__ __align__(16) float staged[];    // [kStageT]
    __shared__ int hist[512];
    __shared__ int sh[8];
    __shared__ int ired[kMaxCtaWarps];
    __shared__ double dred[kMaxCtaWarps];
    __shared__ int stage_fill;
    __shared__ float corner_sm[16];
    __shared__ int corner_n[4];

    const long long o = (long long)s * cap + bi;
    const float4 box = reinterpret_cast<const float4 *>(boxes)[o];
    const long long b0 = trunc_i64(box.x), b1 = trunc_i64(box.y), b2 = trunc_i64(box.z), b3 = trunc_i64(box.w);
    BoxCtx<decltype(make_src(0))> c;
    c.src = make_src(s);
    if (threadIdx.x < 16) {                 // corner patches, fetched up front (read after the barriers below)
        int cy, cx, cn;
        const bool has = corner_elem(threadIdx.x >> 2, threadIdx.x & 3, b0, b1, b2, b3, Hp, Wp, cy, cx, cn);
        corner_sm[threadIdx.x] = has ? c.src.at(cy, cx) : 0.f;
        if ((threadIdx.x & 3) == 0) corner_n[threadIdx.x >> 2] = cn;
    }
    c.ys = py_slice(b1, b3, Hp);
    c.xs = py_slice(b0, b2, Wp);
    c.bw = (unsigned)max(c.xs.hi - c.xs.lo, 0);
    c.npix = c.bw * (unsigned)max(c.ys.hi - c.ys.lo, 0);
    c.staged = staged;
    c.use_stage = false;
    // (box[2]-box[0]) > 800 in int64; wraps like NumPy for the INT64_MIN case
    const long long wbox = wrap_add(b2, (long long)(0ull - (unsigned long long)b0));
    int n = 0;
    if (!(wbox > 800)) {
        // pass 0: count the valid values and stage the first kStageT of them
        // (warp-aggregated append; the order inside the stage is irrelevant)
        if (threadIdx.x == 0) stage_fill = 0;
        __syncthreads();
        for (unsigned i0 = threadIdx.x; i0 < c.npix; i0 += 4 * blockDim.x) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned i = i0 + u * blockDim.x;
                v[u] = -1.f;
                if (i < c.npix) {
                    const unsigned ry = i / c.bw, rx = i - ry * c.bw;
                    v[u] = c.src.at(c.ys.lo + (int)ry, c.xs.lo + (int)rx);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool ok = depth_valid(v[u]);
                const unsigned act = __activemask();
                const unsigned m = __ballot_sync(act, ok);
                int base = 0;
                if (m) {
                    const int leader = __ffs(m) - 1;
                    if (lane == leader) base = atomicAdd(&stage_fill, __popc(m));
                    base = __shfl_sync(act, base, leader);
                    const int pos = base + __popc(m & ((1u << lane) - 1u));
                    if (ok && pos < kStageT) staged[pos] = v[u];
                }
            }
        }
        __syncthreads();
        n = stage_fill;
        c.use_stage = n <= kStageT;
    }
    c.n_valid = n;

    float d = -1.f, scale = 1.f;
    if (!(n < 1 || wbox > 800)) {
        uint32_t val[2];
        int below[2];
        int rk[2] = {n / 2, 0};
        radix_select2_cta(c, 1, rk, val, below, hist, sh);
        const float d_mid = __uint_as_float(val[0]);
        // corner votes on the unfiltered map (:157-160)
        int k = 0;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) k += corner_mean_of(corner_sm + 4 * cc, corner_n[cc]) > d_mid;
        const Slice seg = window_of(n, k);
        const int m = seg.hi - seg.lo;
        if (m <= 0) {
            d = __int_as_float(0x7fc00000);                    // mean of empty -> NaN
        } else {
            rk[0] = seg.lo; rk[1] = seg.hi - 1;
            radix_select2_cta(c, 2, rk, val, below, hist, sh);
            const uint32_t ua = val[0], ub = val[1];
            double sum;
            if (ua == ub) {
                sum = (double)__uint_as_float(ua) * (double)m;
            } else {
                double part = 0.0;
                int eq_a = 0;
                for_each_valid_cta(c, [&](float v) {
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
        scale = scale_of(d);
    }
    if (threadIdx.x == 0) write_box(o, box, d, scale, d_out, scale_out, inflated);
  }
}

// Workspace layout (int32): [0] number of large boxes, [1] number of overflows (boxes the fast
// large-box kernel hands to the general one), [2] work cursor of the persistent large-box kernel,
// [3] number of small boxes, [4] number of huge boxes (raw path), [5..7] spare, then three lists of
// (sequence, row) pairs with list_cap = S * cap entries each: large boxes (the huge ones of the raw
// path grow from the END of this list: entry k at list_cap - 1 - k), overflows, small boxes.
constexpr int kWsHeader = 8;
__device__ __forceinline__ int32_t *large_list(int32_t *ws) { return ws + kWsHeader; }
__device__ __forceinline__ int32_t *overflow_list(int32_t *ws, int list_cap) { return ws + kWsHeader + 2 * list_cap; }
__device__ __forceinline__ int32_t *small_list(int32_t *ws, int list_cap) { return ws + kWsHeader + 4 * list_cap; }

// One thread per box slot: rows that can reuse the statistic of their source row are finished
// here; the others are sorted into the dense small-box and large-box lists, so that the warp
// kernel and the table kernel get balanced work and can run side by side.
//
// Reuse: the statistic depends on the box only through its truncated corners.  The second call
// of a frame (ocsort_disparity.py:100-101) runs on the deflated track boxes, which are the
// inflated-then-deflated detection boxes: whenever the truncated corners equal those of the
// detection the row came from, the value of the first call is the value of this one.
static __global__ void __launch_bounds__(256)
k3_classify(int S, int Hp, int Wp, const float *__restrict__ boxes, const int32_t *__restrict__ counts, int cap,
            float *__restrict__ d_out, float *__restrict__ scale_out, float *__restrict__ inflated,
            int32_t *__restrict__ worklist, int list_cap, const float *__restrict__ prev_boxes,
            const int32_t *__restrict__ prev_index, const float *__restrict__ prev_d, unsigned small_pix,
            unsigned medium_pix /* 0 = no separate list of huge boxes */) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= (long long)S * cap) return;
    const int s = (int)(o / cap), bi = (int)(o - (long long)s * cap);
    const int cnt = counts[s];
    const float4 box = reinterpret_cast<const float4 *>(boxes)[o];
    const int pi = prev_index ? prev_index[o] : -1;
    if (bi >= cnt) return;
    const long long b0 = trunc_i64(box.x), b1 = trunc_i64(box.y), b2 = trunc_i64(box.z), b3 = trunc_i64(box.w);
    if (pi >= 0 && pi < cap) {
        const float4 pb = reinterpret_cast<const float4 *>(prev_boxes)[(long long)s * cap + pi];
        if (trunc_i64(pb.x) == b0 && trunc_i64(pb.y) == b1 && trunc_i64(pb.z) == b2 && trunc_i64(pb.w) == b3) {
            const float d = prev_d[(long long)s * cap + pi];
            write_box(o, box, d, scale_of(d), d_out, scale_out, inflated);
            return;
        }
    }
    const Slice ys = py_slice(b1, b3, Hp), xs = py_slice(b0, b2, Wp);
    const unsigned npix = (unsigned)max(xs.hi - xs.lo, 0) * (unsigned)max(ys.hi - ys.lo, 0);
    if (medium_pix && npix > medium_pix) {
        const int slot = list_cap - 1 - atomicAdd(&worklist[4], 1);
        large_list(worklist)[2 * slot] = s;
        large_list(worklist)[2 * slot + 1] = bi;
    } else if (npix > small_pix) {
        const int slot = atomicAdd(&worklist[0], 1);
        large_list(worklist)[2 * slot] = s;
        large_list(worklist)[2 * slot + 1] = bi;
    } else {
        const int slot = atomicAdd(&worklist[3], 1);
        small_list(worklist, list_cap)[2 * slot] = s;
        small_list(worklist, list_cap)[2 * slot + 1] = bi;
    }
}

struct MakeMapSrc {
    const float *depth;
    int Hp, Wp;
    __device__ __forceinline__ MapSrc operator()(int s) const {
        MapSrc m; m.dm = depth + (long long)s * Hp * Wp; m.Wp = Wp; return m;
    }
};

struct MakeRawSrc {
    const uint16_t *raw;
    const StFrameInputs *tab;          // a replayed frame reads the frame pointer from its input table
    int H, W;
    float bf;
    __device__ __forceinline__ const uint16_t *base() const { return tab ? tab->raw : raw; }
    __device__ __forceinline__ RawSrc operator()(int s) const {
        RawSrc r; r.raw = base() + (long long)s * H * W; r.H = H; r.W = W; r.bf = bf; return r;
    }
};

}  // namespace stk3
