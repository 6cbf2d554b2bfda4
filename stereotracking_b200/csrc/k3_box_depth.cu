// K3: per-box depth statistic + depth-driven box scale.
// Replaces OCSORT_Disparity.extract_depth / bbox_postp_depth
// (mmtrack/models/mot/ocsort_disparity.py:113-130,:136-175) and scale_bbox
// (mmtrack/models/trackers/utils.py:58-73).
//
// The reference sorts the valid depths of the box and averages a 60 % window of
// the sorted list whose position is voted by the four box corners.  Kernels:
//   k3_classify         one thread per box slot: rows that may reuse an earlier statistic
//                       are finished, the rest go to a dense small-box or large-box list
//   k3_box_depth        one WARP per box of <= 512 pixels (airborne objects are small): valid
//                       depths staged in shared memory; the three order statistics that define
//                       the result (median, first and last element of the window) by an
//                       MSB-first bit-wise select (one pass + one REDUX per bit); the window sum
//                       as one masked pass with the boundary duplicates weighted exactly
//   k3_box_depth_table  one CTA per larger box, on an auxiliary stream beside the warp kernel:
//                       shared-memory hash table of the DISTINCT depth values with their
//                       multiplicities, ranked; order statistics and window sum from the
//                       sorted multiset
//   k3_box_depth_cta    overflow of the table (> 3072 distinct values): staged radix select
//                       with shared-memory histograms, 8 bits per pass
#include <cstdlib>
#include <map>
#include <mutex>
#include "k3_select.cuh"

namespace {

constexpr int kWarps = 4;                // boxes per CTA
constexpr int kStage = 512;              // valid depths staged per warp
constexpr int kSmallPix = 512;           // boxes up to this many pixels: one warp; larger: one CTA
// (2048 before: the warp kernel's longest box - ~27 bit passes over 2048 staged values - set its
// duration at ~70 us; 512-2048 pixel boxes cost the table kernel's 512 threads a few microseconds)

using namespace stk3;

struct WarpCtx {
    const float *dm;      // depth map of this sequence
    int Wp;
    Slice ys, xs;
    unsigned bw, npix;
    const float *staged;  // valid values in shared memory (when n_valid <= kStage)
    int n_valid;
    bool use_stage;
};

// visit every valid depth value of the box with the 32 lanes of the warp
template <typename F>
__device__ __forceinline__ void for_each_valid(const WarpCtx &c, int lane, F f) {
    if (c.use_stage) {
        for (int i = lane; i < c.n_valid; i += 32) f(c.staged[i]);
    } else {
        for (unsigned i = lane; i < c.npix; i += 32) {
            const unsigned ry = i / c.bw, rx = i - ry * c.bw;
            const float v = c.dm[(long long)(c.ys.lo + ry) * c.Wp + c.xs.lo + rx];
            if (depth_valid(v)) f(v);
        }
    }
}

// MSB-first binary search on the bit patterns of the staged values (one warp,
// no histogram, no atomics): per bit one pass over shared memory and one
// redux.  Same contract as radix_select2.
__device__ void bit_select2(const float *staged, int n, int lane, int nsel, const int *rank_in,
                            uint32_t *value, int *below) {
    // bits above the highest differing bit are common to every value
    uint32_t vor = 0u, vand = 0xffffffffu;
    for (int i = lane; i < n; i += 32) { const uint32_t u = __float_as_uint(staged[i]); vor |= u; vand &= u; }
    vor = __reduce_or_sync(0xffffffffu, vor);
    vand = __reduce_and_sync(0xffffffffu, vand);
    const uint32_t diff = vor ^ vand;
    const int top = diff ? 31 - __clz(diff) : -1;          // highest bit that varies
    const uint32_t common = top >= 31 ? 0u : (top < 0 ? vand : (vand & ~((2u << top) - 1u)));
    uint32_t prefix[2] = {common, common};
    int rank[2] = {rank_in[0], nsel > 1 ? rank_in[1] : 0};
    int less[2] = {0, 0};
    for (int bit = top; bit >= 0; --bit) {
        const uint32_t hi = ~((2u << bit) - 1u);            // bits above `bit`
        int c0 = 0, c1 = 0;
        for (int i = lane; i < n; i += 32) {
            const uint32_t u = __float_as_uint(staged[i]);
            const bool zero = !((u >> bit) & 1u);
            c0 += zero && ((u & hi) == prefix[0]);
            c1 += zero && ((u & hi) == prefix[1]);
        }
        c0 = __reduce_add_sync(0xffffffffu, c0);
        if (rank[0] >= c0) { rank[0] -= c0; less[0] += c0; prefix[0] |= 1u << bit; }
        if (nsel > 1) {
            c1 = __reduce_add_sync(0xffffffffu, c1);
            if (rank[1] >= c1) { rank[1] -= c1; less[1] += c1; prefix[1] |= 1u << bit; }
        }
    }
    for (int k = 0; k < nsel; ++k) { value[k] = prefix[k]; below[k] = less[k]; }
}

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one small box by one warp; `staged` is the warp's shared-memory stage
__device__ void k3_box_one(const float *__restrict__ depth, int Hp, int Wp, const float *__restrict__ boxes,
                           const int s, const int bi, int cap, float *__restrict__ d_out,
                           float *__restrict__ scale_out, float *__restrict__ inflated, float *staged,
                           const int lane) {
    const long long o = (long long)s * cap + bi;
    const float4 box = reinterpret_cast<const float4 *>(boxes)[o];
    const long long b0 = trunc_i64(box.x), b1 = trunc_i64(box.y), b2 = trunc_i64(box.z), b3 = trunc_i64(box.w);
    WarpCtx c;
    c.dm = depth + (long long)s * Hp * Wp;
    c.Wp = Wp;
    c.ys = py_slice(b1, b3, Hp);
    c.xs = py_slice(b0, b2, Wp);
    c.bw = (unsigned)max(c.xs.hi - c.xs.lo, 0);
    c.npix = c.bw * (unsigned)max(c.ys.hi - c.ys.lo, 0);
    c.staged = staged;
    c.use_stage = true;
    // corner patches: lane 4c+e fetches element e of corner c now, so that the 16 loads
    // are in flight together and hidden behind the staging pass (they were a chain of 16
    // dependent loads after the median before)
    float corner_val = 0.f;
    int corner_cnt = 0;
    if (lane < 16) {
        int cy, cx;
        if (corner_elem(lane >> 2, lane & 3, b0, b1, b2, b3, Hp, Wp, cy, cx, corner_cnt))
            corner_val = c.dm[(long long)cy * Wp + cx];
    }
    // pass 0: count the valid values, staging them (npix <= kSmallPix = kStage always fits)
    int n = 0;
    for (unsigned i0 = 0; i0 < c.npix; i0 += 128) {
        float v[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {              // four independent loads in flight
            const unsigned i = i0 + u * 32 + lane;
            v[u] = 0.f;
            ok[u] = false;
            if (i < c.npix) {
                const unsigned ry = i / c.bw, rx = i - ry * c.bw;
                v[u] = c.dm[(long long)(c.ys.lo + ry) * Wp + c.xs.lo + rx];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            ok[u] = (i0 + u * 32 + lane < c.npix) && depth_valid(v[u]);
            const unsigned m = __ballot_sync(0xffffffffu, ok[u]);
            const int pos = n + __popc(m & ((1u << lane) - 1u));
            if (ok[u] && pos < kStage) staged[pos] = v[u];
            n += __popc(m);
        }
    }
    __syncwarp();
    c.n_valid = n;
    if (n > kStage) c.use_stage = false;
    // (box[2]-box[0]) > 800 in int64; wraps like NumPy for the INT64_MIN case
    const long long wbox = wrap_add(b2, (long long)(0ull - (unsigned long long)b0));

    float d = -1.f, scale = 1.f;
    if (!(n < 1 || wbox > 800)) {
        uint32_t val[2];
        int below[2];
        int rk[2] = {n / 2, 0};
        bit_select2(staged, n, lane, 1, rk, val, below);
        const float d_mid = __uint_as_float(val[0]);
        // corner votes on the unfiltered map (:157-160): lane 4c gathers the elements of corner c
        int k = 0;
        {
            float cv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) cv[e] = __shfl_sync(0xffffffffu, corner_val, (lane & 12) + e);
            const bool vote = lane < 16 && (lane & 3) == 0 && corner_mean_of(cv, corner_cnt) > d_mid;
            k = __popc(__ballot_sync(0xffffffffu, vote));
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
            bit_select2(staged, n, lane, 2, rk, val, below);
            const uint32_t ua = val[0], ub = val[1];
            double sum;
            if (ua == ub) {
                sum = (double)__uint_as_float(ua) * (double)m;
            } else {
                double part = 0.0;
                int eq_a = 0;
                for_each_valid(c, lane, [&](float v) {
                    const uint32_t u = __float_as_uint(v);
                    if (u > ua && u < ub) part += (double)v;
                    eq_a += (u == ua);
                });
                sum = warp_sum(part);
                const int n_eq_a = __reduce_add_sync(0xffffffffu, eq_a);
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
    if (lane == 0) {
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

// Small boxes: a resident wave of 4-warp CTAs, warps stride over the dense small-box list.
__global__ void __launch_bounds__(kWarps * 32)
k3_box_depth(const float *__restrict__ depth, int Hp, int Wp, const float *__restrict__ boxes, int cap,
             float *__restrict__ d_out, float *__restrict__ scale_out, float *__restrict__ inflated,
             const int32_t *__restrict__ worklist, int list_cap) {
    __shared__ float staged_all[kWarps][kStage];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = worklist[3];
    const int32_t *list = worklist + 8 + 4 * list_cap;
    for (int idx = blockIdx.x * kWarps + warp; idx < n; idx += gridDim.x * kWarps) {
        k3_box_one(depth, Hp, Wp, boxes, list[2 * idx], list[2 * idx + 1], cap, d_out, scale_out, inflated,
                   staged_all[warp], lane);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Large boxes, main path: one pass over the pixels builds a shared-memory table of
// the DISTINCT valid depth values with their multiplicities (open-addressing hash on
// the float bit pattern, warp-aggregated inserts).  A depth map that comes from a
// u16 disparity has few distinct values per box (one per object plus the
// background ramp), so the table is tiny compared with the pixel count; the
// order statistics and the window sum are then computed on the sorted table with
// multiplicities - exactly the sorted-list semantics of the reference.  Boxes
// with more distinct values than the table holds go to the overflow list and are
// handled by the staged radix-select kernel below.
constexpr int kTableSlots = 4096;        // hash slots (power of two)
constexpr int kMaxDistinct = 3072;       // distinct values accepted before overflow
constexpr int kTableThreads = 512;
constexpr int kTablePerThread = 8;      // pixel loads in flight per thread (the pass is latency-bound)
constexpr size_t kTableSmem = sizeof(uint32_t) * kTableSlots * 2 + sizeof(uint32_t) * kMaxDistinct * 5;

__global__ void __launch_bounds__(kTableThreads)
k3_box_depth_table(const float *__restrict__ depth, int Hp, int Wp, const float *__restrict__ boxes,
                   int cap, float *__restrict__ d_out, float *__restrict__ scale_out,
                   float *__restrict__ inflated, int32_t *__restrict__ worklist, int list_cap) {
    extern __shared__ __align__(16) uint32_t tsm[];
    uint32_t *keys = tsm;                               // [kTableSlots] 0 = empty
    int *cnts = reinterpret_cast<int *>(tsm + kTableSlots);
    uint32_t *ukey = tsm + 2 * kTableSlots;             // [kMaxDistinct] unordered distinct values
    int *ucnt = reinterpret_cast<int *>(ukey + kMaxDistinct);
    float *sval = reinterpret_cast<float *>(ucnt + kMaxDistinct);    // sorted by value
    int *scnt = reinterpret_cast<int *>(sval + kMaxDistinct);
    int *sbelow = scnt + kMaxDistinct;                  // number of elements strictly smaller
    __shared__ int n_distinct, n_dense, overflow, n_valid_sh, sel[2];
    __shared__ double dred[kMaxCtaWarps];
    __shared__ float corner_sm[16];
    __shared__ int corner_n[4];
    const int lane = threadIdx.x & 31;
    const int n_work = worklist[0];
    __shared__ int next_item;
    // items are fetched from a shared cursor: boxes range from 513 to 65 000+ pixels, and with a
    // static stride one CTA could draw several of the largest while others idled
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) next_item = atomicAdd(&worklist[2], 1);
        __syncthreads();
        const int item = next_item;
        if (item >= n_work) break;
        const int s = worklist[8 + 2 * item], bi = worklist[9 + 2 * item];
        const long long o = (long long)s * cap + bi;
        const float4 box = reinterpret_cast<const float4 *>(boxes)[o];
        const long long b0 = trunc_i64(box.x), b1 = trunc_i64(box.y), b2 = trunc_i64(box.z), b3 = trunc_i64(box.w);
        const float *dm = depth + (long long)s * Hp * Wp;
        const Slice ys = py_slice(b1, b3, Hp), xs = py_slice(b0, b2, Wp);
        const unsigned bw = (unsigned)max(xs.hi - xs.lo, 0);
        const unsigned npix = bw * (unsigned)max(ys.hi - ys.lo, 0);
        const long long wbox = wrap_add(b2, (long long)(0ull - (unsigned long long)b0));
        float d = -1.f, scale = 1.f;
        bool done = wbox > 800;                          // reference: d = -1, scale = 1 (:150)
        if (!done) {
            if (threadIdx.x < 16) {         // corner patches, fetched up front (read after the barriers below)
                int cy, cx, cn;
                const bool has = corner_elem(threadIdx.x >> 2, threadIdx.x & 3, b0, b1, b2, b3, Hp, Wp, cy, cx, cn);
                corner_sm[threadIdx.x] = has ? dm[(long long)cy * Wp + cx] : 0.f;
                if ((threadIdx.x & 3) == 0) corner_n[threadIdx.x >> 2] = cn;
            }
            for (int i = threadIdx.x; i < kTableSlots; i += kTableThreads) { keys[i] = 0u; cnts[i] = 0; }
            if (threadIdx.x == 0) { n_distinct = 0; n_dense = 0; overflow = 0; n_valid_sh = 0; }
            __syncthreads();
            for (unsigned i0 = threadIdx.x; i0 < npix; i0 += kTablePerThread * kTableThreads) {
                float v[kTablePerThread];
#pragma unroll
                for (int u = 0; u < kTablePerThread; ++u) {   // independent loads in flight
                    const unsigned i = i0 + u * kTableThreads;
                    v[u] = -1.f;
                    if (i < npix) {
                        const unsigned ry = i / bw, rx = i - ry * bw;
                        v[u] = dm[(long long)(ys.lo + ry) * Wp + xs.lo + rx];
                    }
                }
#pragma unroll
                for (int u = 0; u < kTablePerThread; ++u) {
                    const bool ok = depth_valid(v[u]);
                    const uint32_t key = ok ? __float_as_uint(v[u]) : 0u;     // valid depths are > 0: bits != 0
                    const unsigned act = __activemask();
                    const unsigned peers = __match_any_sync(act, key);
                    if (ok && lane == __ffs(peers) - 1) {
                        const int c = __popc(peers);
                        uint32_t h = (key * 2654435761u) >> 20;               // top 12 bits -> 4096 slots
                        for (int probe = 0; probe < kTableSlots; ++probe) {
                            const uint32_t prev = atomicCAS(&keys[h], 0u, key);
                            if (prev == 0u || prev == key) {
                                atomicAdd(&cnts[h], c);
                                if (prev == 0u && atomicAdd(&n_distinct, 1) >= kMaxDistinct) overflow = 1;
                                break;
                            }
                            h = (h + 1) & (kTableSlots - 1);
                        }
                    }
                }
                if (overflow) break;                     // benign race: every thread leaves soon after
            }
            __syncthreads();
            if (overflow) {
                if (threadIdx.x == 0) {
                    const int slot = atomicAdd(&worklist[1], 1);
                    worklist[8 + 2 * list_cap + 2 * slot] = s;
                    worklist[9 + 2 * list_cap + 2 * slot] = bi;
                }
                continue;                                // block-uniform
            }
            // dense (unordered) list of the distinct values
            for (int i = threadIdx.x; i < kTableSlots; i += kTableThreads)
                if (keys[i] != 0u) {
                    const int k = atomicAdd(&n_dense, 1);
                    ukey[k] = keys[i];
                    ucnt[k] = cnts[i];
                    atomicAdd(&n_valid_sh, cnts[i]);
                }
            __syncthreads();
            const int T = n_dense, n = n_valid_sh;
            if (n >= 1) {
                // sort by value: position = number of smaller distinct values (keys are unique)
                for (int e = threadIdx.x; e < T; e += kTableThreads) {
                    const uint32_t ke = ukey[e];
                    int pos = 0, below = 0;
                    for (int f = 0; f < T; ++f) {
                        const bool sm = ukey[f] < ke;
                        pos += sm;
                        below += sm ? ucnt[f] : 0;
                    }
                    sval[pos] = __uint_as_float(ke);
                    scnt[pos] = ucnt[e];
                    sbelow[pos] = below;
                }
                __syncthreads();
                auto locate = [&](int rank, int which) {   // sorted[rank] -> sel[which] (entry index)
                    for (int e = threadIdx.x; e < T; e += kTableThreads)
                        if (sbelow[e] <= rank && rank < sbelow[e] + scnt[e]) sel[which] = e;
                };
                locate(n / 2, 0);
                __syncthreads();
                const float d_mid = sval[sel[0]];
                int k = 0;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) k += corner_mean_of(corner_sm + 4 * cc, corner_n[cc]) > d_mid;
                const double frac = fmin(1.0 - (double)k / 4.0, 0.4);
                const double w_start = __dmul_rn(frac, (double)n);
                const double w_end = __dadd_rn(w_start, __dmul_rn(0.6, (double)n));
                Slice seg = py_slice((long long)w_start, (long long)w_end, n);
                if (seg.hi <= seg.lo) seg = py_slice(0, -1, n);       // d_sorted[:-1]
                const int m = seg.hi - seg.lo;
                if (m <= 0) {
                    d = __int_as_float(0x7fc00000);
                } else {
                    // window sum over the sorted multiset, fixed summation order
                    double part = 0.0;
                    for (int e = threadIdx.x; e < T; e += kTableThreads) {
                        const int lo = max(sbelow[e], seg.lo), hi = min(sbelow[e] + scnt[e], seg.hi);
                        if (hi > lo) part += (double)sval[e] * (double)(hi - lo);
                    }
                    const double sum = block_sum(part, dred);
                    d = (float)(sum / (double)m);
                }
                const float dd = __fmul_rn(d, d);
                scale = (dd != dd) ? dd : fmaxf(fminf(dd, 3.f), 1.f);
            }
        }
        if (threadIdx.x == 0) {
            d_out[o] = d;
            scale_out[o] = scale;
            if (inflated) {
                const float cx = __fmul_rn(__fadd_rn(box.x, box.z), 0.5f);
                const float cy = __fmul_rn(__fadd_rn(box.y, box.w), 0.5f);
                const float hw = __fmul_rn(__fmul_rn(__fsub_rn(box.z, box.x), scale), 0.5f);
                const float hh = __fmul_rn(__fmul_rn(__fsub_rn(box.w, box.y), scale), 0.5f);
                reinterpret_cast<float4 *>(inflated)[o] =
                    make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// Ablation reducers (mmtrack/models/mot/depth_extraction_comparison.py:5-150): the same
// patch, validity filter and -1 rule as extract_depth, another statistic of the valid
// depths, and scale = clip(d*d/400, 1, 3).  One CTA per box, values read in place
// (these variants are for comparison runs, not for the frame pipeline):
//   1 truncated mean  mean(sorted[int(0.1 n) : int(0.9 n)]), sorted[:-1] when empty (:29-36)
//   2 mean            mean of the valid depths (:71)
//   3 median          np.median: sorted[n/2], or the mean of the two middle elements (:106)
//   4 centre          the UNFILTERED map at ((b1+b3)//2, (b0+b2)//2) with NumPy indexing:
//                     negative indices wrap; out of range raises IndexError in the
//                     reference (:141) -> d = NaN, scale = 1 and oob[s] is incremented
constexpr int kAblationThreads = 256;

__global__ void __launch_bounds__(kAblationThreads)
k3_box_depth_ablation(const float *__restrict__ depth, int Hp, int Wp, const float *__restrict__ boxes,
                      const int32_t *__restrict__ counts, int cap, int mode, float *__restrict__ d_out,
                      float *__restrict__ scale_out, float *__restrict__ inflated, int32_t *__restrict__ oob) {
    __shared__ int hist[512];
    __shared__ int sh[8];
    __shared__ int ired[kMaxCtaWarps];
    __shared__ double dred[kMaxCtaWarps];
    const int s = blockIdx.y;
    const int n_box = min(counts[s], cap);
    for (int bi = blockIdx.x; bi < n_box; bi += gridDim.x) {
        __syncthreads();                               // shared memory is reused across boxes
        const long long o = (long long)s * cap + bi;
        const float4 box = reinterpret_cast<const float4 *>(boxes)[o];
        const long long b0 = trunc_i64(box.x), b1 = trunc_i64(box.y), b2 = trunc_i64(box.z), b3 = trunc_i64(box.w);
        BoxCtx<MapSrc> c;
        c.src.dm = depth + (long long)s * Hp * Wp;
        c.src.Wp = Wp;
        c.ys = py_slice(b1, b3, Hp);
        c.xs = py_slice(b0, b2, Wp);
        c.bw = (unsigned)max(c.xs.hi - c.xs.lo, 0);
        c.npix = c.bw * (unsigned)max(c.ys.hi - c.ys.lo, 0);
        c.staged = nullptr;
        c.use_stage = false;
        c.n_valid = 0;
        const long long wbox = wrap_add(b2, (long long)(0ull - (unsigned long long)b0));
        int n = 0;
        if (!(wbox > 800)) {
            int mine = 0;
            for_each_valid_cta(c, [&](float) { ++mine; });
            n = block_count(mine, ired);
        }
        c.n_valid = n;
        float d = -1.f, scale = 1.f;
        bool outside = false;
        if (!(n < 1 || wbox > 800)) {
            if (mode == 4) {
                // (box[0] + box[2]) // 2 in int64: wrapping add, floor division
                long long cx = wrap_add(b0, b2) >> 1, cy = wrap_add(b1, b3) >> 1;
                if (cx < 0) cx += Wp;
                if (cy < 0) cy += Hp;
                outside = cx < 0 || cx >= Wp || cy < 0 || cy >= Hp;
                d = outside ? __int_as_float(0x7fc00000) : c.src.at((int)cy, (int)cx);
            } else {
                Slice seg;
                if (mode == 1) {
                    seg = py_slice((long long)__dmul_rn(0.1, (double)n), (long long)__dmul_rn(1.0 - 0.1, (double)n), n);
                    if (seg.hi <= seg.lo) seg = py_slice(0, -1, n);   // d_sorted[:-1]
                } else if (mode == 2) {
                    seg.lo = 0; seg.hi = n;
                } else {
                    seg.lo = (n - 1) / 2; seg.hi = n / 2 + 1;
                }
                const int m = seg.hi - seg.lo;
                if (m <= 0) {
                    d = __int_as_float(0x7fc00000);                   // mean of empty -> NaN
                } else {
                    uint32_t val[2];
                    int below[2];
                    int rk[2] = {seg.lo, seg.hi - 1};
                    radix_select2_cta(c, 2, rk, val, below, hist, sh);
                    const uint32_t ua = val[0], ub = val[1];
                    double sum;
                    if (ua == ub) {
                        sum = (double)__uint_as_float(ua) * (double)m;
                    } else {
                        double part = 0.0;
                        int eq_a = 0;
                        for_each_valid_cta(c, [&](float v) {
                            const uint32_t u = __float_as_uint(v);
                            if (u > ua && u < ub) part += (double)v;
                            eq_a += (u == ua);
                        });
                        sum = block_sum(part, dred);
                        const int n_eq_a = block_count(eq_a, ired);
                        const int in_a = below[0] + n_eq_a - seg.lo;  // duplicates of s[lo] inside
                        const int in_b = seg.hi - below[1];           // duplicates of s[hi-1] inside
                        sum += (double)__uint_as_float(ua) * (double)in_a + (double)__uint_as_float(ub) * (double)in_b;
                    }
                    d = (float)(sum / (double)m);
                }
            }
            if (!outside) {
                // scale = max(min(d*d/400, 3.), 1.) with Python min/max NaN behaviour
                const float dd = __fdiv_rn(__fmul_rn(d, d), 400.f);
                scale = (dd != dd) ? dd : fmaxf(fminf(dd, 3.f), 1.f);
            }
        }
        if (threadIdx.x == 0) {
            d_out[o] = d;
            scale_out[o] = scale;
            if (outside && oob) atomicAdd(&oob[s], 1);
            if (inflated) {
                const float cx = __fmul_rn(__fadd_rn(box.x, box.z), 0.5f);
                const float cy = __fmul_rn(__fadd_rn(box.y, box.w), 0.5f);
                const float hw = __fmul_rn(__fmul_rn(__fsub_rn(box.z, box.x), scale), 0.5f);
                const float hh = __fmul_rn(__fmul_rn(__fsub_rn(box.w, box.y), scale), 0.5f);
                reinterpret_cast<float4 *>(inflated)[o] =
                    make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
            }
        }
    }
}

__global__ void k3_scale_boxes(const float *__restrict__ boxes, const float *__restrict__ scales,
                               const int32_t *__restrict__ counts, int cap, int invert,
                               float *__restrict__ out) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= counts[s] || i >= cap) return;
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

struct K3Dev { int sms, warp_ctas, table_per_sm; };
StPerDevice<K3Dev> g_k3dev;

}  // namespace

extern "C" int st_box_depth(const float *depth, int S, int Hp, int Wp, const float *boxes,
                            const int32_t *counts, int cap, float *d, float *scale,
                            float *inflated, int32_t *workspace, void *stream) {
    return st_box_depth_reuse(depth, S, Hp, Wp, boxes, counts, cap, d, scale, inflated, workspace, nullptr,
                              nullptr, nullptr, stream);
}

extern "C" int st_box_depth_reuse(const float *depth, int S, int Hp, int Wp, const float *boxes,
                                  const int32_t *counts, int cap, float *d, float *scale, float *inflated,
                                  int32_t *workspace, const float *prev_boxes, const int32_t *prev_index,
                                  const float *prev_d, void *stream) {
    if (!depth || !boxes || !counts || !d || !scale || !workspace || S < 0 || cap < 1 || Hp < 1 || Wp < 1)
        return ST_ERR_INVALID_ARG;
    if (prev_index && (!prev_boxes || !prev_d)) return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ST_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(int32_t) * 8, st));
    const int list_cap = S * cap;
    // once per device: shared-memory opt-ins, resident grid of the warp kernel
    const K3Dev *dv = g_k3dev.get([](K3Dev &v) {
        v.sms = st_ctx_sms();
        if (cudaFuncSetAttribute(k3_box_depth_cta<kHugeThreads, kHugeStage, MakeMapSrc>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(float) * kHugeStage)) != cudaSuccess) return false;
        if (cudaFuncSetAttribute(k3_box_depth_table, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)kTableSmem) != cudaSuccess) return false;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3_box_depth, kWarps * 32, 0) != cudaSuccess || per_sm < 1)
            per_sm = 8;
        v.warp_ctas = v.sms * per_sm;
        const char *e = getenv("ST_K3_TABLE_CTAS_PER_SM");
        v.table_per_sm = e ? atoi(e) : 2;
        return true;
    });
    if (!dv) return ST_ERR_CUDA;
    // the large boxes (table kernel, then its overflow list) run on the auxiliary stream paired
    // with the caller's, the small boxes on the caller's stream; join before returning
    cudaStream_t aux;
    cudaEvent_t ev_fork, ev_join;
    if (st_ctx_aux(st, &aux, &ev_fork, &ev_join) != ST_OK) return ST_ERR_CUDA;
    const long long slots = (long long)S * cap;
    k3_classify<<<(int)((slots + 255) / 256), 256, 0, st>>>(S, Hp, Wp, boxes, counts, cap, d, scale, inflated, workspace,
                                                           list_cap, prev_boxes, prev_index, prev_d, (unsigned)kSmallPix, 0u);
    ST_CUDA_OK(cudaEventRecord(ev_fork, st));
    ST_CUDA_OK(cudaStreamWaitEvent(aux, ev_fork, 0));
    long long tgrid = (long long)dv->sms * dv->table_per_sm;
    if (tgrid > list_cap) tgrid = list_cap;
    k3_box_depth_table<<<(int)tgrid, kTableThreads, kTableSmem, aux>>>(
        depth, Hp, Wp, boxes, cap, d, scale, inflated, workspace, list_cap);
    MakeMapSrc mk;
    mk.depth = depth; mk.Hp = Hp; mk.Wp = Wp;
    k3_box_depth_cta<kHugeThreads, kHugeStage, MakeMapSrc><<<dv->sms, kHugeThreads, sizeof(float) * kHugeStage, aux>>>(
        mk, Hp, Wp, boxes, cap, d, scale, inflated, workspace + 1, workspace + 8 + 2 * list_cap);
    ST_CUDA_OK(cudaEventRecord(ev_join, aux));
    {
        long long grid = (slots + kWarps - 1) / kWarps;
        if (grid > dv->warp_ctas) grid = dv->warp_ctas;
        k3_box_depth<<<(int)grid, kWarps * 32, 0, st>>>(depth, Hp, Wp, boxes, cap, d, scale, inflated, workspace, list_cap);
    }
    ST_CUDA_OK(cudaStreamWaitEvent(st, ev_join, 0));
    return st_last_launch();
}

extern "C" int st_box_depth_mode(const float *depth, int S, int Hp, int Wp, const float *boxes,
                                 const int32_t *counts, int cap, int mode, float *d, float *scale,
                                 float *inflated, int32_t *oob, int32_t *workspace, void *stream) {
    if (mode == ST_DEPTH_WINDOW) {
        if (oob && S > 0) ST_CUDA_OK(cudaMemsetAsync(oob, 0, sizeof(int32_t) * S, (cudaStream_t)stream));
        return st_box_depth_reuse(depth, S, Hp, Wp, boxes, counts, cap, d, scale, inflated, workspace, nullptr,
                                  nullptr, nullptr, stream);
    }
    if (!depth || !boxes || !counts || !d || !scale || S < 0 || cap < 1 || Hp < 1 || Wp < 1 ||
        mode < ST_DEPTH_TRUNCATED_MEAN || mode > ST_DEPTH_CENTER)
        return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (oob) ST_CUDA_OK(cudaMemsetAsync(oob, 0, sizeof(int32_t) * S, st));
    k3_box_depth_ablation<<<dim3(cap < 64 ? cap : 64, S), kAblationThreads, 0, st>>>(
        depth, Hp, Wp, boxes, counts, cap, mode, d, scale, inflated, oob);
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
