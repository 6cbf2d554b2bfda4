// K3 on the RAW u16 disparity frame: the per-box depth statistic of
// OCSORT_Disparity.extract_depth (mmtrack/models/mot/ocsort_disparity.py:136-175) without a
// float32 depth map in between.  st_box_depth_raw(raw, ...) returns bit for bit what
// st_box_depth returns on the map st_disp2depth writes for the same frame (resize factor 1).
//
// Why it can: depth is a non-increasing function of the raw code c (Z = B f / (c / 16 + 1e-6),
// two correctly rounded float32 operations), so the sorted list of valid depths of a box is the
// multiset of its valid codes in DESCENDING code order.  Every order statistic the reference
// takes from the sorted list (the median, the two ends of the corner-voted 60 % window) is a
// rank in a histogram of codes, and the window mean is sum(depth(code) * multiplicity inside
// the window) / length.  The map's pad region and code 65535 are disparity 0 = the pad depth,
// which the `< 150` filter drops and the 2 x 2 corner means see as they see it in the map.
// The frame is half the bytes of the map and L2-sized, and K1's 4 H' W' byte write leaves the
// per-frame path unless the caller wants the map.
//
// Kernels (worklists by k3_classify, k3_select.cuh):
//   k3r_box_depth       one WARP per box of <= 512 pixels: valid codes staged as 16-bit keys in
//                       shared memory, ranks by an MSB-first bit-wise select (<= 16 passes, one
//                       REDUX each), window sum as one masked pass
//   k3r_box_depth_hist  a thread-block CLUSTER of 4 CTAs per larger box: each CTA histograms a
//                       quarter of the rows into its own 4096-bin shared-memory histogram (64-bit
//                       loads of 4 pixels, run-length + match.any aggregated atomics), the
//                       histograms are summed over distributed shared memory (every CTA reduces a
//                       quarter of the bins into the leader), the leader scans the bins in
//                       descending code order and takes ranks and window sum from the scan
//   k3_box_depth_cta    (k3_select.cuh, RawSrc) boxes with a valid code >= 4096 (closer than
//                       ~0.63 m at B f = 160): the general radix-select path
#include <cstdio>
#include <cstdlib>
#include <cooperative_groups.h>
#include "k3_select.cuh"

namespace cg = cooperative_groups;
using namespace stk3;

namespace {

constexpr int kWarps = 4;                // boxes per CTA of the warp kernel
constexpr int kStage = 512;              // keys staged per warp
constexpr int kSmallPix = 512;           // boxes up to this many pixels: one warp; larger: one cluster
constexpr int kBins = 4096;              // codes with a direct histogram bin
constexpr int kHistThreads = 512;
constexpr int kCluster = 4;

// MSB-first binary search on 16-bit keys staged in shared memory (one warp, no atomics): per bit
// one pass and one REDUX.  value[k] = key of rank rank_in[k] in ascending key order, below[k] =
// number of keys strictly smaller.
__device__ void bit_select2_u16(const uint16_t *staged, int n, int lane, int nsel, const int *rank_in,
                                uint32_t *value, int *below) {
    uint32_t vor = 0u, vand = 0xffffu;
    for (int i = lane; i < n; i += 32) { const uint32_t u = staged[i]; vor |= u; vand &= u; }
    vor = __reduce_or_sync(0xffffffffu, vor);
    vand = __reduce_and_sync(0xffffffffu, vand);
    const uint32_t diff = vor ^ vand;
    const int top = diff ? 31 - __clz(diff) : -1;          // highest bit that varies
    const uint32_t common = top < 0 ? vand : (vand & ~((2u << top) - 1u));
    uint32_t prefix[2] = {common, common};
    int rank[2] = {rank_in[0], nsel > 1 ? rank_in[1] : 0};
    int less[2] = {0, 0};
    for (int bit = top; bit >= 0; --bit) {
        const uint32_t hi = ~((2u << bit) - 1u);            // bits above `bit`
        int c0 = 0, c1 = 0;
        for (int i = lane; i < n; i += 32) {
            const uint32_t u = staged[i];
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

// one small box by one warp.  Keys are 65535 - code, so ascending key = ascending depth.
__device__ void k3r_box_one(const RawSrc &src, int Hp, int Wp, const float *__restrict__ boxes, const int s,
                            const int bi, int cap, float *__restrict__ d_out, float *__restrict__ scale_out,
                            float *__restrict__ inflated, uint16_t *staged, const int lane) {
    const long long o = (long long)s * cap + bi;
    const float4 box = reinterpret_cast<const float4 *>(boxes)[o];
    const long long b0 = trunc_i64(box.x), b1 = trunc_i64(box.y), b2 = trunc_i64(box.z), b3 = trunc_i64(box.w);
    const Slice ys = py_slice(b1, b3, Hp), xs = py_slice(b0, b2, Wp);
    const unsigned bw = (unsigned)max(xs.hi - xs.lo, 0);
    const unsigned npix = bw * (unsigned)max(ys.hi - ys.lo, 0);
    // corner patches: lane 4c+e fetches element e of corner c now (hidden behind the staging pass)
    float corner_val = 0.f;
    int corner_cnt = 0;
    if (lane < 16) {
        int cy, cx;
        if (corner_elem(lane >> 2, lane & 3, b0, b1, b2, b3, Hp, Wp, cy, cx, corner_cnt))
            corner_val = src.at(cy, cx);
    }
    // pass 0: stage the keys of the valid pixels (npix <= kSmallPix = kStage always fits)
    int n = 0;
    for (unsigned i0 = 0; i0 < npix; i0 += 128) {
        uint32_t cd[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {              // four independent loads in flight
            const unsigned i = i0 + u * 32 + lane;
            cd[u] = 65535u;
            if (i < npix) {
                const unsigned ry = i / bw, rx = i - ry * bw;
                cd[u] = src.code(ys.lo + (int)ry, xs.lo + (int)rx);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool ok = cd[u] != 65535u && depth_valid(st_code_depth(cd[u], src.bf));
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            const int pos = n + __popc(m & ((1u << lane) - 1u));
            if (ok && pos < kStage) staged[pos] = (uint16_t)(65535u - cd[u]);
            n += __popc(m);
        }
    }
    __syncwarp();
    // (box[2]-box[0]) > 800 in int64; wraps like NumPy for the INT64_MIN case
    const long long wbox = wrap_add(b2, (long long)(0ull - (unsigned long long)b0));

    float d = -1.f, scale = 1.f;
    if (!(n < 1 || wbox > 800)) {
        uint32_t val[2];
        int below[2];
        int rk[2] = {n / 2, 0};
        bit_select2_u16(staged, n, lane, 1, rk, val, below);
        const float d_mid = st_code_depth(65535u - val[0], src.bf);
        // corner votes on the unfiltered map (:157-160): lane 4c gathers the elements of corner c
        int k = 0;
        {
            float cv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) cv[e] = __shfl_sync(0xffffffffu, corner_val, (lane & 12) + e);
            const bool vote = lane < 16 && (lane & 3) == 0 && corner_mean_of(cv, corner_cnt) > d_mid;
            k = __popc(__ballot_sync(0xffffffffu, vote));
        }
        const Slice seg = window_of(n, k);
        const int m = seg.hi - seg.lo;
        if (m <= 0) {
            d = __int_as_float(0x7fc00000);                    // mean of empty -> NaN
        } else {
            rk[0] = seg.lo; rk[1] = seg.hi - 1;
            bit_select2_u16(staged, n, lane, 2, rk, val, below);
            const uint32_t ua = val[0], ub = val[1];
            const double da = (double)st_code_depth(65535u - ua, src.bf), db = (double)st_code_depth(65535u - ub, src.bf);
            double sum;
            if (ua == ub) {
                sum = da * (double)m;
            } else {
                double part = 0.0;
                int eq_a = 0;
                for (int i = lane; i < n; i += 32) {
                    const uint32_t u = staged[i];
                    if (u > ua && u < ub) part += (double)st_code_depth(65535u - u, src.bf);
                    eq_a += (u == ua);
                }
                sum = warp_sum(part);
                const int n_eq_a = __reduce_add_sync(0xffffffffu, eq_a);
                const int in_a = below[0] + n_eq_a - seg.lo;   // duplicates of s[lo] inside
                const int in_b = seg.hi - below[1];            // duplicates of s[hi-1] inside
                sum += da * (double)in_a + db * (double)in_b;
            }
            d = (float)(sum / (double)m);
        }
        scale = scale_of(d);
    }
    if (lane == 0) write_box(o, box, d, scale, d_out, scale_out, inflated);
}

__global__ void __launch_bounds__(kWarps * 32)
k3r_box_depth(MakeRawSrc mk, int Hp, int Wp, const float *__restrict__ boxes, int cap,
              float *__restrict__ d_out, float *__restrict__ scale_out, float *__restrict__ inflated,
              const int32_t *__restrict__ worklist, int list_cap) {
    __shared__ uint16_t staged_all[kWarps][kStage];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = worklist[3];
    const int32_t *list = worklist + 8 + 4 * list_cap;
    for (int idx = blockIdx.x * kWarps + warp; idx < n; idx += gridDim.x * kWarps) {
        const int s = list[2 * idx];
        const RawSrc src = mk(s);
        k3r_box_one(src, Hp, Wp, boxes, s, list[2 * idx + 1], cap, d_out, scale_out, inflated, staged_all[warp], lane);
        __syncwarp();
    }
}

// four consecutive codes of row y starting at column x0 (x0 % 4 == 0): one 64-bit load when the
// frame allows it; pixels outside the H x W frame are 65535 (disparity 0, the pad)
__device__ __forceinline__ void load4(const RawSrc &src, bool vec_ok, int y, int x0, uint32_t *c) {
    c[0] = c[1] = c[2] = c[3] = 65535u;
    if (y >= src.H || x0 >= src.W) return;
    const uint16_t *p = src.raw + (long long)y * src.W + x0;
    if (vec_ok && x0 + 3 < src.W) {
        const uint2 v = *reinterpret_cast<const uint2 *>(p);
        c[0] = v.x & 0xffffu; c[1] = v.x >> 16; c[2] = v.y & 0xffffu; c[3] = v.y >> 16;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (x0 + k < src.W) c[k] = p[k];
    }
}

// Histogram pass over rows [r0, r1) x columns xs of one box by the kT threads of a CTA (every
// thread must call it: the trip count is block-uniform).  Codes below kBins are counted in
// `hist` (shared memory); returns true when this thread saw an in-box code without a bin.
template <int kT>
__device__ __forceinline__ bool hist_rows(const RawSrc &src, bool vec_ok, Slice xs, int r0, int r1, uint32_t *hist) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int xa = xs.lo & ~3;
    const int g = xs.hi > xa ? (xs.hi - xa + 3) >> 2 : 0;          // 4-pixel groups per row
    const int total = g * max(r1 - r0, 0);
    bool ovf = false;
    for (int base = 0; base < total; base += 4 * kT) {
        uint32_t c[4][4];
        int x0[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                 // four 64-bit loads in flight
            const int q = base + u * kT + tid;
            x0[u] = -(1 << 20);
            c[u][0] = c[u][1] = c[u][2] = c[u][3] = 65535u;
            if (q < total) {
                const int ry = q / g, gx = q - ry * g;
                x0[u] = xa + 4 * gx;
                load4(src, vec_ok, r0 + ry, x0[u], c[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            // histogram key of each pixel (none outside the box or without a bin), runs of
            // equal keys inside the group merged; then one shared-memory atomic per distinct
            // (key, run length) among the lanes of the warp: object pixels share one code
            uint32_t key[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int x = x0[u] + k;
                const bool inside = x >= xs.lo && x < xs.hi;
                key[k] = (inside && c[u][k] < (uint32_t)kBins) ? c[u][k] : 0xffffu;
                ovf |= inside && c[u][k] >= (uint32_t)kBins && c[u][k] != 65535u;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool start = key[k] != 0xffffu && (k == 0 || key[k] != key[k - 1]);
                if (!__any_sync(0xffffffffu, start)) continue;
                int len = 1;
#pragma unroll
                for (int j = k + 1; j < 4; ++j) len += (len == j - k && key[j] == key[k]) ? 1 : 0;
                const uint32_t tag = start ? (key[k] | ((uint32_t)len << 16)) : 0xffffffffu;
                const unsigned peers = __match_any_sync(0xffffffffu, tag);
                if (start && lane == __ffs(peers) - 1) atomicAdd(&hist[key[k]], (uint32_t)(len * __popc(peers)));
            }
        }
    }
    return ovf;
}

// The statistic from the complete histogram, by the kT threads of one CTA: bins in DESCENDING code
// order = ascending depth; thread t owns positions [B t, B t + B), position p holds code kBins-1-p.
template <int kT>
__device__ __forceinline__ void hist_stats(const uint32_t *hist, float bf, const float *corner_sm, const int *corner_n,
                                           int *wsum, double *dred, float *dmid_sh, float &d, float &scale) {
    constexpr int B = kBins / kT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int cnt[B];
    float dep[B];
    int tot = 0;
    const int c_hi = kBins - 1 - B * tid;                           // code of this thread's first position
#pragma unroll
    for (int q = 0; q < B / 4; ++q) {
        const uint4 h4 = *reinterpret_cast<const uint4 *>(&hist[c_hi - 4 * q - 3]);   // codes c_hi-4q-3 .. c_hi-4q
        const uint32_t hv[4] = {h4.w, h4.z, h4.y, h4.x};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            dep[4 * q + j] = st_code_depth((uint32_t)(c_hi - 4 * q - j), bf);
            cnt[4 * q + j] = depth_valid(dep[4 * q + j]) ? (int)hv[j] : 0;
            tot += cnt[4 * q + j];
        }
    }
    // block-wide exclusive scan of the per-thread totals
    int inc = tot;
    for (int of = 1; of < 32; of <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, of);
        if (lane >= of) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int excl = inc - tot, n = 0;
    for (int w = 0; w < kT / 32; ++w) {
        const int v = wsum[w];
        if (w < warp) excl += v;
        n += v;
    }
    d = -1.f; scale = 1.f;
    if (n >= 1) {                                                   // block-uniform
        // the median: the thread whose rank range holds n / 2
        const int rmid = n / 2;
        if (excl <= rmid && rmid < excl + tot) {
            int acc = excl;
#pragma unroll
            for (int j = 0; j < B; ++j) {
                if (acc <= rmid && rmid < acc + cnt[j]) *dmid_sh = dep[j];
                acc += cnt[j];
            }
        }
        __syncthreads();
        const float d_mid = *dmid_sh;
        int k = 0;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) k += corner_mean_of(corner_sm + 4 * cc, corner_n[cc]) > d_mid;
        const Slice seg = window_of(n, k);
        const int m = seg.hi - seg.lo;
        if (m <= 0) {
            d = __int_as_float(0x7fc00000);                         // mean of empty -> NaN
        } else {
            double part = 0.0;
            int acc = excl;
#pragma unroll
            for (int j = 0; j < B; ++j) {
                const int lo = max(acc, seg.lo), hi = min(acc + cnt[j], seg.hi);
                if (hi > lo) part += (double)dep[j] * (double)(hi - lo);
                acc += cnt[j];
            }
            const double sum = block_sum(part, dred);
            d = (float)(sum / (double)m);
        }
        scale = scale_of(d);
    }
    __syncthreads();                                                // wsum / dmid_sh are reused by the next item
}

// Medium boxes (kSmallPix < pixels <= kMediumPix): one CTA of 256 threads per box, many CTAs per SM.
constexpr int kMedThreads = 256;
constexpr int kMediumPix = 8192;

__global__ void __launch_bounds__(kMedThreads, 4)
k3r_box_depth_med(MakeRawSrc mk, int Hp, int Wp, const float *__restrict__ boxes, int cap,
                  float *__restrict__ d_out, float *__restrict__ scale_out, float *__restrict__ inflated,
                  int32_t *__restrict__ worklist, int list_cap) {
    __shared__ __align__(16) uint32_t hist[kBins];
    __shared__ float dmid_sh;
    __shared__ float corner_sm[16];
    __shared__ int corner_n[4];
    __shared__ int wsum[kMaxCtaWarps];
    __shared__ double dred[kMaxCtaWarps];
    const int tid = threadIdx.x;
    const int n_work = worklist[0];
    const bool vec_ok = (mk.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(mk.base()) & 7u) == 0);
    for (int item = blockIdx.x; item < n_work; item += gridDim.x) {
        const int s = worklist[8 + 2 * item], bi = worklist[9 + 2 * item];
        const long long o = (long long)s * cap + bi;
        const float4 box = reinterpret_cast<const float4 *>(boxes)[o];
        const long long b0 = trunc_i64(box.x), b1 = trunc_i64(box.y), b2 = trunc_i64(box.z), b3 = trunc_i64(box.w);
        const RawSrc src = mk(s);
        const Slice ys = py_slice(b1, b3, Hp), xs = py_slice(b0, b2, Wp);
        const long long wbox = wrap_add(b2, (long long)(0ull - (unsigned long long)b0));
        float d = -1.f, scale = 1.f;
        if (!(wbox > 800)) {                              // reference: d = -1, scale = 1 (:150)
            for (int i = tid; i < kBins; i += kMedThreads) hist[i] = 0u;
            if (tid < 16) {                               // corner patches (read after the barriers below)
                int cy, cx, cn;
                const bool has = corner_elem(tid >> 2, tid & 3, b0, b1, b2, b3, Hp, Wp, cy, cx, cn);
                corner_sm[tid] = has ? src.at(cy, cx) : 0.f;
                if ((tid & 3) == 0) corner_n[tid >> 2] = cn;
            }
            __syncthreads();
            const bool ovf = hist_rows<kMedThreads>(src, vec_ok, xs, ys.lo, ys.hi, hist);
            if (__syncthreads_or(ovf)) {                  // a valid code without a bin: general path
                if (tid == 0) {
                    const int slot = atomicAdd(&worklist[1], 1);
                    worklist[8 + 2 * list_cap + 2 * slot] = s;
                    worklist[9 + 2 * list_cap + 2 * slot] = bi;
                }
                continue;                                 // block-uniform
            }
            hist_stats<kMedThreads>(hist, src.bf, corner_sm, corner_n, wsum, dred, &dmid_sh, d, scale);
        }
        if (tid == 0) write_box(o, box, d, scale, d_out, scale_out, inflated);
    }
}

// Huge boxes (> kMediumPix pixels): a cluster of kCluster CTAs per box, rows split over the CTAs,
// histograms merged over distributed shared memory.  The huge list grows from the END of the
// large-box list region (entry k at list_cap - 1 - k), its count is worklist[4], its cursor [2].
__global__ void __launch_bounds__(kHistThreads, 2)
k3r_box_depth_hist(MakeRawSrc mk, int Hp, int Wp, const float *__restrict__ boxes, int cap,
                   float *__restrict__ d_out, float *__restrict__ scale_out, float *__restrict__ inflated,
                   int32_t *__restrict__ worklist, int list_cap) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned crank = cluster.block_rank();
    __shared__ __align__(16) uint32_t hist[kBins];
    __shared__ int next_item, overflow_sh;
    __shared__ float dmid_sh;
    __shared__ float corner_sm[16];
    __shared__ int corner_n[4];
    __shared__ int wsum[kMaxCtaWarps];
    __shared__ double dred[kMaxCtaWarps];
    const int tid = threadIdx.x;
    int *leader_next = cluster.map_shared_rank(&next_item, 0);
    int *leader_ovf = cluster.map_shared_rank(&overflow_sh, 0);
    const int n_work = worklist[4];
    if (n_work == 0) return;                              // uniform over the grid: nothing to synchronise
    const bool vec_ok = (mk.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(mk.base()) & 7u) == 0);
    while (true) {
        // items come from a shared cursor (boxes range from 8 193 to 500 000+ pixels)
        if (crank == 0 && tid == 0) { next_item = atomicAdd(&worklist[2], 1); overflow_sh = 0; }
        for (int i = tid; i < kBins; i += kHistThreads) hist[i] = 0u;
        cluster.sync();                                   // (A) item published, histograms cleared
        const int item = *leader_next;
        if (item >= n_work) { cluster.sync(); break; }    // nobody leaves while its shared memory may be read
        const int e = list_cap - 1 - item;
        const int s = worklist[8 + 2 * e], bi = worklist[9 + 2 * e];
        const long long o = (long long)s * cap + bi;
        const float4 box = reinterpret_cast<const float4 *>(boxes)[o];
        const long long b0 = trunc_i64(box.x), b1 = trunc_i64(box.y), b2 = trunc_i64(box.z), b3 = trunc_i64(box.w);
        const RawSrc src = mk(s);
        const Slice ys = py_slice(b1, b3, Hp), xs = py_slice(b0, b2, Wp);
        const long long wbox = wrap_add(b2, (long long)(0ull - (unsigned long long)b0));
        const bool skip = wbox > 800;                     // reference: d = -1, scale = 1 (:150)
        if (!skip) {
            if (crank == 0 && tid < 16) {                 // corner patches (read after the barriers below)
                int cy, cx, cn;
                const bool has = corner_elem(tid >> 2, tid & 3, b0, b1, b2, b3, Hp, Wp, cy, cx, cn);
                corner_sm[tid] = has ? src.at(cy, cx) : 0.f;
                if ((tid & 3) == 0) corner_n[tid >> 2] = cn;
            }
            // this CTA's rows of the box
            const int bh = max(ys.hi - ys.lo, 0);
            const int rows_per = (bh + kCluster - 1) / kCluster;
            const int r0 = ys.lo + (int)crank * rows_per, r1 = min(r0 + rows_per, ys.hi);
            if (hist_rows<kHistThreads>(src, vec_ok, xs, r0, r1, hist)) *leader_ovf = 1;
        }
        cluster.sync();                                   // (B) every CTA's histogram is complete
        if (!skip) {
            // distributed merge: this CTA sums its quarter of the bins over the cluster into the leader
            uint32_t *rh[kCluster];
#pragma unroll
            for (int r = 0; r < kCluster; ++r) rh[r] = cluster.map_shared_rank(hist, r);
            for (int b = (int)crank * (kBins / kCluster) + tid; b < ((int)crank + 1) * (kBins / kCluster); b += kHistThreads) {
                uint32_t acc = 0u;
#pragma unroll
                for (int r = 0; r < kCluster; ++r) acc += rh[r][b];
                rh[0][b] = acc;
            }
        }
        cluster.sync();                                   // (C) merged histogram in the leader
        if (crank != 0) continue;
        float d = -1.f, scale = 1.f;
        if (!skip) {
            if (overflow_sh) {                            // a valid code without a bin: general path
                if (tid == 0) {
                    const int slot = atomicAdd(&worklist[1], 1);
                    worklist[8 + 2 * list_cap + 2 * slot] = s;
                    worklist[9 + 2 * list_cap + 2 * slot] = bi;
                }
                continue;                                 // block-uniform
            }
            hist_stats<kHistThreads>(hist, src.bf, corner_sm, corner_n, wsum, dred, &dmid_sh, d, scale);
        }
        if (tid == 0) write_box(o, box, d, scale, d_out, scale_out, inflated);
    }
}

struct K3rDev { int sms, warp_ctas, med_ctas, hist_clusters; };
StPerDevice<K3rDev> g_k3rdev;

int launch_hist(const K3rDev *dv, cudaStream_t st, MakeRawSrc mk, int Hp, int Wp, const float *boxes, int cap, float *d,
                float *scale, float *inflated, int32_t *workspace, int list_cap) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(dv->hist_clusters * kCluster), 1, 1);
    cfg.blockDim = dim3(kHistThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k3r_box_depth_hist, mk, Hp, Wp, boxes, cap, d, scale, inflated, workspace,
                              list_cap) == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

}  // namespace

extern "C" int st_box_depth_raw(const uint16_t *raw, int S, int H, int W, int Hp, int Wp, float baseline,
                                float focal_length, const float *boxes, const int32_t *counts, int cap,
                                float *d, float *scale, float *inflated, int32_t *workspace,
                                const float *prev_boxes, const int32_t *prev_index, const float *prev_d,
                                void *stream) {
    return st_box_depth_raw_tab(raw, nullptr, S, H, W, Hp, Wp, baseline, focal_length, boxes, counts, cap, d, scale,
                                inflated, workspace, prev_boxes, prev_index, prev_d, stream);
}

int st_box_depth_raw_tab(const uint16_t *raw, const StFrameInputs *tab, int S, int H, int W, int Hp, int Wp,
                         float baseline, float focal_length, const float *boxes, const int32_t *counts, int cap,
                         float *d, float *scale, float *inflated, int32_t *workspace, const float *prev_boxes,
                         const int32_t *prev_index, const float *prev_d, void *stream) {
    if (!raw || !boxes || !counts || !d || !scale || !workspace || S < 0 || cap < 1 || H < 1 || W < 1 || Hp < H ||
        Wp < W)
        return ST_ERR_INVALID_ARG;
    if (prev_index && (!prev_boxes || !prev_d)) return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ST_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(int32_t) * 8, st));
    const int list_cap = S * cap;
    const K3rDev *dv = g_k3rdev.get([](K3rDev &v) {
        v.sms = st_ctx_sms();
        if (cudaFuncSetAttribute(k3_box_depth_cta<kHugeThreads, kHugeStage, MakeRawSrc>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(float) * kHugeStage)) != cudaSuccess) return false;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3r_box_depth, kWarps * 32, 0) != cudaSuccess || per_sm < 1)
            per_sm = 8;
        v.warp_ctas = v.sms * per_sm;
        per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k3r_box_depth_med, kMedThreads, 0) != cudaSuccess || per_sm < 1)
            per_sm = 4;
        v.med_ctas = v.sms * per_sm;
        // resident clusters of the histogram kernel
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(v.sms * kCluster), 1, 1);
        cfg.blockDim = dim3(kHistThreads, 1, 1);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&clusters, k3r_box_depth_hist, &cfg) != cudaSuccess || clusters < 1) {
            cudaGetLastError();
            clusters = v.sms / kCluster;
        }
        if (const char *e = getenv("ST_K3R_CLUSTERS")) { const int c = atoi(e); if (c > 0 && c < clusters) clusters = c; }
        v.hist_clusters = clusters;
        if (getenv("ST_K3R_VERBOSE"))
            fprintf(stderr, "[st] k3r: sms %d, warp CTAs %d, medium CTAs %d, resident clusters of %d: %d\n", v.sms,
                    v.warp_ctas, v.med_ctas, kCluster, clusters);
        return true;
    });
    if (!dv) return ST_ERR_CUDA;
    cudaStream_t aux;
    cudaEvent_t ev_fork, ev_join;
    if (st_ctx_aux(st, &aux, &ev_fork, &ev_join) != ST_OK) return ST_ERR_CUDA;
    MakeRawSrc mk;
    mk.raw = raw; mk.tab = tab; mk.H = H; mk.W = W; mk.bf = st_bf(baseline, focal_length);
    const long long slots = (long long)S * cap;
    k3_classify<<<(int)((slots + 255) / 256), 256, 0, st>>>(S, Hp, Wp, boxes, counts, cap, d, scale, inflated, workspace,
                                                           list_cap, prev_boxes, prev_index, prev_d, (unsigned)kSmallPix,
                                                           (unsigned)kMediumPix);
    // fork: the medium boxes (one CTA each) on the auxiliary stream; the huge boxes (clusters), the
    // overflow list of both and the small boxes (warps) on the caller's stream; join before returning
    ST_CUDA_OK(cudaEventRecord(ev_fork, st));
    ST_CUDA_OK(cudaStreamWaitEvent(aux, ev_fork, 0));
    {
        long long grid = dv->med_ctas;
        if (grid > list_cap) grid = list_cap;
        k3r_box_depth_med<<<(int)grid, kMedThreads, 0, aux>>>(mk, Hp, Wp, boxes, cap, d, scale, inflated, workspace, list_cap);
    }
    ST_CUDA_OK(cudaEventRecord(ev_join, aux));
    {
        long long grid = (slots + kWarps - 1) / kWarps;
        if (grid > dv->warp_ctas) grid = dv->warp_ctas;
        k3r_box_depth<<<(int)grid, kWarps * 32, 0, st>>>(mk, Hp, Wp, boxes, cap, d, scale, inflated, workspace, list_cap);
    }
    if (launch_hist(dv, st, mk, Hp, Wp, boxes, cap, d, scale, inflated, workspace, list_cap) != ST_OK) return ST_ERR_CUDA;
    ST_CUDA_OK(cudaStreamWaitEvent(st, ev_join, 0));
    k3_box_depth_cta<kHugeThreads, kHugeStage, MakeRawSrc><<<dv->sms, kHugeThreads, sizeof(float) * kHugeStage, st>>>(
        mk, Hp, Wp, boxes, cap, d, scale, inflated, workspace + 1, workspace + 8 + 2 * list_cap);
    return st_last_launch();
}
