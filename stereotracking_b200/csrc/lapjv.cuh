// Warp-cooperative dense Jonker-Volgenant assignment on the "extended" matrix
// that lap.lapjv(extend_cost=True, cost_limit=L) builds, without materialising
// it: ext[i][j] = cost[i][j] (i<R, j<C), 0 (i>=R, j>=C), L/2 otherwise.
//
// Replaces lap.lapjv at mmtrack/models/trackers/ocsort_tracker_disparity.py
// :260-261 and :312-313.  The phase order and every scan order of the dense JV
// solver are kept (column reduction with reverse sweep, reduction transfer, two
// augmenting-row-reduction passes, shortest-path augmentation with
// READY/SCAN/TODO lists) so that exact cost ties resolve as in the reference;
// what is parallel is only what is order-independent: minima with index
// tie-breaks, relaxations of disjoint columns, and detection of the columns a
// sequential sweep would move (the moves themselves are replayed in order).
#pragma once
#include "st_common.cuh"

namespace stlap {

constexpr double kBig = 1.0e300;

// All arrays have n = R + C entries and live in shared memory.  They are
// volatile because lanes of one warp hand data to each other through them
// between __syncwarp() calls (lane 0 replays serial list updates that the other
// lanes then read).
struct Scratch {
    volatile double *v;            // column potentials
    volatile double *d;            // shortest-path distances
    volatile int *cols, *pred, *x, *y, *free_rows;
    volatile unsigned *mask_a, *mask_b;   // (n+31)/32 words each
    volatile unsigned char *unique;
};

__host__ __device__ inline size_t scratch_bytes(int n) {
    size_t words = (size_t)(n + 31) / 32;
    size_t b = sizeof(double) * 2 * n + sizeof(int) * 5 * n + sizeof(unsigned) * 2 * words + n;
    return (b + 15) & ~(size_t)15;
}

__device__ inline Scratch carve(unsigned char *p, int n) {
    Scratch s;
    size_t words = (size_t)(n + 31) / 32;
    s.v = reinterpret_cast<volatile double *>(p); p += sizeof(double) * n;
    s.d = reinterpret_cast<volatile double *>(p); p += sizeof(double) * n;
    s.cols = reinterpret_cast<volatile int *>(p); p += sizeof(int) * n;
    s.pred = reinterpret_cast<volatile int *>(p); p += sizeof(int) * n;
    s.x = reinterpret_cast<volatile int *>(p); p += sizeof(int) * n;
    s.y = reinterpret_cast<volatile int *>(p); p += sizeof(int) * n;
    s.free_rows = reinterpret_cast<volatile int *>(p); p += sizeof(int) * n;
    s.mask_a = reinterpret_cast<volatile unsigned *>(p); p += sizeof(unsigned) * words;
    s.mask_b = reinterpret_cast<volatile unsigned *>(p); p += sizeof(unsigned) * words;
    s.unique = p;
    return s;
}

struct Problem {
    const float *cost;    // R x C row-major (float32 as the reference's dists)
    int R, C, n;
    double half;          // cost_limit / 2
    int *dbg;             // optional [8] progress words (global), may be null
    int no_greedy;        // diagnostics: != 0 = every non-isolated row goes through the search (no greedy start)
    int *why;             // optional [8] diagnostics of fast_assign_warp (global), may be null: [0] code of the
                          // last refusal, [1] refusals, [2..4] cycles of start / search / certificate of the last
                          // call, [5] rows searched, [6] active rows, [7] tight arcs
    // Admissible-edge statistics of the fast path (cost < L = 2 * half): number of
    // admissible entries per row / column, one admissible column of each row, and
    // adm_misc[0] != 0 when the matrix holds a NaN, adm_misc[1] != 0 when a cost lies within
    // 2e-9 of L.  Filled by whoever wrote the
    // matrix (the track kernel does it while it computes the costs) or by
    // count_admissible().
    int *row_cnt, *col_cnt, *row_adm, *adm_misc;
    __device__ __forceinline__ void mark(int slot, int v) const {
        if (dbg && (threadIdx.x & 31) == 0) ((volatile int *)dbg)[slot] = v;
    }
    __device__ __forceinline__ double at(int i, int j) const {
        if (i < R) return j < C ? (double)cost[(long long)i * C + j] : half;
        return j < C ? half : 0.0;
    }
};

__device__ __forceinline__ double warp_min(double v) {
    for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// (value, index) pairs ordered lexicographically; index < 0 = "none"
__device__ __forceinline__ bool pair_less(double va, int ia, double vb, int ib) {
    if (ib < 0) return ia >= 0;
    if (ia < 0) return false;
    return va < vb || (va == vb && ia < ib);
}

// Solve; on return s.x[i] (row->col) and s.y[j] (col->row) hold the extended
// assignment.  Must be called by all 32 lanes of one warp.  Returns false when
// an iteration guard fired (non-finite costs).
// __noinline__ on purpose: with nvcc 12.9 the copies inlined into the fused
// track kernel mis-executed the warp collectives of the augmentation loop on
// inputs the stand-alone kernel solved (bisected on a B200); one
// shared out-of-line body behaves identically at every call site.
static __device__ __noinline__ bool solve_warp(const Problem &P, Scratch &s) {
    const int lane = threadIdx.x & 31;
    const int n = P.n;
    bool ok = true;

    P.mark(3, 1);
    // ---- column reduction -----------------------------------------------------
    for (int i = lane; i < n; i += 32) { s.x[i] = -1; s.unique[i] = 1; }
    for (int j = lane; j < n; j += 32) {
        double best = kBig;
        int arg = 0;
        if (j < P.C) {
            for (int i = 0; i < P.R; ++i) {
                double c = (double)P.cost[(long long)i * P.C + j];
                if (c < best) { best = c; arg = i; }
            }
            if (P.n > P.R && P.half < best) { best = P.half; arg = P.R; }
        } else {
            if (P.R > 0) { best = P.half; arg = 0; }
            if (P.n > P.R && 0.0 < best) { best = 0.0; arg = P.R; }
        }
        s.v[j] = best;
        s.y[j] = arg;
    }
    __syncwarp();
    // reverse sweep: the highest column that elected a row keeps it
    for (int j = lane; j < n; j += 32) atomicMax(const_cast<int *>(&s.x[s.y[j]]), j);
    __syncwarp();
    for (int j = lane; j < n; j += 32) {
        int i = s.y[j];
        if (s.x[i] != j) { s.unique[i] = 0; s.y[j] = -1; }
    }
    __syncwarp();

    P.mark(3, 2);
    // ---- reduction transfer (row order matters: v[] feeds later rows) ----------
    int nfree = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int ii = i0 + lane;
        const int xi_l = ii < n ? s.x[ii] : 0;
        const bool uq_l = ii < n ? (s.unique[ii] != 0) : false;
        const unsigned freeb = __ballot_sync(0xffffffffu, ii < n && xi_l < 0);
        const unsigned workb = __ballot_sync(0xffffffffu, ii < n && xi_l >= 0 && uq_l);
        unsigned todo = freeb | workb;
        while (todo) {
            const int b = __ffs(todo) - 1;
            todo &= todo - 1;
            const int i = i0 + b;
            if ((freeb >> b) & 1u) {
                if (lane == 0) s.free_rows[nfree] = i;
                ++nfree;
            } else {
                const int j1 = __shfl_sync(0xffffffffu, xi_l, b);
                double m = kBig;
                for (int j2 = lane; j2 < n; j2 += 32)
                    if (j2 != j1) m = fmin(m, P.at(i, j2) - s.v[j2]);
                m = warp_min(m);
                if (lane == 0) s.v[j1] -= m;
                __syncwarp();
            }
        }
    }
    __syncwarp();

    P.mark(3, 3); P.mark(4, nfree);
    // ---- augmenting row reduction, two passes --------------------------------------
    for (int pass = 0; pass < 2 && nfree > 0; ++pass) {
        int cur = 0, nxt = 0;
        unsigned long long visits = 0, guard = 0;
        const unsigned long long guard_max = 64ull * (unsigned long long)n * (unsigned long long)n + 1024ull;
        while (cur < nfree) {
            if (++guard > guard_max) { ok = false; break; }
            ++visits;
            const int fr = s.free_rows[cur++];
            // two smallest reduced costs, lowest index first on ties
            double u1 = kBig, u2 = kBig;
            int j1 = -1, j2 = -1;
            for (int j = lane; j < n; j += 32) {
                const double r = P.at(fr, j) - s.v[j];
                if (pair_less(r, j, u1, j1)) { u2 = u1; j2 = j1; u1 = r; j1 = j; }
                else if (pair_less(r, j, u2, j2)) { u2 = r; j2 = j; }
            }
            for (int o = 16; o; o >>= 1) {
                const double ou1 = __shfl_xor_sync(0xffffffffu, u1, o), ou2 = __shfl_xor_sync(0xffffffffu, u2, o);
                const int oj1 = __shfl_xor_sync(0xffffffffu, j1, o), oj2 = __shfl_xor_sync(0xffffffffu, j2, o);
                if (pair_less(ou1, oj1, u1, j1)) {
                    // other's best wins; second = min(my best, other's second)
                    if (pair_less(u1, j1, ou2, oj2)) { u2 = u1; j2 = j1; } else { u2 = ou2; j2 = oj2; }
                    u1 = ou1; j1 = oj1;
                } else {
                    if (pair_less(ou1, oj1, u2, j2)) { u2 = ou1; j2 = oj1; }
                }
            }
            if (j2 < 0) u2 = kBig;
            int i0 = s.y[j1];
            const double vj1 = s.v[j1];
            const double v1_new = vj1 - (u2 - u1);
            const bool lowers = v1_new < vj1;
            __syncwarp();
            if (visits < (unsigned long long)cur * (unsigned long long)n) {
                if (lowers) { if (lane == 0) s.v[j1] = v1_new; }
                else if (i0 >= 0 && j2 >= 0) { j1 = j2; i0 = s.y[j2]; }
                if (i0 >= 0) {
                    if (lowers) { --cur; if (lane == 0) s.free_rows[cur] = i0; }
                    else { if (lane == 0) s.free_rows[nxt] = i0; ++nxt; }
                }
            } else if (i0 >= 0) {
                if (lane == 0) s.free_rows[nxt] = i0;
                ++nxt;
            }
            if (lane == 0) { s.x[fr] = j1; s.y[j1] = fr; }
            __syncwarp();
        }
        nfree = nxt;
        if (!ok) break;
    }

    // ---- augmentation ----------------------------------------------------------
    P.mark(3, 4); P.mark(4, nfree);
    const int words = (n + 31) >> 5;
    for (int f = 0; f < nfree && ok; ++f) {
        const int start = s.free_rows[f];
        for (int j = lane; j < n; j += 32) {
            s.cols[j] = j;
            s.pred[j] = start;
            s.d[j] = P.at(start, j) - s.v[j];
        }
        __syncwarp();
        int lo = 0, hi = 0, n_ready = 0, final_j = -1;
        int guard = 0;
        while (final_j == -1) {
            if (++guard > 2 * n + 8) { ok = false; break; }
            if (lo == hi) {
                n_ready = lo;
                // -- collect the columns of minimal d into SCAN, replaying the
                //    sequential sweep's swaps (triggers found in parallel)
                double run = s.d[s.cols[lo]];
                for (int w = lane; w < words; w += 32) { s.mask_a[w] = 0u; s.mask_b[w] = 0u; }
                __syncwarp();
                for (int k0 = lo + 1; k0 < n; k0 += 32) {
                    const int k = k0 + lane;
                    double dv = 1.0e300;
                    if (k < n) dv = s.d[s.cols[k]];
                    double inc = dv;                       // inclusive prefix min
                    for (int o = 1; o < 32; o <<= 1) {
                        const double t = __shfl_up_sync(0xffffffffu, inc, o);
                        if (lane >= o) inc = fmin(inc, t);
                    }
                    double exc = __shfl_up_sync(0xffffffffu, inc, 1);
                    exc = lane == 0 ? run : fmin(exc, run);
                    const unsigned strict = __ballot_sync(0xffffffffu, k < n && dv < exc);
                    const unsigned equal = __ballot_sync(0xffffffffu, k < n && dv == exc);
                    if (lane == 0) {
                        // bits are relative to k0; store at bit offset k0 (may straddle words)
                        const int w = k0 >> 5, sh = k0 & 31;
                        s.mask_a[w] |= strict << sh;
                        s.mask_b[w] |= equal << sh;
                        if (sh && w + 1 < words) {
                            s.mask_a[w + 1] |= strict >> (32 - sh);
                            s.mask_b[w + 1] |= equal >> (32 - sh);
                        }
                    }
                    run = fmin(run, __shfl_sync(0xffffffffu, inc, 31));
                    __syncwarp();
                }
                if (lane == 0) {
                    int h2 = lo + 1;
                    for (int w = 0; w < words; ++w) {
                        unsigned any = s.mask_a[w] | s.mask_b[w];
                        while (any) {
                            const int b = __ffs(any) - 1;
                            any &= any - 1;
                            const int k = (w << 5) + b;
                            if ((s.mask_a[w] >> b) & 1u) h2 = lo;
                            const int j = s.cols[k];
                            s.cols[k] = s.cols[h2];
                            s.cols[h2++] = j;
                        }
                    }
                    s.mask_a[0] = (unsigned)h2;
                }
                __syncwarp();
                hi = (int)s.mask_a[0];
                __syncwarp();
                // last unassigned column in SCAN order wins
                int cand = -1;
                for (int k = lo + lane; k < hi; k += 32)
                    if (s.y[s.cols[k]] < 0) cand = k;
                cand = warp_max_i(cand);
                if (cand >= 0) final_j = s.cols[cand];
            }
            if (final_j == -1) {
                // -- relax TODO columns from every SCAN column
                while (lo != hi) {
                    if (lo < 0 || hi > n || lo > hi) { ok = false; final_j = -2; break; }
                    const int js = s.cols[lo++];
                    const int i = s.y[js];
                    const double mind = s.d[js];
                    const double h = P.at(i, js) - s.v[js] - mind;
                    for (int w = lane; w < words; w += 32) s.mask_a[w] = 0u;
                    __syncwarp();
                    int first_free = 0x7fffffff;
                    for (int k0 = hi; k0 < n; k0 += 32) {
                        const int k = k0 + lane;
                        bool hit = false, freehit = false;
                        if (k < n) {
                            const int j = s.cols[k];
                            const double r = P.at(i, j) - s.v[j] - h;
                            if (r < s.d[j]) {
                                s.d[j] = r;
                                s.pred[j] = i;
                                if (r == mind) { hit = true; freehit = s.y[j] < 0; }
                            }
                        }
                        const unsigned hb = __ballot_sync(0xffffffffu, hit);
                        const unsigned fb = __ballot_sync(0xffffffffu, freehit);
                        if (fb && first_free == 0x7fffffff) first_free = k0 + __ffs(fb) - 1;
                        if (lane == 0 && hb) {
                            const int rel = k0 - hi;          // bit index relative to hi
                            const int w = rel >> 5, sh = rel & 31;
                            s.mask_a[w] |= hb << sh;
                            if (sh && w + 1 < words) s.mask_a[w + 1] |= hb >> (32 - sh);
                        }
                    }
                    __syncwarp();
                    if (first_free != 0x7fffffff) { final_j = s.cols[first_free]; break; }
                    if (lane == 0) {
                        int h2 = hi;
                        const int span = n - hi;
                        for (int w = 0; w * 32 < span; ++w) {
                            unsigned any = s.mask_a[w];
                            while (any) {
                                const int b = __ffs(any) - 1;
                                any &= any - 1;
                                const int k = hi + (w << 5) + b;
                                const int j = s.cols[k];
                                s.cols[k] = s.cols[h2];
                                s.cols[h2++] = j;
                            }
                        }
                        s.mask_b[0] = (unsigned)h2;
                    }
                    __syncwarp();
                    hi = (int)s.mask_b[0];
                    __syncwarp();
                }
                if (final_j != -1) lo = n_ready;      // caller's lo is unchanged by an early return
            }
        }
        if (!ok || final_j < 0) { ok = false; break; }
        {
            const double mind = s.d[s.cols[lo]];
            for (int k = lane; k < n_ready; k += 32) {
                const int j = s.cols[k];
                s.v[j] += s.d[j] - mind;
            }
        }
        __syncwarp();
        if (lane == 0) {
            int i = -1, j = final_j, steps = 0;
            while (i != start && steps++ <= n) {
                i = s.pred[j];
                s.y[j] = i;
                const int t = j; j = s.x[i]; s.x[i] = t;
            }
        }
        __syncwarp();
    }
    P.mark(3, ok ? 5 : -5);
    return ok;
}

// ---------------------------------------------------------------------------------
// Fast path.  The extended problem is equivalent to a maximum-weight bipartite
// matching with "stay unmatched" options: minimise  sum_M c_ij + L/2 * (unmatched
// rows + unmatched cols).  It is solved here as a rectangular assignment of the R
// real rows to C real columns (cost c_ij - L/2) or to a private dummy column (cost
// L/2) with the row-by-row shortest-augmenting-path (Hungarian) method, one warp,
// float64 potentials.  Any exact solver returns the reference's assignment when the
// optimum is unique, so the result is accepted only with a uniqueness certificate
// from the optimal duals (u_i + v_j <= a_ij, v_j <= 0, v_j = 0 on free columns).
// Every other optimal assignment uses tight edges only and keeps every column
// with v_j < 0 occupied, so it differs from this one by re-routings along arcs
// sigma(i) -> j (one per tight non-matching edge (i,j), dummy edges included):
//   (1) a path from an occupied column that may be vacated (v_j >= -eps) to a
//       free column, or
//   (2) a directed cycle.
// If neither exists (tolerance eps on tightness) the real part is unique and equals
// the reference's.  Without the certificate -
// exact cost ties such as duplicate boxes - the caller runs solve_warp, the literal
// emulation of the reference's solver, so tie-breaks still match.
// Rows whose best cost is >= L can never be matched: they are priced u_i = L/2
// and skipped.
// Scratch use: s.v = column potentials v_j (C real + R dummy columns), s.d =
// minv, (s.cols,s.pred) reinterpreted as the R row potentials, s.x = p (column ->
// row), s.y = way, s.unique = visited flags, s.free_rows = active row list.
// On success s.y[j] (j < C) holds the row of column j or -1, s.x[i] the column of
// row i or -1 - the same contract as solve_warp after truncation.
constexpr double kFastEps = 1.0e-9;

// One warp fills the admissible-edge statistics of P (see Problem).
__device__ inline void count_admissible(const Problem &P) {
    const int lane = threadIdx.x & 31;
    const double L = 2.0 * P.half;
    for (int j = lane; j < P.C; j += 32) P.col_cnt[j] = 0;
    if (lane == 0) { P.adm_misc[0] = 0; P.adm_misc[1] = 0; }
    __syncwarp();
    for (int i = lane; i < P.R; i += 32) {
        int cnt = 0, adm = -1;
        bool nan = false;
        for (int j = 0; j < P.C; ++j) {
            const double c = (double)P.cost[(long long)i * P.C + j];
            nan |= !(c == c);
            if (fabs(c - L) <= 2.0e-9) P.adm_misc[1] = 1;
            if (c < L) { ++cnt; adm = j; atomicAdd(&P.col_cnt[j], 1); }
        }
        P.row_cnt[i] = cnt;
        P.row_adm[i] = adm;
        if (nan) P.adm_misc[0] = 1;
    }
    __syncwarp();
}

// argmin over the warp of (value, index) pairs, lowest index among equal values,
// index < 0 = "none"; three REDUX instead of a 5-level shuffle tree on 12 bytes.
__device__ __forceinline__ void warp_argmin(double &val, int &idx) {
    // order-preserving map of the double onto an unsigned 64-bit key; "none" sorts last
    unsigned long long k = (unsigned long long)__double_as_longlong(val);
    k = (k >> 63) ? ~k : (k | 0x8000000000000000ull);
    if (idx < 0) k = ~0ull;
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    const bool mine = hi == mhi && lo == mlo && idx >= 0;
    const unsigned midx = __reduce_min_sync(0xffffffffu, mine ? (unsigned)idx : 0xffffffffu);
    const unsigned who = __ballot_sync(0xffffffffu, mine && (unsigned)idx == midx);
    if (who == 0u) { idx = -1; return; }
    const int src = __ffs(who) - 1;
    val = __shfl_sync(0xffffffffu, val, src);
    idx = (int)midx;
}

// Row-by-row Hungarian with the per-column state (minv, v, visited, way) held in
// registers: lane l owns the real columns l, l+32, ... (KC per lane, C <= 32*KC).
// The private dummy columns are not materialised: a row that joins the
// alternating tree contributes the candidate "leave through my own dummy" with
// slack L/2 - u_row, and all such slacks shrink together, so one warp-uniform
// scalar (dmin, drow, dcol) tracks the best of them.  Real columns win ties
// against dummies (lower index in the materialised formulation).
// u (row potentials), p (column -> row) and way live in shared memory: they are
// touched once per tree step, not per column.  Returns false on a guard.
template <int KC>
__device__ bool hungarian_regs(const Problem &P, volatile double *u, volatile double *v_out,
                               volatile int *p, volatile int *way_sm, volatile int *active, int n_active) {
    const int lane = threadIdx.x & 31;
    const int R = P.R, C = P.C;
    const double half = P.half, INF = 1.0e300;
    const int VIRT = 1 << 30;                      // the virtual start column
    double vv[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) vv[k] = 0.0;
    for (int t = 0; t < n_active; ++t) {
        const int row = active[t];
        if (row < 0) continue;                     // matched by the greedy start
        double minv[KC];
        int way[KC];
        unsigned used = 0u;
#pragma unroll
        for (int k = 0; k < KC; ++k) { minv[k] = INF; way[k] = VIRT; }
        int j0 = VIRT, i0 = row;
        double dmin = half - u[row];               // the new row's own dummy
        int drow = row, dcol = VIRT;
        bool to_dummy = false;
        int guard = 0;
        while (true) {
            if (++guard > R + C + 2) return false;
            const double ui0 = u[i0];
            double best = INF;
            int bj = -1;
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int j = lane + 32 * k;
                if (j < C && !((used >> k) & 1u)) {
                    const double cur = ((double)P.cost[(long long)i0 * C + j] - half) - ui0 - vv[k];
                    if (cur < minv[k]) { minv[k] = cur; way[k] = j0; }
                    if (minv[k] < best) { best = minv[k]; bj = j; }
                }
            }
            warp_argmin(best, bj);
            to_dummy = bj < 0 || dmin < best;
            const double delta = to_dummy ? dmin : best;
            if (!(delta < INF)) return false;
            // dual update
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int j = lane + 32 * k;
                if (j < C) {
                    if ((used >> k) & 1u) { u[p[j]] += delta; vv[k] -= delta; }
                    else minv[k] -= delta;
                }
            }
            if (lane == 0) u[row] += delta;
            dmin -= delta;
            __syncwarp();
            if (to_dummy) break;
            j0 = bj;
            if ((j0 & 31) == lane) used |= 1u << (j0 >> 5);
            i0 = p[j0];
            if (i0 < 0) break;                     // free real column: augment
            const double cand = half - u[i0];
            if (cand < dmin) { dmin = cand; drow = i0; dcol = j0; }
        }
        // publish `way` and flip the alternating path (lane 0)
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int j = lane + 32 * k;
            if (j < C) way_sm[j] = way[k];
        }
        __syncwarp();
        if (lane == 0) {
            int jc = to_dummy ? dcol : j0;         // column that receives a new row (VIRT: none)
            if (to_dummy && dcol != VIRT) {
                // drow leaves its column for its dummy; that column is re-filled along the path
            }
            int steps = 0;
            while (jc != VIRT && steps++ <= R + C) {
                const int jp = way_sm[jc];
                p[jc] = jp == VIRT ? row : p[jp];
                jc = jp;
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int j = lane + 32 * k;
        if (j < C) v_out[j] = vv[k];
    }
    for (int i = lane; i < R; i += 32) v_out[C + i] = 0.0;     // dummy columns keep v = 0
    __syncwarp();
    return true;
}

// Uniqueness analysis of an optimal assignment from its tight non-matching edges, given as arcs
// sigma(i) -> j over the column nodes [0, m) (real columns < C, then the private dummy column C + i of
// row i): p[j] = row of real column j or -1, sigma[i] = column node of row i, vnode(j) = potential of
// column node j (0 for dummies), flags = 4 m bytes of scratch.  Returns 0 when no equal-cost
// alternative exists, else the refusal code (6 / 7): (1) a free column is reachable from a column that
// can be vacated at no cost, or (2) the arcs hold a directed cycle.  One warp.
template <typename VNode>
__device__ inline int arcs_unique(int C, int m, int n_arcs, const volatile int *arc_from, const volatile int *arc_to,
                                  const volatile int *p, const volatile int *sigma, VNode vnode,
                                  volatile unsigned char *flags, long long *t_mid) {
    const int lane = threadIdx.x & 31;
    if (n_arcs > 0 && n_arcs <= 32) {
        // few arcs (the usual case): one arc per lane, fixpoints by shuffles, no shared memory.
        // `from` is always an assigned column (sigma of a row); only columns that occur in arcs matter.
        __syncwarp();
        const bool have = lane < n_arcs;
        const int from = have ? arc_from[lane] : -1, to = have ? arc_to[lane] : -2;
        bool to_assigned = false, from_vac = false;
        if (have) {
            to_assigned = to < C ? p[to] >= 0 : sigma[to - C] == to;
            from_vac = vnode(from) >= -kFastEps;                 // may be vacated at no cost
        }
        // (1) a free column reachable from a vacatable one
        bool reach = have && !to_assigned;
        for (int round = 0; round <= n_arcs; ++round) {
            bool changed = false;
            for (int k = 0; k < n_arcs; ++k) {
                const int f = __shfl_sync(0xffffffffu, from, k);
                const bool r = __shfl_sync(0xffffffffu, (int)reach, k) != 0;
                if (have && !reach && r && f == to) { reach = true; changed = true; }
            }
            if (!__any_sync(0xffffffffu, changed)) break;
        }
        if (t_mid) *t_mid = clock64();
        bool unsure = reach && from_vac;
        // (2) a directed cycle: an arc stays alive while its head has an alive outgoing arc
        bool alive = have && to_assigned;
        for (int round = 0; round <= n_arcs; ++round) {
            bool has_next = false;
            for (int k = 0; k < n_arcs; ++k) {
                const int f = __shfl_sync(0xffffffffu, from, k);
                const bool al = __shfl_sync(0xffffffffu, (int)alive, k) != 0;
                has_next |= al && f == to;
            }
            const bool drop = alive && !has_next;
            if (drop) alive = false;
            if (!__any_sync(0xffffffffu, drop)) break;
        }
        unsure |= alive;
        if (__any_sync(0xffffffffu, unsure)) return 6;
    } else if (n_arcs > 0) {
        volatile unsigned char *assigned = flags;
        volatile unsigned char *reach = assigned + m, *alive = assigned + 2 * m, *has_out = assigned + 3 * m;
        bool vac_l[16];
        int cnt = 0;
        for (int j = lane; j < m; j += 32, ++cnt) {
            const bool asg = j < C ? p[j] >= 0 : sigma[j - C] == j;
            if (cnt < 16) vac_l[cnt] = asg && vnode(j) >= -kFastEps;     // may be vacated at no cost
            assigned[j] = asg; reach[j] = !asg; alive[j] = asg;
        }
        __syncwarp();
        // (1) columns from which a free column is reachable along arcs
        for (int round = 0; round <= m; ++round) {
            bool changed = false;
            for (int k = lane; k < n_arcs; k += 32)
                if (reach[arc_to[k]] && !reach[arc_from[k]]) { reach[arc_from[k]] = 1; changed = true; }
            __syncwarp();
            if (!__any_sync(0xffffffffu, changed)) break;
        }
        if (t_mid) *t_mid = clock64();
        bool unsure = false;
        cnt = 0;
        for (int j = lane; j < m; j += 32, ++cnt)
            if (cnt < 16 && vac_l[cnt] && reach[j]) unsure = true;   // a free re-routing exists
        // (2) directed cycle among assigned columns (peel columns without live successors)
        for (int round = 0; round <= m; ++round) {
            for (int j = lane; j < m; j += 32) has_out[j] = 0;
            __syncwarp();
            for (int k = lane; k < n_arcs; k += 32)
                if (alive[arc_from[k]] && alive[arc_to[k]]) has_out[arc_from[k]] = 1;
            __syncwarp();
            bool changed = false;
            for (int j = lane; j < m; j += 32)
                if (alive[j] && !has_out[j]) { alive[j] = 0; changed = true; }
            __syncwarp();
            if (!__any_sync(0xffffffffu, changed)) break;
        }
        for (int j = lane; j < m; j += 32) unsure |= alive[j] != 0;    // survivors lie on a cycle
        if (__any_sync(0xffffffffu, unsure)) return 7;
    }
    return 0;
}

static __device__ __noinline__ bool fast_assign_warp(const Problem &P, Scratch &s) {
    const int lane = threadIdx.x & 31;
    const int R = P.R, C = P.C, m = R + C;
    const double half = P.half, L = 2.0 * P.half;
    volatile double *u = reinterpret_cast<volatile double *>(s.cols);   // [R] (cols,pred) = 2n ints
    volatile double *v = s.v, *minv = s.d;
    volatile int *p = s.x, *way = s.y, *active = s.free_rows;
    volatile unsigned char *used = s.unique;
    const double INF = 1.0e300;
    const long long t_f0 = clock64();
#define ST_REFUSE(code)                                                                     \
    do {                                                                                    \
        if (P.why && lane == 0) { P.why[0] = (code); atomicAdd(&P.why[1], 1); }             \
        return false;                                                                       \
    } while (0)
    // ---- row pruning, NaN screen, isolated pairs ----------------------------------------
    // Rows without an admissible entry (cost < L) can never be matched: u_i = L/2.
    // A row whose only admissible column has no other admissible row is matched
    // to it right away with u_i = a_ij, v_j = 0 - the state the row-by-row method
    // reaches when it inserts that row first (every other entry of the row and of
    // the column costs at least the dummy).  In tracking most pairs are of this
    // kind, so the shortest-path search runs only for the crowded remainder.
    if (P.adm_misc[0] || P.adm_misc[1]) ST_REFUSE(1);      // NaN, or a cost within 2e-9 of L (tie with the dummies)
    for (int j = lane; j < m; j += 32) { v[j] = 0.0; p[j] = -1; way[j] = 0x7fffffff; }
    __syncwarp();
    // Greedy start for the remaining rows: u_i = the row's smallest reduced entry (duals stay feasible:
    // v = 0, every entry of the row >= u_i, the dummy costs half > u_i) and the row takes that column
    // unless a lower row wants it too (way[] collects the lowest claimant).  Only the rows that lose
    // a claim go through the shortest-path search - in a crowd nearly every track still has "its"
    // detection.  `active` lists every non-isolated row with admissible entries (the certificate scans
    // them all); rows matched by the greedy start are stored as ~row and skipped by the search.
    int n_active = 0;
    for (int i0 = 0; i0 < R; i0 += 32) {
        const int i = i0 + lane;
        bool act = false;
        int best_j = -1;
        if (i < R) {
            const int cnt = P.row_cnt[i];
            double ui = half;
            if (cnt == 1 && P.col_cnt[P.row_adm[i]] == 1) {
                const int j = P.row_adm[i];
                p[j] = i;
                ui = (double)P.cost[(long long)i * C + j] - half;
            } else if (cnt > 0) {
                act = true;
                float bc = 3.0e38f;
                const float *crow = P.cost + (long long)i * C;
                for (int j = 0; j < C; ++j) { const float c = crow[j]; if (c < bc) { bc = c; best_j = j; } }
                ui = P.no_greedy ? 0.0 : (double)bc - half;
                atomicMin(const_cast<int *>(way) + best_j, i);
            }
            u[i] = ui;
        }
        const unsigned mb = __ballot_sync(0xffffffffu, act);
        if (act) active[n_active + __popc(mb & ((1u << lane) - 1u))] = i | (best_j << 16);
        n_active += __popc(mb);
    }
    __syncwarp();
    int n_search = 0;
    for (int t = lane; t < n_active; t += 32) {
        const int packed = active[t], i = packed & 0xffff, j = packed >> 16;
        const bool won = way[j] == i && !P.no_greedy;
        if (won) p[j] = i;
        active[t] = won ? ~i : i;
        n_search += won ? 0 : 1;
    }
    __syncwarp();
    const int n_search_total = __reduce_add_sync(0xffffffffu, n_search);
    // a(i, j): real column -> c_ij - half ; own dummy column C+i -> half ; else INF
    auto a = [&](int i, int j) -> double {
        if (j < C) return (double)P.cost[(long long)i * C + j] - half;
        return (j - C == i) ? half : INF;
    };
    // ---- Hungarian, rows inserted one by one ---------------------------------------------
    const long long t_h0 = clock64();
    bool solved = false;
    if (C <= 256) {
        bool ok;
        if (C <= 32) ok = hungarian_regs<1>(P, u, v, p, way, active, n_active);
        else if (C <= 64) ok = hungarian_regs<2>(P, u, v, p, way, active, n_active);
        else if (C <= 128) ok = hungarian_regs<4>(P, u, v, p, way, active, n_active);
        else ok = hungarian_regs<8>(P, u, v, p, way, active, n_active);
        if (!ok) ST_REFUSE(2);
        solved = true;
    }
    if (!solved) {
    for (int t = 0; t < n_active; ++t) {
        const int row = active[t];
        if (row < 0) continue;                     // matched by the greedy start
        for (int j = lane; j < m; j += 32) { minv[j] = INF; used[j] = 0; }
        __syncwarp();
        int j0 = m;                       // virtual start column holding `row`
        int guard = 0;
        while (true) {
            if (++guard > m + 2) ST_REFUSE(3);
            if (j0 < m && lane == 0) used[j0] = 1;
            __syncwarp();
            const int i0 = j0 == m ? row : p[j0];
            const double ui0 = u[i0];
            double delta = INF;
            int j1 = -1;
            for (int j = lane; j < m; j += 32) {
                if (used[j]) continue;
                // dummy columns of other rows are unreachable
                if (j >= C && j - C != i0) {
                    const double mv = minv[j];
                    if (mv < delta) { delta = mv; j1 = j; }
                    continue;
                }
                const double cur = a(i0, j) - ui0 - v[j];
                double mv = minv[j];
                if (cur < mv) { mv = cur; minv[j] = cur; way[j] = j0; }
                if (mv < delta) { delta = mv; j1 = j; }
            }
            for (int o = 16; o; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, delta, o);
                const int oj = __shfl_xor_sync(0xffffffffu, j1, o);
                if (pair_less(od, oj, delta, j1)) { delta = od; j1 = oj; }
            }
            if (j1 < 0 || !(delta < INF)) ST_REFUSE(4);
            __syncwarp();
            // dual update: visited columns (and the virtual one) move with their rows
            for (int j = lane; j < m; j += 32) {
                if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
                else minv[j] -= delta;
            }
            if (lane == 0) u[row] += delta;          // the virtual column's row
            __syncwarp();
            j0 = j1;
            if (p[j0] < 0) break;
        }
        if (lane == 0) {
            int steps = 0;
            while (j0 != m && steps++ <= m) {       // flip the alternating path
                const int jp = way[j0];
                p[j0] = jp == m ? row : p[jp];
                j0 = jp;
            }
        }
        __syncwarp();
    }
    }
    const long long t_h1 = clock64();
    // ---- uniqueness certificate ---------------------------------------------------------
    // sigma[i] = column of row i (own dummy C+i when unmatched); arcs sigma(i) -> j for
    // every tight non-matching edge (i, j), dummy edges included.
    volatile int *sigma = way;                         // `way` is dead after the solve
    for (int i = lane; i < R; i += 32) sigma[i] = C + i;
    __syncwarp();
    for (int j = lane; j < C; j += 32) { const int r = p[j]; if (r >= 0) sigma[r] = j; }
    __syncwarp();
    volatile int *arc_from = reinterpret_cast<volatile int *>(s.d);     // s.d = 2n ints
    volatile int *arc_to = arc_from + m;
    int n_arcs = 0;
    // Only the rows the search handled can have a tight non-matching real edge: a row without
    // admissible entries has u = L/2 and an isolated pair keeps u = a_ij < L/2, so every other
    // entry of such a row has slack >= (cost - L) > eps (no cost lies within 2e-9 of L:
    // adm_misc[1], checked above).  The real edges of the active rows, then one dummy edge per row:
    // one active row per lane, all lanes stepping through the columns together.  The test runs
    // in float32 first (dependent float64 operations cost ~50 cycles each on this part, and
    // nearly every edge is far from tight): the float32 slack differs from the float64 one by
    // < 1e-6, so only edges within 1e-4 are re-evaluated exactly.  The float32 copy of the
    // column potentials sits behind the R row potentials in the (cols, pred) region (n doubles).
    // (plain pointer: the __syncwarp() below orders the writes before the reads, and a volatile
    // generic access to shared memory costs far more than an ordinary one)
    float *vf = reinterpret_cast<float *>(const_cast<double *>(u + R));
    for (int j = lane; j < C; j += 32) vf[j] = (float)v[j];
    __syncwarp();
    // Two traversals of the same edge set: with few active rows, row by row with the lanes over
    // the columns (cost ~ rows x ceil(C / 32) steps); with many, one row per lane and all lanes
    // stepping through the columns together (cost ~ C steps per 32 rows).
    if (n_active * ((C + 31) / 32) * 2 < C) {
        for (int t = 0; t < n_active; ++t) {
            const int i = active[t] < 0 ? ~active[t] : active[t];
            const double ui = u[i];
            const float hu = (float)(half + ui);
            const int si = sigma[i];
            const float *crow = P.cost + (long long)i * C;
            for (int j0 = 0; j0 < C; j0 += 32) {
                const int j = j0 + lane;
                bool tight = j < C && j != si && (crow[j] - hu) - vf[j] <= 1.0e-4f;
                if (tight) tight = ((double)crow[j] - half) - ui - v[j] <= kFastEps;
                const unsigned tb = __ballot_sync(0xffffffffu, tight);
                if (tb == 0u) continue;
                const int pos = n_arcs + __popc(tb & ((1u << lane) - 1u));
                if (tight && pos < m) { arc_from[pos] = si; arc_to[pos] = j; }
                n_arcs += __popc(tb);
            }
        }
    } else
    for (int t0 = 0; t0 < n_active; t0 += 32) {
        const int t = t0 + lane;
        const bool hv = t < n_active;
        const int i = hv ? (active[t] < 0 ? ~active[t] : active[t]) : 0;
        const double ui = hv ? u[i] : 0.0;
        const float hu = (float)(half + ui);
        const int si = hv ? sigma[i] : -1;
        const float *crow = P.cost + (long long)i * C;
        // four columns per step: one vote decides whether any of them needs a closer look (the
        // loop-carried vote + branch, not the loads, set the pace of a one-column step)
        for (int j4 = 0; j4 < C; j4 += 4) {
            bool near[4];
            bool any_near = false;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j4 + q;
                near[q] = hv && j < C && j != si && (crow[j < C ? j : 0] - hu) - vf[j < C ? j : 0] <= 1.0e-4f;
                any_near |= near[q];
            }
            if (!__any_sync(0xffffffffu, any_near)) continue;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = j4 + q;
                bool tight = near[q];
                if (tight) tight = ((double)crow[j] - half) - ui - v[j] <= kFastEps;
                const unsigned tb = __ballot_sync(0xffffffffu, tight);
                if (tb == 0u) continue;
                const int pos = n_arcs + __popc(tb & ((1u << lane) - 1u));
                if (tight && pos < m) { arc_from[pos] = si; arc_to[pos] = j; }
                n_arcs += __popc(tb);
            }
        }
    }
    for (int i0 = 0; i0 < R; i0 += 32) {               // the dummy edge of every matched row
        const int i = i0 + lane;
        bool tight = false;
        int from = 0;
        if (i < R) {
            from = sigma[i];
            if (from != C + i) tight = half - u[i] - v[C + i] <= kFastEps;
        }
        const unsigned tb = __ballot_sync(0xffffffffu, tight);
        const int pos = n_arcs + __popc(tb & ((1u << lane) - 1u));
        if (tight && pos < m) { arc_from[pos] = from; arc_to[pos] = C + i; }
        n_arcs += __popc(tb);
    }
    if (n_arcs > m) ST_REFUSE(5);                       // more tight edges than the list holds
    __syncwarp();
    const long long t_c1 = clock64();
    long long t_c2 = t_c1;
    {
        const int verdict = arcs_unique(C, m, n_arcs, arc_from, arc_to, p, sigma, [&](int j) { return (double)v[j]; },
                                        reinterpret_cast<volatile unsigned char *>(s.cols), &t_c2);
        if (verdict) ST_REFUSE(verdict);
    }
    __syncwarp();
    if (P.why && lane == 0) {
        P.why[2] = (int)(t_h0 - t_f0); P.why[3] = (int)(t_h1 - t_h0); P.why[4] = (int)(clock64() - t_h1);
        P.why[5] = n_search_total; P.why[6] = n_active; P.why[7] = n_arcs;
    }
    // diagnostics (debug words 14 / 15): cycles of the search and of the certificate, last call
    if (P.dbg && lane == 0 && P.dbg[0] == 1) {          // stage 1 of the track step only
        P.dbg[14] = (int)(t_h1 - t_h0); P.dbg[15] = (int)(clock64() - t_h1);
        P.dbg[16] = (int)(t_c1 - t_h1); P.dbg[17] = (int)(t_c2 - t_c1); P.dbg[18] = (int)(clock64() - t_c2);
        P.dbg[19] = n_arcs;
        P.dbg[3] = n_active; P.dbg[4] = R * 1000 + C;
    }
    // ---- publish in solve_warp's layout: y = col -> row, x = row -> col -------------------
    // (p aliases s.x and way aliases s.y, so go through registers)
    int colrow[16];                                   // m <= 512 per lane stride 32
    int cnt = 0;
    for (int j = lane; j < C && cnt < 16; j += 32) colrow[cnt++] = p[j];
    __syncwarp();
    for (int i = lane; i < m; i += 32) s.x[i] = -1;
    __syncwarp();
    cnt = 0;
    for (int j = lane; j < C; j += 32) {
        const int r = cnt < 16 ? colrow[cnt] : -1;
        ++cnt;
        s.y[j] = r;
        if (r >= 0) s.x[r] = j;
    }
    __syncwarp();
    return true;
}
#undef ST_REFUSE

// ---------------------------------------------------------------------------------
// Sparse certified solver (crowded or wide problems: more than 64 rows or columns).
// Only ADMISSIBLE entries (cost < L) can be part of an optimum, and during the shortest-path search a
// non-admissible real column never beats the private dummy of the same row (its slack is
// c - L/2 - u_i - v_j >= L/2 - u_i because v_j <= 0, with equality excluded by the 2e-9 screen), so the
// whole method runs on the admissible edges alone: rows / columns without admissible entries are
// dropped, the rest is renumbered compactly (nR rows, nC columns) and the edges are held as a CSR list
// (a hundred tracks in a crowd have ~4 admissible detections each: ~600 edges instead of a
// 170 x 240 matrix in global memory).
//   start:   u_i = the row's smallest reduced entry; the row takes that column unless a lower row wants
//            it too (isolated pairs are the trivial case) - feasible duals, v = 0;
//   search:  the rows that lost a claim, one after the other, Hungarian shortest-path search with the
//            same rules as hungarian_regs (row's own dummy tracked as one scalar, real columns win
//            ties against dummies, lowest column index among equal slacks); the per-column state
//            lives in shared memory and only the columns touched by the search (the frontier, a
//            handful) are scanned and updated per step;
//   certificate: tight non-matching admissible edges + the dummy edge of every matched row -> arcs ->
//            arcs_unique (non-admissible edges are never tight: slack >= c - L > 2e-9).
// One warp.  Results in COMPACT numbering: p[j] = row of column j or -1.  Returns 0, or a refusal code.
struct Sparse {
    int nR, nC, E;
    double half;
    const unsigned short *off;       // [nR + 1]
    const unsigned short *erow;      // [E] row of every edge
    const unsigned short *ecol;      // [E] column of every edge
    const float *ecost;              // [E]
    // state (shared memory)
    volatile double *u;              // [nR]
    volatile double *v, *minv;       // [nC]
    volatile int *p, *way;           // [nC]
    volatile unsigned short *fcol;   // [nC] frontier of the current search
    volatile unsigned char *used;    // [nC]
    volatile int *freel;             // [nR] rows to search
    volatile int *sigma;             // [nR]
    volatile int *arc_from, *arc_to; // [nR + nC]
    volatile unsigned char *flags;   // [4 (nR + nC)]
    int *why;                        // optional diagnostics as Problem::why
};

__host__ __device__ inline size_t sparse_state_bytes(int nR, int nC) {
    size_t b = sizeof(double) * (size_t)(nR + 2 * nC) + sizeof(int) * (size_t)(2 * nC + 2 * nR + 2 * (nR + nC));
    b += sizeof(unsigned short) * (size_t)nC + (size_t)nC + 4 * (size_t)(nR + nC);
    return (b + 15) & ~(size_t)15;
}

// carves the state arrays of q (nR, nC set) out of mem (sparse_state_bytes(nR, nC) bytes, 8-byte aligned)
__device__ inline void sparse_carve_state(Sparse &q, unsigned char *mem) {
    q.u = reinterpret_cast<volatile double *>(mem); mem += sizeof(double) * q.nR;
    q.v = reinterpret_cast<volatile double *>(mem); mem += sizeof(double) * q.nC;
    q.minv = reinterpret_cast<volatile double *>(mem); mem += sizeof(double) * q.nC;
    q.p = reinterpret_cast<volatile int *>(mem); mem += sizeof(int) * q.nC;
    q.way = reinterpret_cast<volatile int *>(mem); mem += sizeof(int) * q.nC;
    q.freel = reinterpret_cast<volatile int *>(mem); mem += sizeof(int) * q.nR;
    q.sigma = reinterpret_cast<volatile int *>(mem); mem += sizeof(int) * q.nR;
    q.arc_from = reinterpret_cast<volatile int *>(mem); mem += sizeof(int) * (q.nR + q.nC);
    q.arc_to = reinterpret_cast<volatile int *>(mem); mem += sizeof(int) * (q.nR + q.nC);
    q.fcol = reinterpret_cast<volatile unsigned short *>(mem); mem += sizeof(unsigned short) * q.nC;
    q.used = mem; mem += q.nC;
    q.flags = mem;
}

static __device__ __noinline__ int sparse_assign_warp(const Sparse &q) {
    const int lane = threadIdx.x & 31;
    const int nR = q.nR, nC = q.nC, m = nR + nC;
    const double half = q.half, INF = 1.0e300;
    const int VIRT = 1 << 30;
    const long long t0 = clock64();
    for (int j = lane; j < nC; j += 32) { q.v[j] = 0.0; q.minv[j] = INF; q.p[j] = -1; q.way[j] = 0x7fffffff; q.used[j] = 0; }
    __syncwarp();
    // ---- greedy start ----
    for (int i = lane; i < nR; i += 32) {
        float bc = 3.0e38f;
        int bj = 0x7fffffff;
        for (int e = q.off[i]; e < q.off[i + 1]; ++e) {
            const float c = q.ecost[e];
            const int j = q.ecol[e];
            if (c < bc || (c == bc && j < bj)) { bc = c; bj = j; }
        }
        q.u[i] = (double)bc - half;
        q.sigma[i] = bj;                                   // the wanted column, until the certificate
        atomicMin(const_cast<int *>(q.way) + bj, i);
    }
    __syncwarp();
    int n_free = 0;
    for (int i0 = 0; i0 < nR; i0 += 32) {
        const int i = i0 + lane;
        bool lost = false;
        if (i < nR) {
            const int j = q.sigma[i];
            if (q.way[j] == i) q.p[j] = i; else lost = true;
        }
        const unsigned mb = __ballot_sync(0xffffffffu, lost);
        if (lost) q.freel[n_free + __popc(mb & ((1u << lane) - 1u))] = i;
        n_free += __popc(mb);
    }
    __syncwarp();
    const long long t1 = clock64();
    // ---- shortest-path searches for the rows that lost their claim ----
    for (int t = 0; t < n_free; ++t) {
        const int row = q.freel[t];
        int n_f = 0;
        int j0 = VIRT, i0 = row;
        double dmin = half - q.u[row];                     // the new row's own dummy
        int drow = row, dcol = VIRT;
        bool to_dummy = false;
        int guard = 0;
        while (true) {
            if (++guard > m + 2) return 2;
            const double ui0 = q.u[i0];
            // relax the admissible edges of i0
            const int e0 = q.off[i0], e1 = q.off[i0 + 1];
            for (int eb = e0; eb < e1; eb += 32) {
                const int e = eb + lane;
                bool fresh = false;
                int j = 0;
                if (e < e1) {
                    j = q.ecol[e];
                    if (!q.used[j]) {
                        const double cur = ((double)q.ecost[e] - half) - ui0 - q.v[j];
                        const double mv = q.minv[j];
                        fresh = !(mv < INF);
                        if (cur < mv) { q.minv[j] = cur; q.way[j] = j0; }
                    }
                }
                const unsigned fb = __ballot_sync(0xffffffffu, fresh);
                if (fresh) q.fcol[n_f + __popc(fb & ((1u << lane) - 1u))] = (unsigned short)j;
                n_f += __popc(fb);
            }
            __syncwarp();
            // the closest unvisited column of the frontier (lowest index among equals)
            double best = INF;
            int bj = -1;
            for (int f = lane; f < n_f; f += 32) {
                const int j = q.fcol[f];
                if (q.used[j]) continue;
                const double mv = q.minv[j];
                if (mv < best || (mv == best && j < bj)) { best = mv; bj = j; }
            }
            warp_argmin(best, bj);
            to_dummy = bj < 0 || dmin < best;
            const double delta = to_dummy ? dmin : best;
            if (!(delta < INF)) return 4;
            // dual update (only frontier columns have state)
            for (int f = lane; f < n_f; f += 32) {
                const int j = q.fcol[f];
                if (q.used[j]) { q.u[q.p[j]] += delta; q.v[j] -= delta; }
                else q.minv[j] -= delta;
            }
            if (lane == 0) q.u[row] += delta;
            dmin -= delta;
            __syncwarp();
            if (to_dummy) break;
            j0 = bj;
            if (lane == 0) q.used[j0] = 1;
            __syncwarp();
            i0 = q.p[j0];
            if (i0 < 0) break;                             // free real column: augment
            const double cand = half - q.u[i0];
            if (cand < dmin) { dmin = cand; drow = i0; dcol = j0; }
        }
        (void)drow;
        if (lane == 0) {
            int jc = to_dummy ? dcol : j0;                 // column that receives a new row (VIRT: none)
            int steps = 0;
            while (jc != VIRT && steps++ <= m) {
                const int jp = q.way[jc];
                q.p[jc] = jp == VIRT ? row : q.p[jp];
                jc = jp;
            }
        }
        __syncwarp();
        for (int f = lane; f < n_f; f += 32) { const int j = q.fcol[f]; q.minv[j] = INF; q.used[j] = 0; }
        __syncwarp();
    }
    const long long t2 = clock64();
    // ---- uniqueness certificate ----
    for (int i = lane; i < nR; i += 32) q.sigma[i] = nC + i;
    __syncwarp();
    for (int j = lane; j < nC; j += 32) { const int r = q.p[j]; if (r >= 0) q.sigma[r] = j; }
    __syncwarp();
    int n_arcs = 0;
    for (int eb = 0; eb < q.E; eb += 32) {
        const int e = eb + lane;
        bool tight = false;
        int from = 0, to = 0;
        if (e < q.E) {
            const int i = q.erow[e];
            to = q.ecol[e];
            from = q.sigma[i];
            tight = to != from && ((double)q.ecost[e] - half) - q.u[i] - q.v[to] <= kFastEps;
        }
        const unsigned tb = __ballot_sync(0xffffffffu, tight);
        if (tb == 0u) continue;
        const int pos = n_arcs + __popc(tb & ((1u << lane) - 1u));
        if (tight && pos < m) { q.arc_from[pos] = from; q.arc_to[pos] = to; }
        n_arcs += __popc(tb);
    }
    for (int i0 = 0; i0 < nR; i0 += 32) {                  // the dummy edge of every matched row
        const int i = i0 + lane;
        bool tight = false;
        int from = 0;
        if (i < nR) {
            from = q.sigma[i];
            if (from != nC + i) tight = half - q.u[i] <= kFastEps;
        }
        const unsigned tb = __ballot_sync(0xffffffffu, tight);
        const int pos = n_arcs + __popc(tb & ((1u << lane) - 1u));
        if (tight && pos < m) { q.arc_from[pos] = from; q.arc_to[pos] = nC + i; }
        n_arcs += __popc(tb);
    }
    if (n_arcs > m) return 5;
    __syncwarp();
    const int verdict = arcs_unique(nC, m, n_arcs, q.arc_from, q.arc_to, q.p, q.sigma,
                                    [&](int j) { return j < nC ? (double)q.v[j] : 0.0; }, q.flags, (long long *)nullptr);
    if (q.why && lane == 0) {
        q.why[2] = (int)(t1 - t0); q.why[3] = (int)(t2 - t1); q.why[4] = (int)(clock64() - t2);
        q.why[5] = n_free; q.why[6] = nR; q.why[7] = n_arcs;
    }
    __syncwarp();
    return verdict;
}

// The sparse solver for a whole problem, called by EVERY thread of the CTA (blockDim.x a multiple of 32;
// a one-warp CTA works too): drops the rows / columns without admissible entries, builds the CSR edge
// list from the dense matrix (P.row_cnt / col_cnt / adm_misc must be filled), runs sparse_assign_warp on
// warp 0 and writes x / y of the whole problem.  Everything lives in the part of the whole-problem
// scratch in front of x / y (v, d, cols, pred: `arena_bytes` = 24 bytes per node of its capacity).
// hdr: 4 ints of shared memory.  Returns the block-uniform verdict: 0 = solved with certificate,
// 1 / 6 / 7 = the dense certified path would refuse as well (NaN, near-limit cost, equal-cost
// alternative), 8 = does not fit, other = a guard.
template <typename F>
__device__ __forceinline__ int compact_by_warp(int n, F flag, int *dst) {
    const int lane = threadIdx.x & 31;
    int total = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const bool f = i < n && flag(i);
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (f) dst[total + __popc(b & ((1u << lane) - 1u))] = i;
        total += __popc(b);
    }
    __syncwarp();
    return total;
}

__device__ __forceinline__ int sparse_assign_block(const Problem &P, Scratch &s, size_t arena_bytes, int *hdr, int *why) {
    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, lane = tid & 31;
    const int R = P.R, C = P.C;
    const double L = 2.0 * P.half;
    if (P.adm_misc[0] != 0 || P.adm_misc[1] != 0) return 1;
    size_t at = ((size_t)8 * (R + C) + 15) & ~(size_t)15;
    if (at > arena_bytes) return 8;
    unsigned char *arena = reinterpret_cast<unsigned char *>(const_cast<double *>(s.v));
    int *rmap = reinterpret_cast<int *>(arena), *cmap = rmap + R, *rlist = cmap + C, *clist = rlist + R;
    for (int r = tid; r < R; r += nthr) s.x[r] = -1;
    for (int c = tid; c < C; c += nthr) s.y[c] = -1;
    if (warp == 0) {
        const int n = compact_by_warp(R, [&](int r) { return P.row_cnt[r] > 0; }, rlist);
        for (int k = lane; k < n; k += 32) rmap[rlist[k]] = k;
        if (lane == 0) hdr[0] = n;
    }
    if (warp == (nthr > 32 ? 1 : 0)) {
        const int n = compact_by_warp(C, [&](int c) { return P.col_cnt[c] > 0; }, clist);
        for (int k = lane; k < n; k += 32) cmap[clist[k]] = k;
        if (lane == 0) hdr[1] = n;
    }
    __syncthreads();
    const int nR = hdr[0], nC = hdr[1];
    int *cursor = reinterpret_cast<int *>(arena + at); at += ((size_t)4 * nR + 15) & ~(size_t)15;
    unsigned short *off = reinterpret_cast<unsigned short *>(arena + at); at += ((size_t)2 * (nR + 1) + 15) & ~(size_t)15;
    // edge offsets: exclusive prefix of the admissible counts of the kept rows (warp 0)
    if (warp == 0 && at <= arena_bytes) {
        int run = 0;
        for (int k0 = 0; k0 < nR; k0 += 32) {
            const int k = k0 + lane;
            const int cnt = k < nR ? P.row_cnt[rlist[k]] : 0;
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int up = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += up; }
            if (k < nR) { off[k] = (unsigned short)min(run + incl - cnt, 65535); cursor[k] = run + incl - cnt; }
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) { off[nR] = (unsigned short)min(run, 65535); hdr[2] = run; }
    }
    __syncthreads();
    if (at > arena_bytes) return 8;
    const int E = hdr[2];
    unsigned short *erow = reinterpret_cast<unsigned short *>(arena + at); at += ((size_t)2 * E + 15) & ~(size_t)15;
    unsigned short *ecol = reinterpret_cast<unsigned short *>(arena + at); at += ((size_t)2 * E + 15) & ~(size_t)15;
    float *ecost = reinterpret_cast<float *>(arena + at); at += ((size_t)4 * E + 15) & ~(size_t)15;
    if (at + sparse_state_bytes(nR, nC) > arena_bytes || E >= 65535 || nR + nC > 512) return 8;
    // the admissible entries of the matrix, grouped by row (the order inside a row does not matter)
    for (int e = tid; e < R * C; e += nthr) {
        const float dist = P.cost[e];
        if ((double)dist < L) {
            const int r = e / C, j = e - r * C;
            const int k = atomicAdd(&cursor[rmap[r]], 1);
            erow[k] = (unsigned short)rmap[r]; ecol[k] = (unsigned short)cmap[j]; ecost[k] = dist;
        }
    }
    __syncthreads();
    if (warp == 0) {
        Sparse q;
        q.nR = nR; q.nC = nC; q.E = E; q.half = P.half;
        q.off = off; q.erow = erow; q.ecol = ecol; q.ecost = ecost; q.why = why;
        sparse_carve_state(q, arena + at);
        const int code = nR == 0 ? 0 : sparse_assign_warp(q);
        if (code == 0)
            for (int j = lane; j < nC; j += 32) {
                const int rr = q.p[j];
                if (rr >= 0) { s.y[clist[j]] = rlist[rr]; s.x[rlist[rr]] = clist[j]; }
            }
        else if (why && lane == 0) { why[0] = code; atomicAdd(&why[1], 1); }
        if (lane == 0) hdr[3] = code;
    }
    __syncthreads();
    return hdr[3];
}

// ---------------------------------------------------------------------------------
// Component path (problems with R, C <= 64).  The objective of the extended problem is
// sum over matched pairs of (c_ij - L): only ADMISSIBLE entries (c_ij < L) can be part of an
// optimum, so rows in different connected components of the admissible bipartite graph are
// independent problems.  In tracking the graph is sparse: most rows are isolated pairs, the
// rest form components of two or three tracks around the same detections (98.8 % of the
// problems of the bench workload have only components <= 4 x 4, profiles/r01_assign_components_c2.txt).
//   1. one warp: isolated pairs are matched, the other rows are grouped into components by
//      closing row / column sets over the admissibility bitmasks (64-bit, REDUX per round);
//   2. every warp of the CTA takes components: a component of <= 4 rows and <= 4 columns is
//      solved by ENUMERATING its <= 209 partial matchings (5^rows codes spread over the lanes):
//      the minimum is the optimum, and the runner-up being worse by more than kFastEps is the
//      uniqueness certificate - a unique optimum is what any exact solver, including the
//      reference's, returns;
//   3. a larger component, a tie (duplicate boxes), a NaN or a cost within 2e-9 of L sends the
//      whole problem to assign_warp (certified Hungarian, then the literal JV emulation), so
//      tie-breaks still follow the reference's scan order.
struct Masks {
    unsigned long long *rmask;        // [64] admissible columns of every row
    unsigned long long *cmask;        // [64] admissible rows of every column
    unsigned long long *comp_rows;    // [64] components of <= 4 x 4 found by prepare_components
    unsigned long long *comp_cols;    // [64]
    unsigned long long *big;          // [2] union of the rows / columns of all larger components
    int *n_comp;                      // [0] number of small components; [2] number of larger ones (block version:
                                      //     listed one by one in comp_rows / comp_cols[63 - k])
    int *fallback;                    // [1] != 0: the component path does not apply / failed
};

constexpr int kSubMax = 32;           // rows / columns of the sub-problem made of the larger components

__host__ __device__ inline size_t masks_bytes() { return sizeof(unsigned long long) * (64 * 4 + 2) + 16; }

__device__ inline Masks carve_masks(unsigned char *p) {
    Masks m;
    m.rmask = reinterpret_cast<unsigned long long *>(p);
    m.cmask = m.rmask + 64;
    m.comp_rows = m.cmask + 64;
    m.comp_cols = m.comp_rows + 64;
    m.big = m.comp_cols + 64;
    m.n_comp = reinterpret_cast<int *>(m.big + 2);
    m.fallback = m.n_comp + 1;
    return m;
}

__device__ __forceinline__ unsigned long long warp_or64(unsigned long long v) {
    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
    const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
    return ((unsigned long long)hi << 32) | lo;
}

// One warp builds the bitmasks from the dense matrix (the track kernel builds them while it
// gathers the matrix; this is for the stand-alone batch entry).
__device__ inline void build_masks(const Problem &P, const Masks &m) {
    const int lane = threadIdx.x & 31;
    const double L = 2.0 * P.half;
    for (int i = lane; i < 64; i += 32) { m.rmask[i] = 0ull; m.cmask[i] = 0ull; }
    __syncwarp();
    for (int i = lane; i < P.R; i += 32) {
        unsigned long long rm = 0ull;
        for (int j = 0; j < P.C; ++j)
            if ((double)P.cost[(long long)i * P.C + j] < L) { rm |= 1ull << j; atomicOr(&m.cmask[j], 1ull << i); }
        m.rmask[i] = rm;
    }
    __syncwarp();
}

// Step 1, by ONE warp.  Writes x / y for the isolated pairs (-1 elsewhere), the component list
// and the fallback flag.  P.adm_misc must be filled (NaN / near-limit screens).  Components by
// min-label propagation: every active row starts with its own index, a column takes the minimum
// label of its rows, a row the minimum of its columns, until nothing changes (rounds ~ diameter of
// the largest component, two rows per lane); rows with equal labels form a component.
__device__ inline void prepare_components(const Problem &P, Scratch &s, const Masks &m) {
    const int lane = threadIdx.x & 31;
    const int R = P.R, C = P.C;
    for (int j = lane; j < C; j += 32) s.y[j] = -1;
    __syncwarp();
    unsigned long long rm[2], cm[2];
    int label[2];
    bool act[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int r = lane + 32 * h;
        rm[h] = r < R ? m.rmask[r] : 0ull;
        cm[h] = r < C ? m.cmask[r] : 0ull;            // (this lane's two columns)
        act[h] = false;
        if (r < R) {
            int xr = -1;
            if (rm[h]) {
                const int c = __ffsll((long long)rm[h]) - 1;
                if ((rm[h] & (rm[h] - 1)) == 0ull && m.cmask[c] == (1ull << r)) { xr = c; s.y[c] = r; }   // isolated pair
                else act[h] = true;
            }
            s.x[r] = xr;
        }
        label[h] = act[h] ? r : 64;
    }
    bool fb = P.adm_misc[0] != 0 || P.adm_misc[1] != 0;
    // labels live in registers; a lane reads another row's / column's label by shuffle
    int clab[2] = {64, 64};
    for (int round = 0; round < 130; ++round) {
        // column labels: minimum over the column's rows
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            unsigned long long t = cm[h];
            int best = 64;
            // all lanes step together: the shuffles need every lane, so iterate to the longest set
            int steps = __reduce_max_sync(0xffffffffu, __popcll(t));
            for (int k = 0; k < steps; ++k) {
                const int r = t ? __ffsll((long long)t) - 1 : 0;
                const int l0 = __shfl_sync(0xffffffffu, label[0], r & 31), l1 = __shfl_sync(0xffffffffu, label[1], r & 31);
                if (t) best = min(best, r < 32 ? l0 : l1);
                t &= t - 1;
            }
            clab[h] = best;
        }
        bool changed = false;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            unsigned long long t = act[h] ? rm[h] : 0ull;
            int best = label[h];
            int steps = __reduce_max_sync(0xffffffffu, __popcll(t));
            for (int k = 0; k < steps; ++k) {
                const int c = t ? __ffsll((long long)t) - 1 : 0;
                const int l0 = __shfl_sync(0xffffffffu, clab[0], c & 31), l1 = __shfl_sync(0xffffffffu, clab[1], c & 31);
                if (t) best = min(best, c < 32 ? l0 : l1);
                t &= t - 1;
            }
            if (best < label[h]) { label[h] = best; changed = true; }
        }
        if (!__any_sync(0xffffffffu, changed)) break;
    }
    // roots = active rows that kept their own index; one component per root, in ascending order
    unsigned long long roots = (unsigned long long)__ballot_sync(0xffffffffu, act[0] && label[0] == lane) |
                               ((unsigned long long)__ballot_sync(0xffffffffu, act[1] && label[1] == lane + 32) << 32);
    int n = 0;
    unsigned long long big_rows = 0ull, big_cols = 0ull;
    while (roots) {
        const int root = __ffsll((long long)roots) - 1;
        roots &= roots - 1;
        const unsigned long long rows = (unsigned long long)__ballot_sync(0xffffffffu, label[0] == root) |
                                        ((unsigned long long)__ballot_sync(0xffffffffu, label[1] == root) << 32);
        const unsigned long long cols = warp_or64((label[0] == root ? rm[0] : 0ull) | (label[1] == root ? rm[1] : 0ull));
        if (__popcll(rows) > 4 || __popcll(cols) > 4) {
            big_rows |= rows; big_cols |= cols;           // solved together as one compact sub-problem
        } else {
            if (lane == 0) { m.comp_rows[n] = rows; m.comp_cols[n] = cols; }
            ++n;
        }
    }
    if (lane == 0) { *m.n_comp = n; *m.fallback = fb ? 1 : 0; m.big[0] = big_rows; m.big[1] = big_cols; }
    __syncwarp();
}

// A component of at most 2 rows and 2 columns (four out of five components in tracking) in closed
// form, by ONE thread: the empty matching, the up to four single pairs and the up to two pairings
// are its seven candidates; the best wins if the runner-up is worse by more than kFastEps.
// Writes x / y of its rows and columns; false = tie.
__device__ inline bool solve_tiny_idx(const Problem &P, Scratch &s, int r0, int r1, int c0, int c1);

__device__ inline bool solve_tiny_component(const Problem &P, Scratch &s, unsigned long long rows, unsigned long long cols) {
    const int r0 = __ffsll((long long)rows) - 1, c0 = __ffsll((long long)cols) - 1;
    rows &= rows - 1; cols &= cols - 1;
    const int r1 = rows ? __ffsll((long long)rows) - 1 : -1, c1 = cols ? __ffsll((long long)cols) - 1 : -1;
    return solve_tiny_idx(P, s, r0, r1, c0, c1);
}

// rows r0 (, r1) and columns c0 (, c1); -1 = absent
__device__ inline bool solve_tiny_idx(const Problem &P, Scratch &s, int r0, int r1, int c0, int c1) {
    const double L = 2.0 * P.half, INF = 1.0e300;
    auto a = [&](int r, int c) -> double {
        if (r < 0 || c < 0) return INF;
        const double v = (double)P.cost[(long long)r * P.C + c];
        return v < L ? v - L : INF;
    };
    const double a00 = a(r0, c0), a01 = a(r0, c1), a10 = a(r1, c0), a11 = a(r1, c1);
    // candidate k: (value, column of r0, column of r1)
    const double val[7] = {0.0, a00, a01, a10, a11, a00 + a11, a01 + a10};
    double best = INF, second = INF;
    int bk = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const double v = val[k];
        if (!(v < INF)) continue;
        if (v < best) { second = best; best = v; bk = k; }
        else if (v < second) second = v;
    }
    if (!((second - best) > kFastEps)) return false;
    const int x0 = (bk == 1 || bk == 5) ? c0 : ((bk == 2 || bk == 6) ? c1 : -1);
    const int x1 = (bk == 3 || bk == 6) ? c0 : ((bk == 4 || bk == 5) ? c1 : -1);
    s.x[r0] = x0;
    if (x0 >= 0) s.y[x0] = r0;
    if (r1 >= 0) {
        s.x[r1] = x1;
        if (x1 >= 0) s.y[x1] = r1;
    }
    return true;
}

// Step 1 for a whole CTA (>= 64 threads, every thread calls it; it ends with a barrier): the same
// outputs as prepare_components, except that components of <= 2 x 2 are solved on the spot by the
// thread that finds them (solve_tiny_component) and do not enter the component list.  One thread per row: the rows sharing a column with it, then the
// transitive closure of that relation by repeated squaring on 64-bit row sets in shared memory
// (rounds ~ log2 of the largest component), instead of a label propagation whose every step is a
// chain of dependent shuffles on one warp.  m.comp_rows doubles as the closure's work array.
__device__ inline void prepare_components_block(const Problem &P, Scratch &s, const Masks &m) {
    const int t = threadIdx.x;
    const int R = P.R, C = P.C;
    __shared__ unsigned long long adj[64];
    __shared__ int changed_sh;
    for (int j = t; j < C; j += blockDim.x) s.y[j] = -1;
    if (t == 0) {
        *m.n_comp = 0; m.n_comp[2] = 0; m.big[0] = 0ull; m.big[1] = 0ull;
        *m.fallback = (P.adm_misc[0] != 0 || P.adm_misc[1] != 0) ? 1 : 0;
    }
    __syncthreads();
    bool act = false;
    unsigned long long rm = 0ull, mine = 0ull;
    if (t < 64) {
        if (t < R) {
            rm = m.rmask[t];
            int xr = -1;
            if (rm) {
                const int c = __ffsll((long long)rm) - 1;
                if ((rm & (rm - 1)) == 0ull && m.cmask[c] == (1ull << t)) { xr = c; s.y[c] = t; }   // isolated pair
                else act = true;
            }
            s.x[t] = xr;
        }
        if (act) {                                         // rows sharing a column with this one (itself included)
            unsigned long long q = rm;
            while (q) { mine |= m.cmask[__ffsll((long long)q) - 1]; q &= q - 1; }
        }
        adj[t] = mine;
    }
    __syncthreads();
    for (int round = 0; round < 7; ++round) {             // 2^7 >= 64 rows: enough for any chain
        if (t == 0) changed_sh = 0;
        __syncthreads();
        unsigned long long next = mine;
        if (t < 64 && act) {
            unsigned long long q = mine;
            while (q) { next |= adj[__ffsll((long long)q) - 1]; q &= q - 1; }
            if (next != mine) changed_sh = 1;
        }
        __syncthreads();
        if (t < 64) { adj[t] = next; mine = next; }
        __syncthreads();
        if (!changed_sh) break;
    }
    // the lowest row of a component publishes it
    if (t < 64 && act && (mine & (0ull - mine)) == (1ull << t)) {
        unsigned long long cols = 0ull, q = mine;
        while (q) { cols |= m.rmask[__ffsll((long long)q) - 1]; q &= q - 1; }
        if (__popcll(mine) > 4 || __popcll(cols) > 4) {
            atomicOr(&m.big[0], mine);                     // the union, for a caller that solves them as one sub-problem
            atomicOr(&m.big[1], cols);
            const int kb = atomicAdd(m.n_comp + 2, 1);     // and one by one, from the end of the component list
            m.comp_rows[63 - kb] = mine;                   // (components have distinct rows: small + larger ones <= 64)
            m.comp_cols[63 - kb] = cols;
        } else if (__popcll(mine) <= 2 && __popcll(cols) <= 2 && *m.fallback == 0) {
            if (!solve_tiny_component(P, s, mine, cols)) *m.fallback = 1;      // a tie: the exact path decides
        } else {
            const int k = atomicAdd(m.n_comp, 1);
            m.comp_rows[k] = mine;
            m.comp_cols[k] = cols;
        }
    }
    __syncthreads();
    __syncthreads();
}

// Step 2, by one warp per component: enumerate the partial matchings of a component of <= 4 rows
// and <= 4 columns.  Writes x / y of its rows and columns when the optimum is unique; returns false
// on a tie (within kFastEps).
__device__ inline bool enumerate_idx(const Problem &P, Scratch &s, const int (&ri)[4], const int (&ci)[4], int k, int mcols);

__device__ inline bool enumerate_component(const Problem &P, Scratch &s, unsigned long long rows, unsigned long long cols) {
    int ri[4], ci[4];
    const int k = __popcll(rows), mcols = __popcll(cols);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ri[i] = rows ? __ffsll((long long)rows) - 1 : 0;
        rows &= rows - 1;
        ci[i] = cols ? __ffsll((long long)cols) - 1 : 0;
        cols &= cols - 1;
    }
    return enumerate_idx(P, s, ri, ci, k, mcols);
}

// the same for a component given by its row / column indices (k <= 4 rows, mcols <= 4 columns)
__device__ inline bool enumerate_idx(const Problem &P, Scratch &s, const int (&ri)[4], const int (&ci)[4], int k, int mcols) {
    const int lane = threadIdx.x & 31;
    const double L = 2.0 * P.half, INF = 1.0e300;
    double a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double v = INF;
            if (i < k && j < mcols) {
                const double c = (double)P.cost[(long long)ri[i] * P.C + ci[j]];
                if (c < L) v = c - L;
            }
            a[i][j] = v;
        }
    // digits of a code: the option of each row, 0 = unmatched, j + 1 = column j; base = columns + 1
    const int base = mcols + 1;
    int total = 1;
    for (int i = 0; i < k; ++i) total *= base;
    double best = INF, second = INF;
    int best_code = -1;
    for (int code = lane; code < total; code += 32) {
        int t = code;
        unsigned used = 0u;
        double obj = 0.0;
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int dgt = t % base;
            t /= base;
            if (i < k && dgt) {
                const int j = dgt - 1;
                double v = INF;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) v = (jj == j) ? a[i][jj] : v;
                if ((used >> j) & 1u || !(v < INF)) ok = false;
                used |= 1u << j;
                obj += v;
            }
        }
        if (!ok) continue;
        if (obj < best) { second = best; best = obj; best_code = code; }
        else if (obj < second) second = obj;
    }
    const double B = warp_min(best);
    const unsigned holders = __ballot_sync(0xffffffffu, best == B);
    const int holder = __ffs(holders) - 1;
    const double S = warp_min(lane == holder ? second : best);
    const bool unique = __popc(holders) == 1 && (S - B) > kFastEps;
    if (!unique) return false;
    const int code = __shfl_sync(0xffffffffu, best_code, holder);
    if (lane == 0) {
        int t = code;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int dgt = t % base;
            t /= base;
            if (i < k) {
                int c = -1;
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) c = (dgt == jj + 1) ? ci[jj] : c;
                s.x[ri[i]] = c;
                if (c >= 0) s.y[c] = ri[i];
            }
        }
    }
    __syncwarp();
    return true;
}

// The components larger than 4 x 4 (crowded groups) together as ONE sub-problem for the sparse certified
// solver on one warp, while the other warps enumerate the small components.  Rows of these components
// have admissible entries only in their own columns, so the sub-problem's unique optimum is the
// restriction of the whole problem's.  The edge list comes straight from the 64-bit admissibility masks
// (at most kSubEdges edges, else the whole-problem path); `sub_mem` (sub_bytes()) holds the CSR arrays,
// the compact -> whole index lists and the solver's state.  Writes x / y of the whole problem for its
// rows and columns; false = no certificate.
// (Round 2 first copied a dense <= 32 x 32 sub-matrix and ran the dense certified Hungarian on it: ~20 k
// cycles for a component of 5-9 tracks, most of them in the admissible-edge statistics, the per-column
// state set-up and the certificate's column walk of the dense form.)
__host__ __device__ inline size_t sub_bytes() {
    return scratch_bytes(2 * kSubMax) + sizeof(float) * kSubMax * kSubMax + sizeof(int) * (5 * kSubMax + 4);
}

constexpr int kSubEdges = 256;         // admissible edges the sub-problem of the larger components may hold
constexpr size_t kSubFront = 144 + 2 * 2 * kSubEdges + 4 * kSubEdges + 128;     // off | erow | ecol | ecost | lrow, lcol

__device__ inline bool solve_component_sparse(const Problem &P, Scratch &s, const Masks &m, unsigned long long rows,
                                              unsigned long long cols, unsigned char *sub_mem);

__device__ inline bool solve_big_components(const Problem &P, Scratch &s, const Masks &m, unsigned char *sub_mem) {
    return solve_component_sparse(P, s, m, m.big[0], m.big[1], sub_mem);
}

// one component (or a union of components) given by its row / column sets; sub_mem: sub_bytes() of scratch
__device__ inline bool solve_component_sparse(const Problem &P, Scratch &s, const Masks &m, unsigned long long rows,
                                              unsigned long long cols, unsigned char *sub_mem) {
    static_assert(kSubFront % 16 == 0, "state must start 16-byte aligned");
    const int lane = threadIdx.x & 31;
    if (!rows) return true;
    const int nR = __popcll(rows), nC = __popcll(cols);
    if (kSubFront + sparse_state_bytes(nR, nC) > sub_bytes()) return false;
    unsigned short *off = reinterpret_cast<unsigned short *>(sub_mem);
    unsigned short *erow = off + 72, *ecol = erow + kSubEdges;
    float *ecost = reinterpret_cast<float *>(ecol + kSubEdges);
    unsigned char *lrow = reinterpret_cast<unsigned char *>(ecost + kSubEdges), *lcol = lrow + 64;
    // compact numbering: the k-th set bit of `rows` / `cols`; lane l takes rows l and l + 32
    int gr[2] = {0, 0}, deg[2] = {0, 0};
    unsigned long long adj[2] = {0ull, 0ull};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        if (k < nR) {
            unsigned long long t = rows;
            for (int q = 0; q < k; ++q) t &= t - 1;
            gr[h] = __ffsll((long long)t) - 1;
            adj[h] = m.rmask[gr[h]] & cols;
            deg[h] = __popcll(adj[h]);
            lrow[k] = (unsigned char)gr[h];
        }
        if (k < nC) {
            unsigned long long t = cols;
            for (int q = 0; q < k; ++q) t &= t - 1;
            lcol[k] = (unsigned char)(__ffsll((long long)t) - 1);
        }
    }
    int run = 0, first[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        int incl = deg[h];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int up = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += up; }
        first[h] = run + incl - deg[h];
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    const int E = run;
    if (E > kSubEdges) return false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = lane + 32 * h;
        if (k < nR) {
            off[k] = (unsigned short)first[h];
            unsigned long long a = adj[h];
            int e = first[h];
            while (a) {
                const int j = __ffsll((long long)a) - 1;
                a &= a - 1;
                erow[e] = (unsigned short)k;
                ecol[e] = (unsigned short)__popcll(cols & ((1ull << j) - 1ull));
                ecost[e] = P.cost[(long long)gr[h] * P.C + j];
                ++e;
            }
            s.x[gr[h]] = -1;
        }
    }
    if (lane == 0) off[nR] = (unsigned short)E;
    __syncwarp();
    Sparse q;
    q.nR = nR; q.nC = nC; q.E = E; q.half = P.half;
    q.off = off; q.erow = erow; q.ecol = ecol; q.ecost = ecost; q.why = P.why;
    sparse_carve_state(q, sub_mem + kSubFront);
    if (sparse_assign_warp(q) != 0) return false;
    for (int j = lane; j < nC; j += 32) {
        const int r = q.p[j];
        s.y[lcol[j]] = r >= 0 ? (int)lrow[r] : -1;
        if (r >= 0) s.x[lrow[r]] = lcol[j];
    }
    __syncwarp();
    return true;
}

// The whole component path for a problem of at most 32 x 32 on ONE warp, without a block
// barrier (stages 2 and 3 of the tracker are a handful of rows and columns; a CTA-wide barrier
// costs more than their whole solution): `masks` must hold the admissibility bitmasks (low 32
// bits used).  Isolated pairs, closure of the row adjacency by repeated squaring with one row
// per lane, components of <= 2 x 2 in closed form by their root lanes, components of <= 4 x 4
// enumerated by the warp one after the other, larger ones as the compact certified
// sub-problem.  Returns false when the problem needs the whole-problem path (tie, NaN, ...).
__device__ inline bool solve_small_warp(const Problem &P, Scratch &s, const Masks &m, unsigned char *sub_mem) {
    const int lane = threadIdx.x & 31;
    const int R = P.R, C = P.C;
    if (lane < C) s.y[lane] = -1;
    __syncwarp();
    const unsigned rm = lane < R ? (unsigned)m.rmask[lane] : 0u;
    bool act = false;
    if (lane < R) {
        int xr = -1;
        if (rm) {
            const int c = __ffs(rm) - 1;
            if ((rm & (rm - 1)) == 0u && (unsigned)m.cmask[c] == (1u << lane)) { xr = c; s.y[c] = lane; }   // isolated pair
            else act = true;
        }
        s.x[lane] = xr;
    }
    bool bad = P.adm_misc[0] != 0 || P.adm_misc[1] != 0;
    // rows sharing a column with this lane's row, then the transitive closure (work array: comp_rows)
    unsigned mine = 0u;
    if (act) { unsigned q = rm; while (q) { mine |= (unsigned)m.cmask[__ffs(q) - 1]; q &= q - 1; } }
    m.comp_rows[lane] = mine;
    __syncwarp();
    for (int round = 0; round < 6; ++round) {
        unsigned next = mine, q = mine;
        while (q) { next |= (unsigned)m.comp_rows[__ffs(q) - 1]; q &= q - 1; }
        const bool changed = __any_sync(0xffffffffu, next != mine);
        __syncwarp();
        m.comp_rows[lane] = next;
        mine = next;
        __syncwarp();
        if (!changed) break;
    }
    const bool root = act && (mine & (0u - mine)) == (1u << lane);
    unsigned cols = 0u;
    if (root) { unsigned q = mine; while (q) { cols |= (unsigned)m.rmask[__ffs(q) - 1]; q &= q - 1; } }
    const int nr = __popc(mine), nc = __popc(cols);
    const bool tiny = root && nr <= 2 && nc <= 2, big = root && (nr > 4 || nc > 4), mid = root && !tiny && !big;
    if (tiny && !solve_tiny_component(P, s, mine, cols)) bad = true;
    bad = __any_sync(0xffffffffu, bad);
    unsigned mids = __ballot_sync(0xffffffffu, mid);
    while (mids && !bad) {
        const int src = __ffs(mids) - 1;
        mids &= mids - 1;
        const unsigned rr = __shfl_sync(0xffffffffu, mine, src), cc = __shfl_sync(0xffffffffu, cols, src);
        if (!enumerate_component(P, s, rr, cc)) bad = true;
    }
    const unsigned big_rows = __reduce_or_sync(0xffffffffu, big ? mine : 0u), big_cols = __reduce_or_sync(0xffffffffu, big ? cols : 0u);
    if (big_rows && !bad) {
        if (lane == 0) { m.big[0] = big_rows; m.big[1] = big_cols; }
        __syncwarp();
        if (!solve_big_components(P, s, m, sub_mem)) bad = true;
    }
    __syncwarp();
    return !bad;
}

// Fast path with certified uniqueness, else the exact emulation of the reference's solver.
__device__ inline bool assign_warp(const Problem &P, Scratch &s, int *fallback_counter) {
    if (P.n <= 512 && fast_assign_warp(P, s)) return true;
    if (fallback_counter && (threadIdx.x & 31) == 0) *fallback_counter += 1;
    __syncwarp();
    return solve_warp(P, s);
}

}  // namespace stlap
