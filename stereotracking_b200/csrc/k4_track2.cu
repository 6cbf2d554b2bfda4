// K4 + K5, round-2 kernel: one frame of the OC-SORT tracker with depth-scaled boxes for S
// sequences in lockstep, one CTA per sequence, the whole frame in one launch.  Same semantics
// and arithmetic as k4_track.cu (which it replaces as the default; ST_K4_V1=1 selects the old
// one), restructured around what bounded that kernel - dependent latencies, not bandwidth:
//   * the frame's working set (per-track scalars, observations, detection rows) is fetched ONCE
//     into shared memory in one round of independent loads; every later phase works on shared
//     memory and the write-back is plain stores (the old kernel walked slot -> counter ->
//     ring entry chains of dependent global loads in every stage and again in the update);
//   * the association costs of all three stages are computed in ONE pass over all
//     (live track, filtered detection) pairs - OCM cost with the velocity-direction term and
//     OCR cost on the last observations - instead of three serial row-data + cost phases with a
//     float64 acos each; a stage then only gathers its sub-matrix;
//   * assignment by connected components of the admissible graph (lapjv.cuh): isolated pairs
//     directly, components of <= 4 x 4 enumerated one per warp with the uniqueness certificate
//     from the same enumeration; anything else takes the certified Hungarian / literal JV path;
//   * the online-smoothing chain of a re-found track runs in the same 8-lane group, back to
//     back with that track's Kalman update, while the other groups update their own tracks
//     (it was a phase of its own that the whole CTA waited for);
//   * compactions of independent lists run on different warps between the same two barriers;
//     diagnostics are fire-and-forget.
// Replaces OCSORTTracker_Disparity.track and everything it calls:
//   mmtrack/models/trackers/ocsort_tracker_disparity.py:105-185,:187-343,:345-618
//   mmtrack/models/trackers/kalman_tracker_base.py:50-88
//   mmtrack/models/trackers/base_tracker.py:34-121
//   mmtrack/models/motion/kalman_filter.py:60-189
#include "st_common.cuh"
#include "kalman.cuh"
#include "lapjv.cuh"
#include "k4_common.cuh"

using namespace stk4;

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
// shared-memory floats for the cost matrices (all-pairs OCM + OCR, and the dense matrix of the
// current stage) when the capacities leave room; larger problems use the global workspace
constexpr int kCostFloatsMax = 14336;
constexpr size_t kSmemBudget = 220 * 1024;

struct Smem {
    stlap::Scratch lap;
    stlap::Masks masks;
    // per ORDER position (live tracks in dict order)
    long long *id;
    int *slot, *last_frame, *num_obs, *miss_run, *hits;
    unsigned char *tent, *tracked0, *hit, *valid, *knext_ok, *has_saved;
    float *box_ocm;              // [T,4] box from the Kalman mean (after predict for confirmed tracks)
    float *box_ocr;              // [T,4] last observation
    float *knext;                // [T,4] observation the velocity of a matched track is taken against
    float *vel;                  // [T,2]
    float *kc;                   // [T,2] centre of the k-step observation
    // per input detection
    float *in_box;               // [D,4]
    float *in_score;             // [D]
    int *det_pos;                // [D] matched order position or -1
    int *out_src;                // [D] output order -> input det index
    int *dets;                   // [D] filtered column -> input det index
    int *scol;                   // [D] stage column -> filtered column
    unsigned char *det_stage;    // [D] 0 dropped, 1..3 matched in stage, 4 leftover
    // stage
    int *rows;                   // [T] stage-3 row -> order position
    int *conf, *tentl;           // [T] confirmed / tentative order positions (stage-1 / stage-2 rows)
    int *row_cnt, *row_adm, *col_cnt, *adm_misc;
    unsigned char *sub;          // scratch of the sub-problem made of the larger components (lapjv.cuh)
    float *cost;                 // [cost_cap]
    int cost_cap;
};

__host__ __device__ inline size_t smem2_fixed_bytes(int T, int D) {
    size_t b = stlap::scratch_bytes(T + D) + align16(stlap::masks_bytes());
    b += align16(sizeof(long long) * T) + align16(sizeof(int) * T) * 5 + align16(T) * 6;
    b += align16(sizeof(float) * T * 4) * 3 + align16(sizeof(float) * T * 2) * 2;
    b += align16(sizeof(float) * D * 4) + align16(sizeof(float) * D) + align16(sizeof(int) * D) * 4 + align16(D);
    b += align16(sizeof(int) * T) * 5 + align16(sizeof(int) * D) + 16;
    b += align16(stlap::sub_bytes());
    return b;
}

// floats of the cost region for these capacities (0 = the fixed part alone does not fit)
__host__ inline int cost_floats_for(int T, int D) {
    const size_t fixed = smem2_fixed_bytes(T, D);
    // the camera-motion step borrows the region as its per-warp scratch
    const size_t min_bytes = sizeof(double) * kWarps * stkf::kScratchDoubles;
    if (fixed + min_bytes > kSmemBudget) return 0;
    size_t fl = (kSmemBudget - fixed) / sizeof(float);
    if (fl > (size_t)kCostFloatsMax) fl = kCostFloatsMax;
    return (int)(fl & ~(size_t)3);
}

__device__ inline Smem carve2(unsigned char *p, int T, int D, int cost_floats) {
    Smem m;
    auto take = [&](size_t bytes) { unsigned char *q = p; p += align16(bytes); return q; };
    m.lap = stlap::carve(p, T + D); p += stlap::scratch_bytes(T + D);
    m.masks = stlap::carve_masks(take(stlap::masks_bytes()));
    m.id = reinterpret_cast<long long *>(take(sizeof(long long) * T));
    m.slot = reinterpret_cast<int *>(take(sizeof(int) * T));
    m.last_frame = reinterpret_cast<int *>(take(sizeof(int) * T));
    m.num_obs = reinterpret_cast<int *>(take(sizeof(int) * T));
    m.miss_run = reinterpret_cast<int *>(take(sizeof(int) * T));
    m.hits = reinterpret_cast<int *>(take(sizeof(int) * T));
    m.tent = take(T); m.tracked0 = take(T); m.hit = take(T); m.valid = take(T); m.knext_ok = take(T);
    m.has_saved = take(T);
    m.box_ocm = reinterpret_cast<float *>(take(sizeof(float) * T * 4));
    m.box_ocr = reinterpret_cast<float *>(take(sizeof(float) * T * 4));
    m.knext = reinterpret_cast<float *>(take(sizeof(float) * T * 4));
    m.vel = reinterpret_cast<float *>(take(sizeof(float) * T * 2));
    m.kc = reinterpret_cast<float *>(take(sizeof(float) * T * 2));
    m.in_box = reinterpret_cast<float *>(take(sizeof(float) * D * 4));
    m.in_score = reinterpret_cast<float *>(take(sizeof(float) * D));
    m.det_pos = reinterpret_cast<int *>(take(sizeof(int) * D));
    m.out_src = reinterpret_cast<int *>(take(sizeof(int) * D));
    m.dets = reinterpret_cast<int *>(take(sizeof(int) * D));
    m.scol = reinterpret_cast<int *>(take(sizeof(int) * D));
    m.det_stage = take(D);
    m.rows = reinterpret_cast<int *>(take(sizeof(int) * T));
    m.conf = reinterpret_cast<int *>(take(sizeof(int) * T));
    m.tentl = reinterpret_cast<int *>(take(sizeof(int) * T));
    m.row_cnt = reinterpret_cast<int *>(take(sizeof(int) * T));
    m.row_adm = reinterpret_cast<int *>(take(sizeof(int) * T));
    m.col_cnt = reinterpret_cast<int *>(take(sizeof(int) * D));
    m.adm_misc = reinterpret_cast<int *>(take(16));
    m.sub = take(stlap::sub_bytes());
    m.cost = reinterpret_cast<float *>(p);
    m.cost_cap = cost_floats;
    return m;
}

// ordered (stable) compaction of the indices i in [0,n) with flag(i) into dst by ONE warp (all 32
// lanes call it); the count goes to *cnt_sh.  No barrier inside: callers run several of these on
// different warps and synchronise once.
template <typename F>
__device__ __forceinline__ void compact_warp(int n, F flag, int *dst, int *cnt_sh) {
    const int lane = threadIdx.x & 31;
    int total = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        const bool f = i < n && flag(i);
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (f) dst[total + __popc(b & ((1u << lane) - 1u))] = i;
        total += __popc(b);
    }
    if (lane == 0) *cnt_sh = total;
}

// track box of the association from a Kalman mean: cxcyah_to_xyxy(float32(mean[:4])) (:217-223)
__device__ __forceinline__ void mean_to_box(float cx, float cy, float a, float h, float *o) {
    const float hw = __fmul_rn(__fmul_rn(a, h), 0.5f), hh = __fmul_rn(h, 0.5f);
    o[0] = __fsub_rn(cx, hw); o[1] = __fsub_rn(cy, hh); o[2] = __fadd_rn(cx, hw); o[3] = __fadd_rn(cy, hh);
}

// OCM cost of (track at order position p, input detection i) exactly as the reference evaluates it
// (:187-258): 1 - IoU (x score) + vel_consist_weight * (acos(cos) - pi/2) / pi for rows with a
// valid velocity and k-step observation.
// `lazy`: most pairs of a frame do not overlap at all.  For such a pair IoU is exactly 0, the cost is
// 1 + w (angle - pi/2) / pi, and it can only come near the assignment's cost limit when the
// detection lies almost exactly in the track's direction of motion: if the direction cosine
// (one approximate square root, no division) is below pr.screen_cos the cost exceeds the limit by
// more than 1e-4 whatever its last bits are, and pr.placeholder is returned instead.  Such an
// entry can never be matched, and only the literal JV emulation (exact ties) looks at its value:
// associate2 re-evaluates the matrix exactly before it runs that path.  *lazy_used is set when a
// placeholder was returned.  (The all-pairs pass was issue-bound on the five IEEE divisions, two
// square roots and the float64 acos per pair; a screened pair costs ~30 instructions.)
// The screen of ocm_cost alone: returns false and the final value (1 or the placeholder) for a pair the lazy
// rule settles, true when the pair needs the full evaluation.
__device__ __forceinline__ bool ocm_screen(const Params &pr, const Smem &sm, int p, int i, float *val, int *lazy_used) {
    const float *tb = sm.box_ocm + p * 4, *db = sm.in_box + i * 4;
    float w = __fsub_rn(fminf(tb[2], db[2]), fmaxf(tb[0], db[0]));
    float h = __fsub_rn(fminf(tb[3], db[3]), fmaxf(tb[1], db[1]));
    w = w < 0.f ? 0.f : w;
    h = h < 0.f ? 0.f : h;
    const float inter = __fmul_rn(w, h);
    const float a12 = __fadd_rn(__fmul_rn(__fsub_rn(tb[2], tb[0]), __fsub_rn(tb[3], tb[1])),
                                __fmul_rn(__fsub_rn(db[2], db[0]), __fsub_rn(db[3], db[1])));
    const float sc = pr.weight_iou ? sm.in_score[i] : 1.f;
    if (inter == 0.f && fabsf(a12) < 3.0e38f && fabsf(sc) < 3.0e38f) {
        if (!sm.valid[p]) { *val = 1.f; return false; }
        const float dcx = __fmul_rn(__fadd_rn(db[0], db[2]), 0.5f), dcy = __fmul_rn(__fadd_rn(db[1], db[3]), 0.5f);
        const float dy = __fsub_rn(dcy, sm.kc[p * 2 + 1]), dx = __fsub_rn(dcx, sm.kc[p * 2]);
        const float dot = dy * sm.vel[p * 2] + dx * sm.vel[p * 2 + 1];
        const float norm = sqrtf(dy * dy + dx * dx) + 1e-6f;
        if (dot < pr.screen_cos * norm) { *lazy_used = 1; *val = pr.placeholder; return false; }
    }
    return true;
}

__device__ __forceinline__ float ocm_cost(const Params &pr, const Smem &sm, int p, int i, bool lazy, int *lazy_used) {
    const float *tb = sm.box_ocm + p * 4, *db = sm.in_box + i * 4;
    if (lazy) {
        float w = __fsub_rn(fminf(tb[2], db[2]), fmaxf(tb[0], db[0]));
        float h = __fsub_rn(fminf(tb[3], db[3]), fmaxf(tb[1], db[1]));
        w = w < 0.f ? 0.f : w;
        h = h < 0.f ? 0.f : h;
        const float inter = __fmul_rn(w, h);
        const float a12 = __fadd_rn(__fmul_rn(__fsub_rn(tb[2], tb[0]), __fsub_rn(tb[3], tb[1])),
                                    __fmul_rn(__fsub_rn(db[2], db[0]), __fsub_rn(db[3], db[1])));
        const float sc = pr.weight_iou ? sm.in_score[i] : 1.f;
        // inter == 0 with a finite union and score: IoU (x score) is exactly +-0 and 1 - it exactly 1
        if (inter == 0.f && fabsf(a12) < 3.0e38f && fabsf(sc) < 3.0e38f) {
            if (!sm.valid[p]) return 1.f;                 // no direction term: this IS the exact cost
            const float dcx = __fmul_rn(__fadd_rn(db[0], db[2]), 0.5f), dcy = __fmul_rn(__fadd_rn(db[1], db[3]), 0.5f);
            const float dy = __fsub_rn(dcy, sm.kc[p * 2 + 1]), dx = __fsub_rn(dcx, sm.kc[p * 2]);
            const float dot = dy * sm.vel[p * 2] + dx * sm.vel[p * 2 + 1];
            const float norm = sqrtf(dy * dy + dx * dx) + 1e-6f;
            if (dot < pr.screen_cos * norm) { *lazy_used = 1; return pr.placeholder; }
        }
    }
    float iou = iou_pair(tb, db);
    if (pr.weight_iou) iou = __fmul_rn(iou, sm.in_score[i]);
    float dist = __fsub_rn(1.f, iou);
    if (!sm.valid[p]) return dist;       // the term is +-0 * weight: adding it leaves dist unchanged
    const float kHalfPi = 1.57079637f, kPi = 3.14159274f;
    const float dcx = __fmul_rn(__fadd_rn(db[0], db[2]), 0.5f), dcy = __fmul_rn(__fadd_rn(db[1], db[3]), 0.5f);
    float dy = __fsub_rn(dcy, sm.kc[p * 2 + 1]), dx = __fsub_rn(dcx, sm.kc[p * 2]);
    const float norm = __fadd_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx))), 1e-6f);
    dy = __fdiv_rn(dy, norm);
    dx = __fdiv_rn(dx, norm);
    float cs = __fadd_rn(__fmul_rn(dy, sm.vel[p * 2]), __fmul_rn(dx, sm.vel[p * 2 + 1]));
    cs = cs < -1.f ? -1.f : (cs > 1.f ? 1.f : cs);
    const float ang = st_acos_f32(cs);
    const float na = __fdiv_rn(__fsub_rn(ang, kHalfPi), kPi);
    return __fadd_rn(dist, __fmul_rn(na, pr.vel_w));
}

// One association stage: rows = order positions rows[0..R), columns = filtered columns
// sm.scol[0..C); `ap` = the all-pairs OCM matrix (leading dimension C1), or nullptr for the
// observation-centric recovery, whose cost 1 - IoU(last observation, detection) (:273-308) is
// evaluated right here (its rows and columns are the few left over after stages 1 and 2).
// The sub-matrix is gathered into `dense` together with its admissible-edge statistics, solved,
// and the matches are scattered into det_pos / det_stage / hit.
// (inlined at its three call sites on purpose: out of line, the shared-memory pointers of `sm` are
// re-read through local memory and every sub-phase ran 1.5-2x slower)
__device__ __forceinline__ void associate2(const Params &pr, Smem &sm, int32_t *dbg, int32_t *status_word, const int *rows, int R, int C,
                           const float *ap, int C1, float *dense, int stage_code, bool ocm_lazy,
                           float *dump = nullptr /* diagnostics: receives the matrix of a refused stage-1 problem */) {
    if (R == 0 || C == 0) return;           // dists.size == 0 -> nothing matched (:263-265)
    const double lap_limit = 2.0 * pr.lap_half;
    const bool small = R <= 64 && C <= 64;
    __shared__ int sp_hdr[4];                            // sparse path: nR, nC, E, verdict
    const int warp = threadIdx.x >> 5;
    // debug words 14..18 (stage 1 only): cycles of gather / components / solve / fallback / scatter
    long long t_sub = clock64();
#define ST_SUB(idx)                                                              \
    do {                                                                         \
        if (threadIdx.x == 0 && stage_code == 1) {                               \
            const long long t_now = clock64();                                   \
            dbg[(idx)] = (int)(t_now - t_sub);                                   \
            t_sub = t_now;                                                       \
        }                                                                        \
    } while (0)
    stlap::Problem P;
    P.cost = dense; P.R = R; P.C = C; P.n = R + C; P.half = pr.lap_half;
    P.dbg = nullptr;                                     // no progress marks on the critical path
    P.no_greedy = (g_skip >> 8) & 1;                     // diagnostics bit 8
    P.why = stage_code == 1 ? dbg + 24 : nullptr;        // debug words 24..31: the whole-problem fast path of stage 1
    P.row_cnt = sm.row_cnt; P.col_cnt = sm.col_cnt; P.row_adm = sm.row_adm; P.adm_misc = sm.adm_misc;
    const bool skip = (g_skip >> (stage_code - 1)) & 1;
    bool need_warp_path, sparse_refused = false;
    {
    for (int r = threadIdx.x; r < R; r += kThreads) { sm.row_cnt[r] = 0; sm.row_adm[r] = -1; }
    for (int c = threadIdx.x; c < C; c += kThreads) sm.col_cnt[c] = 0;
    if (threadIdx.x < 64) { sm.masks.rmask[threadIdx.x] = 0ull; sm.masks.cmask[threadIdx.x] = 0ull; }
    if (threadIdx.x == 0) { sm.adm_misc[0] = 0; sm.adm_misc[1] = 0; *sm.masks.fallback = 0; *sm.masks.n_comp = 0; }
    __syncthreads();
    bool any_edge = false;                               // an admissible entry, a NaN or a cost at the limit
    for (int e = threadIdx.x; e < R * C; e += kThreads) {
        const int r = e / C, j = e - r * C;
        float dist;
        if (ap) dist = ap[rows[r] * C1 + sm.scol[j]];
        else {
            const int i = sm.dets[sm.scol[j]];
            float iou = iou_pair(sm.box_ocr + rows[r] * 4, sm.in_box + i * 4);
            if (pr.weight_iou) iou = __fmul_rn(iou, sm.in_score[i]);
            dist = __fsub_rn(1.f, iou);
        }
        dense[e] = dist;
        // admissible-edge statistics for the assignment (lapjv.cuh)
        if ((double)dist < lap_limit) {
            atomicAdd(&sm.row_cnt[r], 1);
            atomicAdd(&sm.col_cnt[j], 1);
            sm.row_adm[r] = j;
            if (small) { atomicOr(&sm.masks.rmask[r], 1ull << j); atomicOr(&sm.masks.cmask[j], 1ull << r); }
            any_edge = true;
        }
        if (dist != dist) { sm.adm_misc[0] = 1; any_edge = true; }
        if (fabs((double)dist - lap_limit) <= 2.0e-9) { sm.adm_misc[1] = 1; any_edge = true; }
    }
    const bool nothing = __syncthreads_or(any_edge) == 0;
    ST_SUB(14);
    if (threadIdx.x == 0) atomicAdd(&dbg[5], 1);        // assignments solved (fire and forget)
    // every cost is above the limit by more than 2e-9 (the usual case in stage 3: what could be matched was matched
    // in stage 1): the reference's solver leaves every row and column unmatched, and so does this stage
    if (nothing && !dump) return;
    need_warp_path = !small;
    if (!skip && R <= 32 && C <= 32) {
        // a handful of rows and columns (always so in stages 2 and 3): the component path on warp 0
        // alone, without block barriers; the other warps go straight to the next barrier
        if (warp == 0) {
            const bool ok = stlap::solve_small_warp(P, sm.lap, sm.masks, sm.sub);
            if (threadIdx.x == 0) sm.adm_misc[2] = ok ? 0 : 1;
        }
        __syncthreads();
        need_warp_path = sm.adm_misc[2] != 0;
    } else if (skip) {
        for (int c = threadIdx.x; c < C; c += kThreads) sm.lap.y[c] = -1;
        need_warp_path = false;
    } else if (small) {
        stlap::prepare_components_block(P, sm.lap, sm.masks);     // ends with a barrier
        ST_SUB(15);
        if (*sm.masks.fallback == 0 && *sm.masks.n_comp == 0 && sm.masks.big[0] == 0ull) {
            need_warp_path = false;                      // isolated pairs and <= 2 x 2 components only: all solved already
        } else if (*sm.masks.fallback == 0) {
            const int n_comp = *sm.masks.n_comp, n_big = sm.masks.n_comp[2];
            bool bad = false;
            // the larger components (crowded groups) each go to the sparse certified solver on a warp of their own, as
            // many at a time as there is scratch for: sm.sub, then pieces of the whole-problem scratch in front of
            // x / y (unused on this path); the other warps enumerate the small components meanwhile
            const size_t piece = (stlap::sub_bytes() + 15) & ~(size_t)15;
            const int n_slots = 1 + min(3, (int)(((size_t)24 * (pr.T + pr.D)) / piece));
            const int n_par = min(n_big, n_slots);
            if (warp < n_par) {
                unsigned char *mem = warp == 0 ? sm.sub
                                               : reinterpret_cast<unsigned char *>(const_cast<double *>(sm.lap.v)) + (warp - 1) * piece;
                for (int k = warp; k < n_big; k += n_par)
                    bad |= !stlap::solve_component_sparse(P, sm.lap, sm.masks, sm.masks.comp_rows[63 - k], sm.masks.comp_cols[63 - k], mem);
            }
            const int first = n_par, nw = kWarps - first;
            if (warp >= first)
                for (int k = warp - first; k < n_comp; k += nw)
                    bad |= !stlap::enumerate_component(P, sm.lap, sm.masks.comp_rows[k], sm.masks.comp_cols[k]);
            need_warp_path = __syncthreads_or(bad) != 0;
        } else {
            need_warp_path = true;
        }
    } else {
        // crowded or wide problems: the sparse certified solver on the admissible edges (lapjv.cuh); its arena is
        // the part of the whole-problem scratch in front of x / y (v, d, cols, pred: 24 bytes per node)
        const int verdict = stlap::sparse_assign_block(P, sm.lap, (size_t)24 * (pr.T + pr.D), sp_hdr,
                                                       stage_code == 1 ? dbg + 24 : nullptr);
        ST_SUB(15);
        need_warp_path = verdict != 0;
        sparse_refused = verdict == 1 || verdict == 6 || verdict == 7;   // the dense certified path would refuse as well
    }
    ST_SUB(16);
    }
    if (need_warp_path) {                                // block-uniform
        // certified Hungarian on one warp; without its certificate (exact ties) the literal JV emulation
        if (warp == 0) {
            const bool ok = !sparse_refused && P.n <= 512 && stlap::fast_assign_warp(P, sm.lap);
            if (threadIdx.x == 0) sm.adm_misc[3] = ok ? 1 : 0;
        }
        __syncthreads();
        if (sm.adm_misc[3] == 0) {
            if (ocm_lazy) {
                // the emulation follows the reference's scan order through EVERY entry: entries that
                // were only bounded from below get their exact value first
                int dummy = 0;
                for (int e = threadIdx.x; e < R * C; e += kThreads) {
                    const int r = e / C, j = e - r * C;
                    dense[e] = ocm_cost(pr, sm, rows[r], sm.dets[sm.scol[j]], false, &dummy);
                }
                __syncthreads();
            }
            if (dump) {                                  // diagnostics bit 9: keep the refused matrix for st_debug_read_cost
                for (int e = threadIdx.x; e < R * C; e += kThreads) dump[e] = dense[e];
                if (threadIdx.x == 0) { dbg[23] = stage_code; dbg[26] = R; dbg[27] = C; }
            }
            if (warp == 0) {
                if (threadIdx.x == 0) atomicAdd(&dbg[6], 1);     // solved by the exact fallback
                const bool ok = stlap::solve_warp(P, sm.lap);
                if (!ok && threadIdx.x == 0) atomicOr(status_word, ST_STATUS_LAP_GUARD);
            }
        }
    }
    __syncthreads();
    ST_SUB(17);
    // col -> row; rows >= R are the solver's dummy rows (= unmatched)
    for (int j = threadIdx.x; j < C; j += kThreads) {
        const int r = sm.lap.y[j];
        if (r >= 0 && r < R) {
            const int i = sm.dets[sm.scol[j]], pos = rows[r];
            sm.det_pos[i] = pos;
            sm.det_stage[i] = (unsigned char)stage_code;
            sm.hit[pos] = 1;
        }
    }
    __syncthreads();
    ST_SUB(18);
#undef ST_SUB
}

__global__ void __launch_bounds__(kThreads)
k4_track_step2(TrackState st, Params pr, int cost_floats, const int32_t *__restrict__ frame_ids,
               const float *__restrict__ det_boxes, const float *__restrict__ det_scores,
               const int64_t *__restrict__ det_labels, const float *__restrict__ det_scales,
               const float *__restrict__ det_depth, const int32_t *__restrict__ det_counts, int det_cap,
               float *cost_ws /* [S, 3, T*D] (two thirds used): written and re-read in this kernel: no __restrict__ */,
               float *__restrict__ out_boxes, float *__restrict__ out_scores, int64_t *__restrict__ out_labels,
               float *__restrict__ out_scales, float *__restrict__ out_depth, int64_t *__restrict__ out_ids,
               int32_t *__restrict__ out_counts, int32_t *__restrict__ out_src_g /* may be null */,
               const float *__restrict__ cmc_warp /* [S,6] camera-motion affine, may be null */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int sh_cnt[8];                            // [6]: pairs queued for the full OCM evaluation
    const int s = blockIdx.x;
    const int T = pr.T, D = pr.D;
    int32_t *dbg = st.debug + s * ST_DEBUG_WORDS;
    // debug words 8..13: SM cycles the last step spent in each phase (thread 0's clock); 20: cycles of
    // the whole step, 21 / 22: low 32 bits of the global nanosecond timer at the CTA's start / end
    long long t_prev = clock64();
    const long long t_begin = t_prev;
    if (threadIdx.x == 0 && ((g_skip >> 7) & 1)) {       // diagnostics bit 7: global-timer stamps of the CTA
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        dbg[21] = (int)(unsigned)ns;
    }
#define ST_PHASE(idx)                                                         \
    do {                                                                      \
        if (threadIdx.x == 0) {                                               \
            const long long t_now = clock64();                                \
            dbg[(idx)] = (int)(t_now - t_prev);                               \
            t_prev = t_now;                                                   \
        }                                                                     \
    } while (0)
    Smem sm = carve2(smem_raw, T, D, cost_floats);
    const long long sbase = (long long)s * T;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int32_t *order = st.order + sbase;
    int32_t *free_slots = st.free_slots + sbase;
    float *gws = cost_ws + (long long)s * 3 * T * D;
    const float *g_boxes = det_boxes + (long long)s * det_cap * 4;
    const float *g_scores = det_scores + (long long)s * det_cap;
    const int64_t *g_labels = det_labels + (long long)s * det_cap;
    const float *g_scales = det_scales + (long long)s * det_cap;
    const float *g_depth = det_depth + (long long)s * det_cap;
    // ---- the frame's working set: three rounds of independent loads ---------------------------
    // round 1: everything whose address does not depend on another load - the per-sequence
    // scalars, and speculatively this thread's `order` entry and detection row (the counts that say
    // whether they are live arrive in the same round; a stale entry is in bounds and is dropped)
    const int frame = frame_ids[s];
    int nd_in = det_counts[s];
    const int n_live_g = st.n_live[s], n_free_g = st.n_free[s];
    const long long id0_g = st.next_id[s];
    const int slot_spec = threadIdx.x < T ? order[threadIdx.x] : 0;
    float4 det_spec = make_float4(0.f, 0.f, 0.f, 0.f);
    float score_spec = 0.f;
    if (threadIdx.x < det_cap && threadIdx.x < D) {
        det_spec = reinterpret_cast<const float4 *>(g_boxes)[threadIdx.x];
        score_spec = g_scores[threadIdx.x];
    }
    if (nd_in > D) {
        if (threadIdx.x == 0) atomicOr(&st.status[s], ST_STATUS_DET_OVERFLOW);
        nd_in = D;
    }
    if (nd_in > det_cap) nd_in = det_cap;

    // frame 0: reset() (:385-387)
    int n_live, n_free;
    long long id0;
    if (frame == 0) {
        for (int i = threadIdx.x; i < T; i += kThreads) free_slots[i] = T - 1 - i;
        n_live = 0; n_free = T; id0 = 0;
    } else {
        n_live = n_live_g; n_free = n_free_g; id0 = id0_g;
    }
    for (int i = threadIdx.x; i < nd_in; i += kThreads) {
        sm.det_pos[i] = -1;
        sm.det_stage[i] = 0;
        const float4 b = i < kThreads ? det_spec : reinterpret_cast<const float4 *>(g_boxes)[i];
        sm.in_box[i * 4 + 0] = b.x; sm.in_box[i * 4 + 1] = b.y; sm.in_box[i * 4 + 2] = b.z; sm.in_box[i * 4 + 3] = b.w;
        sm.in_score[i] = i < kThreads ? score_spec : g_scores[i];
    }
    // rounds 2 and 3: the per-track scalars (addresses from the slot), then the two ring entries
    // (addresses from the observation count)
    for (int p = threadIdx.x; p < n_live; p += kThreads) {
        const int slot = p < kThreads ? slot_spec : order[p];
        const long long g = sbase + slot;
        const int n = st.num_obs[g];
        const int tent = st.tentative[g];
        const float v0 = st.velocity[g * 2], v1 = st.velocity[g * 2 + 1];
        const float4 lo = reinterpret_cast<const float4 *>(st.last_obs)[g];
        // k_step_observation (:173-185) as the association sees it, and as the velocity update of a
        // track matched in this frame will see it (one more observation appended)
        const int r1 = ((n - 1 - pr.vel_dt) % ST_RING + ST_RING) % ST_RING;
        const int r2 = ((n - pr.vel_dt) % ST_RING + ST_RING) % ST_RING;
        const float4 o1 = reinterpret_cast<const float4 *>(st.obs_ring)[g * ST_RING + r1];
        const float4 o2 = reinterpret_cast<const float4 *>(st.obs_ring)[g * ST_RING + r2];
        const bool ok1 = n > pr.vel_dt && st.obs_valid[g * ST_RING + r1];
        const bool ok2 = pr.vel_dt >= 1 && n >= pr.vel_dt && st.obs_valid[g * ST_RING + r2];
        sm.slot[p] = slot;
        sm.id[p] = st.id[g];
        sm.tent[p] = (unsigned char)tent;
        sm.tracked0[p] = st.tracked[g];
        sm.has_saved[p] = st.has_saved[g];
        sm.last_frame[p] = st.last_frame[g];
        sm.num_obs[p] = n;
        sm.miss_run[p] = st.miss_run[g];
        sm.hits[p] = st.hits[g];
        sm.hit[p] = 0;
        sm.vel[p * 2] = v0; sm.vel[p * 2 + 1] = v1;
        sm.box_ocr[p * 4 + 0] = lo.x; sm.box_ocr[p * 4 + 1] = lo.y; sm.box_ocr[p * 4 + 2] = lo.z; sm.box_ocr[p * 4 + 3] = lo.w;
        const float4 ko = ok1 ? o1 : lo;
        sm.kc[p * 2] = __fmul_rn(__fadd_rn(ko.x, ko.z), 0.5f);
        sm.kc[p * 2 + 1] = __fmul_rn(__fadd_rn(ko.y, ko.w), 0.5f);
        const float kob[4] = {ko.x, ko.y, ko.z, ko.w};
        sm.valid[p] = (__fadd_rn(v0, v1) != -2.f) && (sum4(kob) != -4.f);
        sm.knext_ok[p] = ok2;
        sm.knext[p * 4 + 0] = o2.x; sm.knext[p * 4 + 1] = o2.y; sm.knext[p * 4 + 2] = o2.z; sm.knext[p * 4 + 3] = o2.w;
        if (tent)      // tentative tracks are matched with their un-predicted mean (:431-441 predicts confirmed ones only)
            mean_to_box((float)st.mean[g * 8 + 0], (float)st.mean[g * 8 + 1], (float)st.mean[g * 8 + 2],
                        (float)st.mean[g * 8 + 3], sm.box_ocm + p * 4);
    }
    __syncthreads();
    int n_out = 0, first_new = 0;

    if (n_live == 0 || nd_in == 0) {
        // ---- init-only branch (:391-403): keep scores > init_track_thr, all new ids
        if (warp == 0) compact_warp(nd_in, [&](int i) { return sm.in_score[i] > pr.init_track_thr; }, sm.out_src, &sh_cnt[0]);
        __syncthreads();
        n_out = sh_cnt[0];
        first_new = 0;
        for (int i = threadIdx.x; i < n_out; i += kThreads) sm.det_stage[sm.out_src[i]] = 4;
        __syncthreads();
    } else {
        // ---- detection filter (:412-421), confirmed and tentative tracks: three warps side by side
        if (warp == 0)
            compact_warp(nd_in, [&](int i) {
                const float *b = sm.in_box + i * 4;
                const float area = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
                return sm.in_score[i] > pr.obj_score_thr && area > 100.f;
            }, sm.dets, &sh_cnt[0]);
        else if (warp == 1) compact_warp(n_live, [&](int p) { return !sm.tent[p]; }, sm.conf, &sh_cnt[1]);
        else if (warp == 2) compact_warp(n_live, [&](int p) { return sm.tent[p] != 0; }, sm.tentl, &sh_cnt[2]);
        __syncthreads();
        const int C1 = sh_cnt[0], Rc = sh_cnt[1], Rt = sh_cnt[2];
        for (int c = threadIdx.x; c < C1; c += kThreads) { sm.det_stage[sm.dets[c]] = 4; sm.scol[c] = c; }   // leftover unless matched
        // ---- confirmed tracks: predict (:431-441), four tracks per warp (kalman.cuh)
        for (int r0 = warp * 4; r0 < Rc; r0 += kWarps * 4) {
            const int r = r0 + (lane >> 3), gl = lane & 7;
            const bool active = r < Rc;
            const int pos = active ? sm.conf[r] : 0;
            const long long g = active ? sbase + sm.slot[pos] : 0;
            stkf::Row8 row = stkf::load8(st.mean + g * 8, st.cov + g * 64, gl, active);
            if (active) {
                if (sm.last_frame[pos] != frame - 1 && gl == 7) row.m = 0.0;        // lost last frame: vh = 0
                if (sm.tracked0[pos]) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) st.saved_cov[g * 64 + gl * 8 + j] = row.p[j];
                    st.saved_mean[g * 8 + gl] = row.m;
                    if (gl == 0) { st.has_saved[g] = 1; sm.has_saved[pos] = 1; }
                }
            }
            if (!(g_skip & 32)) stkf::predict8(row, gl);
            stkf::store8(row, st.mean + g * 8, st.cov + g * 64, gl, active);
            const float cx = (float)stkf::shfl8(row.m, 0), cy = (float)stkf::shfl8(row.m, 1);
            const float a = (float)stkf::shfl8(row.m, 2), h = (float)stkf::shfl8(row.m, 3);
            if (active && gl == 0) mean_to_box(cx, cy, a, h, sm.box_ocm + pos * 4);
        }
        __syncthreads();
        // ---- camera-motion compensation of the confirmed tracks (:443-445, gmc.py:20-45):
        // centre and its velocity through R (+ t), h and vh by s = sqrt(max(det R, 1e-12)),
        // covariance A P A^T with A = diag-blocks (R, 1, s, R, 1, s).  A NaN first entry
        // means "no warp this frame" (the reference's estimator returned None).
        if (cmc_warp != nullptr && cmc_warp[s * 6] == cmc_warp[s * 6]) {
            const float *wm = cmc_warp + s * 6;
            const double r00 = (double)wm[0], r01 = (double)wm[1], t0 = (double)wm[2];
            const double r10 = (double)wm[3], r11 = (double)wm[4], t1 = (double)wm[5];
            const double sc = sqrt(fmax(__dsub_rn(__dmul_rn(r00, r11), __dmul_rn(r01, r10)), 1e-12));
            double *kfws = reinterpret_cast<double *>(sm.cost) + warp * stkf::kScratchDoubles;   // the cost region is still free
            for (int r = warp; r < Rc; r += kWarps) {
                const int pos = sm.conf[r];
                const long long g = sbase + sm.slot[pos];
                double *P = kfws, *X = kfws + 64, *m = kfws + 128;
                for (int e = lane; e < 64; e += 32) P[e] = st.cov[g * 64 + e];
                if (lane < 8) m[lane] = st.mean[g * 8 + lane];
                __syncwarp();
                // row map: 0,1 and 4,5 mix through R; 3 and 7 scale; 2 and 6 stay
                auto mix = [&](const double *src, int i, int j, bool rows) -> double {
                    const int k = rows ? i : j;              // index being transformed
                    auto at = [&](int q) { return rows ? src[q * 8 + j] : src[i * 8 + q]; };
                    const int base = k & 4;                  // 0 or 4
                    switch (k & 3) {
                        case 0: return __dadd_rn(__dmul_rn(r00, at(base)), __dmul_rn(r01, at(base + 1)));
                        case 1: return __dadd_rn(__dmul_rn(r10, at(base)), __dmul_rn(r11, at(base + 1)));
                        case 3: return __dmul_rn(sc, at(k));
                        default: return at(k);
                    }
                };
                for (int e = lane; e < 64; e += 32) X[e] = mix(P, e >> 3, e & 7, true);      // X = A P
                __syncwarp();
                for (int e = lane; e < 64; e += 32) st.cov[g * 64 + e] = mix(X, e >> 3, e & 7, false);   // X A^T
                if (lane == 0) {
                    const double n0 = __dadd_rn(__dadd_rn(__dmul_rn(r00, m[0]), __dmul_rn(r01, m[1])), t0);
                    const double n1 = __dadd_rn(__dadd_rn(__dmul_rn(r10, m[0]), __dmul_rn(r11, m[1])), t1);
                    const double n3 = __dmul_rn(m[3], sc);
                    st.mean[g * 8 + 0] = n0;
                    st.mean[g * 8 + 1] = n1;
                    st.mean[g * 8 + 3] = n3;
                    st.mean[g * 8 + 4] = __dadd_rn(__dmul_rn(r00, m[4]), __dmul_rn(r01, m[5]));
                    st.mean[g * 8 + 5] = __dadd_rn(__dmul_rn(r10, m[4]), __dmul_rn(r11, m[5]));
                    st.mean[g * 8 + 7] = __dmul_rn(m[7], sc);
                    mean_to_box((float)n0, (float)n1, (float)m[2], (float)n3, sm.box_ocm + pos * 4);
                }
                __syncwarp();
            }
            __syncthreads();
        }
        ST_PHASE(8);
        // ---- all-pairs OCM costs (:187-258): every live track against every filtered detection, float32
        // in the reference's evaluation order (stage 1 takes the confirmed rows, stage 2 the tentative ones)
        const int A = n_live * C1;                       // <= T * D <= 2^20
        float *ap_ocm, *dense_sm;
        int dense_cap;
        const int cost_avail = sm.cost_cap;
        if (A <= cost_avail) {
            ap_ocm = sm.cost; dense_sm = sm.cost + A; dense_cap = cost_avail - A;
        } else {
            ap_ocm = gws; dense_sm = sm.cost; dense_cap = cost_avail;
        }
        float *dense_g = gws + (long long)T * D;
        const bool lazy = !((g_skip >> 6) & 1) && pr.screen_cos > -1.5f;   // diagnostics bit 6: evaluate every entry exactly
        int lazy_used = 0;
        // Two passes: the screen settles the pairs that do not overlap (nearly all of them) and queues the rest;
        // the full evaluation (five IEEE divisions, two square roots, a float64 acos) then runs with every lane
        // busy - evaluated in place, one overlapping pair per warp held the other 31 lanes for ~1 200 cycles.
        // The queue borrows the room of the stage's dense matrix; what does not fit is evaluated in place.
        int *queue = reinterpret_cast<int *>(dense_sm);
        const int queue_cap = lazy ? dense_cap : 0;
        if (threadIdx.x == 0) sh_cnt[6] = 0;
        __syncthreads();
        for (int e0 = warp * 32; e0 < A; e0 += kThreads) {        // 32-bit index math: a 64-bit division per pair cost more than the pair
            const int e = e0 + (threadIdx.x & 31);
            bool need = false;
            int p = 0, c = 0;
            if (e < A) {
                p = e / C1; c = e - p * C1;
                float val = 0.f;
                need = !lazy || ocm_screen(pr, sm, p, sm.dets[c], &val, &lazy_used);
                if (!need) ap_ocm[e] = val;
            }
            const unsigned mb = __ballot_sync(0xffffffffu, need);
            if (mb == 0u) continue;
            int base = 0;
            if ((threadIdx.x & 31) == 0) base = atomicAdd(&sh_cnt[6], __popc(mb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (need) {
                const int pos = base + __popc(mb & ((1u << (threadIdx.x & 31)) - 1u));
                if (pos < queue_cap) queue[pos] = e;
                else ap_ocm[e] = ocm_cost(pr, sm, p, sm.dets[c], false, &lazy_used);
            }
        }
        __syncthreads();
        {
            const int n_queued = min(sh_cnt[6], queue_cap);
            int unused = 0;
            for (int k = threadIdx.x; k < n_queued; k += kThreads) {
                const int e = queue[k], p = e / C1, c = e - p * C1;
                ap_ocm[e] = ocm_cost(pr, sm, p, sm.dets[c], false, &unused);
            }
        }
        const bool ocm_lazy = __syncthreads_or(lazy_used) != 0;
        if (threadIdx.x == 0) dbg[19] = (int)(clock64() - t_prev);      // cycles of the all-pairs pass (t_prev = end of phase 8)
        // ---- stage 1: confirmed tracks vs filtered dets (:448-474)
        associate2(pr, sm, dbg, &st.status[s], sm.conf, Rc, C1, ap_ocm, C1,
                   Rc * C1 <= dense_cap ? dense_sm : dense_g, 1, ocm_lazy,
                   ((g_skip >> 9) & 1) ? gws + 2ll * T * D : nullptr);
        ST_PHASE(9);
        // ---- stage 2: tentative tracks vs leftovers (:476-503)
        if (warp == 0) compact_warp(C1, [&](int c) { return sm.det_stage[sm.dets[c]] == 4; }, sm.scol, &sh_cnt[3]);
        __syncthreads();
        const int C2 = sh_cnt[3];
        associate2(pr, sm, dbg, &st.status[s], sm.tentl, Rt, C2, ap_ocm, C1,
                   Rt * C2 <= dense_cap ? dense_sm : dense_g, 2, ocm_lazy, ((g_skip >> 9) & 1) ? gws + 2ll * T * D : nullptr);
        ST_PHASE(10);
        // ---- stage 3: OCR, every unmatched track's last observation (:505-569)
        if (warp == 0) compact_warp(C1, [&](int c) { return sm.det_stage[sm.dets[c]] == 4; }, sm.scol, &sh_cnt[4]);
        else if (warp == 1) compact_warp(n_live, [&](int p) { return !sm.hit[p]; }, sm.rows, &sh_cnt[5]);
        __syncthreads();
        const int C3 = sh_cnt[4], R3 = sh_cnt[5];
        associate2(pr, sm, dbg, &st.status[s], sm.rows, R3, C3, nullptr, C1,
                   R3 * C3 <= dense_cap ? dense_sm : dense_g, 3, false, ((g_skip >> 9) & 1) ? gws + 2ll * T * D : nullptr);
        ST_PHASE(11);
        // ---- unmatched tracks: tracked = False, obs.append(None) (:579-582): plain stores
        for (int p = threadIdx.x; p < n_live; p += kThreads)
            if (!sm.hit[p]) {
                const long long g = sbase + sm.slot[p];
                const int n = sm.num_obs[p];
                st.tracked[g] = 0;
                st.obs_valid[g * ST_RING + n % ST_RING] = 0;
                st.miss_run[g] = sm.miss_run[p] + 1;
                st.num_obs[g] = n + 1;
            }
        // ---- output order: stage 1, 2, 3 matches then leftovers (:584-596)
        if (warp == 0) {
            int total = 0, matched = 0;
            for (int code = 1; code <= 4; ++code) {
                if (code == 4) matched = total;
                for (int i0 = 0; i0 < nd_in; i0 += 32) {
                    const int i = i0 + lane;
                    const bool f = i < nd_in && sm.det_stage[i] == code;
                    const unsigned b = __ballot_sync(0xffffffffu, f);
                    if (f) sm.out_src[total + __popc(b & ((1u << lane) - 1u))] = i;
                    total += __popc(b);
                }
            }
            if (lane == 0) { sh_cnt[0] = total; sh_cnt[6] = matched; }
        }
        __syncthreads();
        n_out = sh_cnt[0];
        first_new = sh_cnt[6];      // new tracks are exactly the outputs without a track; they are last in order
    }

    ST_PHASE(12);
    // ---- ids + BaseTracker.update (:598-605, base_tracker.py:54-91) ------------
    const int n_new = n_out - first_new;
    int n_create = n_new;
    if (n_live + n_new > T) {
        n_create = T - n_live;
        if (threadIdx.x == 0) atomicOr(&st.status[s], ST_STATUS_TRACK_OVERFLOW);
    }
    // four output rows per warp: 8-lane groups run the online smoothing (:571-577, :319-343) and the
    // Kalman update of their row, or the birth of a new track
    for (int k0 = warp * 4; k0 < n_out; k0 += kWarps * 4) {
        const int k = k0 + (lane >> 3), gl = lane & 7;
        const bool active = k < n_out;
        const int i = active ? sm.out_src[k] : 0;
        const float *box = sm.in_box + i * 4;
        float zc[4];
        stkf::xyxy_to_cxcyah(box, zc);
        const int pos = active ? sm.det_pos[i] : -1;
        const bool upd = active && pos >= 0;
        int slot = upd ? sm.slot[pos] : -1;
        long long id = 0;
        {
            // update_track (:128-146 + kalman_tracker_base.py:65-75); every lane runs the shuffles
            const long long g = upd ? sbase + slot : 0;
            const bool smooth = upd && !sm.tracked0[pos] && !(g_skip & 8);        // the track was lost before this step
            const int miss = smooth ? sm.miss_run[pos] : 0;
            const bool from_saved = smooth && sm.has_saved[pos];
            stkf::Row8 row = stkf::load8(from_saved ? st.saved_mean + g * 8 : st.mean + g * 8,
                                         from_saved ? st.saved_cov + g * 64 : st.cov + g * 64, gl, upd);
            const int max_miss = __reduce_max_sync(0xffffffffu, miss);
            if (max_miss > 0) {
                // virtual observations on the line from the last match to the new one, no predict in between
                float last[4], shift[4];
                const float denom = (float)(miss + 1);
                for (int q = 0; q < 4; ++q) {
                    last[q] = smooth ? sm.box_ocr[pos * 4 + q] : 0.f;
                    shift[q] = __fdiv_rn(__fsub_rn(box[q], last[q]), denom);
                }
                for (int t = 0; t < max_miss; ++t) {
                    float vb[4], vz[4];
                    double z[4];
                    for (int q = 0; q < 4; ++q) vb[q] = __fadd_rn(last[q], __fmul_rn((float)(t + 1), shift[q]));
                    stkf::xyxy_to_cxcyah(vb, vz);
                    for (int q = 0; q < 4; ++q) z[q] = (double)vz[q];
                    stkf::Row8 tmp = row;
                    stkf::update8(tmp, z, gl);
                    if (t < miss) row = tmp;
                }
            }
            double z[4];
            for (int q = 0; q < 4; ++q) z[q] = (double)zc[q];
            if (!(g_skip & 16)) stkf::update8(row, z, gl);
            stkf::store8(row, st.mean + g * 8, st.cov + g * 64, gl, upd);
            if (upd) id = sm.id[pos];
            if (upd && gl == 0) {
                const int h = sm.hits[pos] + 1, n = sm.num_obs[pos];
                st.hits[g] = h;
                st.last_frame[g] = frame;
                if (sm.tent[pos] && h >= pr.num_tentatives) st.tentative[g] = 0;
                st.tracked[g] = 1;
                // obs.append(box)
                const int r = n % ST_RING;
                for (int q = 0; q < 4; ++q) { st.obs_ring[(g * ST_RING + r) * 4 + q] = box[q]; st.last_obs[g * 4 + q] = box[q]; }
                st.obs_valid[g * ST_RING + r] = 1;
                st.miss_run[g] = 0;
                st.num_obs[g] = n + 1;
                // velocity against the k-step observation of the extended list (:140-146)
                const float *ko = sm.knext_ok[pos] ? sm.knext + pos * 4 : box;
                vel_direction(ko, box, st.velocity + g * 2);
            }
        }
        if (active && !upd) {
            const int r = k - first_new;
            id = id0 + r;
            if (r < n_create) {
                // init_track (:105-126 + kalman_tracker_base.py:55-63)
                slot = free_slots[n_free - 1 - r];
                const long long g = sbase + slot;
                stkf::initiate8(st.mean + g * 8, st.cov + g * 64, zc, gl);
                if (gl == 0) {
                    st.id[g] = id;
                    st.hits[g] = 1;
                    st.last_frame[g] = frame;
                    st.tentative[g] = frame != 0;
                    st.tracked[g] = 1;
                    st.has_saved[g] = 0;
                    st.num_obs[g] = 1;
                    st.miss_run[g] = 0;
                    for (int q = 0; q < 4; ++q) { st.obs_ring[(g * ST_RING) * 4 + q] = box[q]; st.last_obs[g * 4 + q] = box[q]; }
                    st.obs_valid[g * ST_RING] = 1;
                    st.velocity[g * 2] = -1.f;
                    st.velocity[g * 2 + 1] = -1.f;
                    sm.slot[n_live + r] = slot;
                }
            }
        }
        if (active && gl == 0) {
            const long long o = (long long)s * det_cap + k;
            const float score = sm.in_score[i], scale = g_scales[i], depth = g_depth[i];
            const int64_t label = g_labels[i];
            if (slot >= 0) {
                const long long g = sbase + slot;
                for (int q = 0; q < 4; ++q) st.last_box[g * 4 + q] = box[q];
                st.last_score[g] = score;
                st.last_scale[g] = scale;
                st.last_depth[g] = depth;
                st.last_label[g] = label;
            }
            for (int q = 0; q < 4; ++q) out_boxes[o * 4 + q] = box[q];
            out_scores[o] = score;
            out_labels[o] = label;
            out_scales[o] = scale;
            out_depth[o] = depth;
            out_ids[o] = id;
            if (out_src_g) out_src_g[o] = i;
        }
    }
    __syncthreads();
    ST_PHASE(13);
    // ---- pop_invalid_tracks (kalman_tracker_base.py:77-88): from the shared-memory copies ----
    const int n_now = n_live + n_create;
    if (threadIdx.x < 32) {
        int kept = 0, freed = 0;
        for (int p0 = 0; p0 < n_now; p0 += 32) {
            const int p = p0 + lane;
            int slot = -1;
            bool dead = false;
            if (p < n_now) {
                slot = sm.slot[p];
                int lf = frame;
                bool tent_now = frame != 0;                       // a track born in this frame
                if (p < n_live) {
                    const bool hit = sm.hit[p] != 0;
                    lf = hit ? frame : sm.last_frame[p];
                    tent_now = sm.tent[p] && !(hit && sm.hits[p] + 1 >= pr.num_tentatives);
                }
                dead = (frame - lf >= pr.retain) || (tent_now && lf != frame);
            }
            const unsigned alive_b = __ballot_sync(0xffffffffu, p < n_now && !dead);
            const unsigned dead_b = __ballot_sync(0xffffffffu, p < n_now && dead);
            if (p < n_now && !dead) order[kept + __popc(alive_b & ((1u << lane) - 1u))] = slot;
            if (p < n_now && dead)
                free_slots[(n_free - n_create) + freed + __popc(dead_b & ((1u << lane) - 1u))] = slot;
            kept += __popc(alive_b);
            freed += __popc(dead_b);
        }
        if (lane == 0) {
            st.n_live[s] = kept;
            st.n_free[s] = n_free - n_create + freed;
            st.next_id[s] = id0 + n_new;
            out_counts[s] = n_out;
            dbg[20] = (int)(clock64() - t_begin);
            if ((g_skip >> 7) & 1) {
                unsigned long long ns;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
                dbg[22] = (int)(unsigned)ns;
            }
        }
    }
#undef ST_PHASE
}

}  // namespace

int st_k4v2_set_skip(int mask) {
    return cudaMemcpyToSymbol(g_skip, &mask, sizeof(int)) == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

// shared-memory opt-in of the round-2 kernel for these capacities (st_create); ST_ERR_CAPACITY when
// its fixed part does not fit (the round-1 kernel is then the only one available)
int st_k4v2_configure(int max_tracks, int max_dets) {
    const int cf = cost_floats_for(max_tracks, max_dets);
    if (cf <= 0) return ST_ERR_CAPACITY;
    const size_t smem = smem2_fixed_bytes(max_tracks, max_dets) + sizeof(float) * (size_t)cf;
    return cudaFuncSetAttribute(k4_track_step2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess
               ? ST_OK : ST_ERR_CAPACITY;
}

int st_k4v2_launch(st_handle *h, const stk4::Params &pr, int S, const int32_t *frame_ids, const float *det_boxes,
                   const float *det_scores, const int64_t *det_labels, const float *det_scales, const float *det_depth,
                   const int32_t *det_counts, int det_cap, float *out_boxes, float *out_scores, int64_t *out_labels,
                   float *out_scales, float *out_depth, int64_t *out_ids, int32_t *out_counts, int32_t *out_src,
                   const float *warp, cudaStream_t stream) {
    const int cf = cost_floats_for(pr.T, pr.D);
    const size_t smem = smem2_fixed_bytes(pr.T, pr.D) + sizeof(float) * (size_t)cf;
    k4_track_step2<<<S, kThreads, smem, stream>>>(h->st, pr, cf, frame_ids, det_boxes, det_scores, det_labels, det_scales,
                                                 det_depth, det_counts, det_cap, h->cost, out_boxes, out_scores,
                                                 out_labels, out_scales, out_depth, out_ids, out_counts, out_src, warp);
    return st_last_launch();
}
