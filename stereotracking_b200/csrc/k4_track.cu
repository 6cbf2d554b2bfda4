// K4 + K5: one frame of the OC-SORT tracker with depth-scaled boxes for S
// sequences in lockstep, one CTA per sequence, the whole frame in one launch:
//   detection filter -> Kalman predict of confirmed tracks (8-lane groups, four tracks per
//   warp, state in registers) -> optional camera-motion compensation
//   -> OCM cost (1 - IoU + velocity-direction term) -> LAPJV  (stage 1)
//   -> tentative tracks vs leftovers                           (stage 2)
//   -> observation-centric recovery on last observations       (stage 3)
//   -> online smoothing, Kalman update / track birth, eviction, id allocation.
// Replaces OCSORTTracker_Disparity.track and everything it calls:
//   mmtrack/models/trackers/ocsort_tracker_disparity.py:105-185,:187-343,:345-618
//   mmtrack/models/trackers/kalman_tracker_base.py:50-88
//   mmtrack/models/trackers/base_tracker.py:34-121
//   mmtrack/models/motion/kalman_filter.py:60-189
// Track state lives in the handle as a structure of arrays; `order` keeps the
// live slots in ascending-id order, which is the reference's dict order.
#include <new>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "st_common.cuh"
#include "kalman.cuh"
#include "lapjv.cuh"
#include "k4_common.cuh"

using namespace stk4;

namespace {

#ifndef ST_K4_THREADS
#define ST_K4_THREADS 512
#endif
constexpr int kThreads = ST_K4_THREADS;
// cost matrices up to this many entries stay in shared memory (56 KB: two CTAs per SM fit)
constexpr int kCostSmemFloats = 14336;
constexpr int kWarps = kThreads / 32;


struct Smem {                    // carved from dynamic shared memory
    stlap::Scratch lap;
    double *kf;                  // [kWarps][kScratchDoubles]
    float *row_box;              // [T,4] track boxes of the current stage (xyxy)
    float *row_vel;              // [T,2]
    float *row_kc;               // [T,2] centre of the k-step observation (cx, cy)
    float *det_box;              // [D,4] boxes of the dets in the current column list
    int *rows;                   // [T] slot of each row of the current stage
    int *dets;                   // [D] input index of each column of the current stage
    int *det_slot;               // [D] per INPUT det: matched slot or -1
    int *out_src;                // [D] output order -> input det index
    unsigned char *row_valid;    // [T]
    unsigned char *det_stage;    // [D] per INPUT det: 0 dropped, 1..3 matched in stage, 4 leftover
    unsigned char *slot_hit;     // [T] per ORDER position: matched this frame
    float *cost;                 // [cost_cap] cost matrix of the current stage when it fits
    int cost_cap;
    int *row_cnt, *col_cnt, *row_adm, *adm_misc;   // admissible-edge statistics (stlap::Problem)
    // the frame's working set, fetched once (every later phase is a chain of dependent reads)
    int *ord_slot;               // [T] order[p]
    int *slot_pos;               // [T] slot -> order position (live slots only)
    unsigned char *ord_tent;     // [T] tentative flag per order position
    unsigned char *ord_tracked;  // [T] tracked flag per order position (as of the start of the frame)
    float *in_box;               // [D,4] detection boxes
    float *in_score;             // [D]
};

__host__ __device__ inline size_t smem_bytes(int T, int D) {
    size_t b = stlap::scratch_bytes(T + D);
    b += align16(sizeof(double) * kWarps * stkf::kScratchDoubles);
    b += align16(sizeof(float) * T * 4) + align16(sizeof(float) * T * 2) * 2 + align16(sizeof(float) * D * 4);
    b += align16(sizeof(int) * T) + align16(sizeof(int) * D) * 3;
    b += align16(T) + align16(D) + align16(T);
    b += align16(sizeof(float) * (size_t)((long long)T * D < kCostSmemFloats ? T * D : kCostSmemFloats));
    b += align16(sizeof(int) * T) * 2 + align16(sizeof(int) * D) + 16;
    b += align16(sizeof(int) * T) * 2 + align16(T) * 2 + align16(sizeof(float) * D * 4) + align16(sizeof(float) * D);
    return b;
}

__device__ inline Smem carve(unsigned char *p, int T, int D) {
    Smem m;
    m.lap = stlap::carve(p, T + D); p += stlap::scratch_bytes(T + D);
    m.kf = reinterpret_cast<double *>(p); p += align16(sizeof(double) * kWarps * stkf::kScratchDoubles);
    m.row_box = reinterpret_cast<float *>(p); p += align16(sizeof(float) * T * 4);
    m.row_vel = reinterpret_cast<float *>(p); p += align16(sizeof(float) * T * 2);
    m.row_kc = reinterpret_cast<float *>(p); p += align16(sizeof(float) * T * 2);
    m.det_box = reinterpret_cast<float *>(p); p += align16(sizeof(float) * D * 4);
    m.rows = reinterpret_cast<int *>(p); p += align16(sizeof(int) * T);
    m.dets = reinterpret_cast<int *>(p); p += align16(sizeof(int) * D);
    m.det_slot = reinterpret_cast<int *>(p); p += align16(sizeof(int) * D);
    m.out_src = reinterpret_cast<int *>(p); p += align16(sizeof(int) * D);
    m.row_valid = p; p += align16(T);
    m.det_stage = p; p += align16(D);
    m.slot_hit = p; p += align16(T);
    m.cost_cap = (long long)T * D < kCostSmemFloats ? T * D : kCostSmemFloats;
    m.cost = reinterpret_cast<float *>(p); p += align16(sizeof(float) * m.cost_cap);
    m.row_cnt = reinterpret_cast<int *>(p); p += align16(sizeof(int) * T);
    m.row_adm = reinterpret_cast<int *>(p); p += align16(sizeof(int) * T);
    m.col_cnt = reinterpret_cast<int *>(p); p += align16(sizeof(int) * D);
    m.adm_misc = reinterpret_cast<int *>(p); p += 16;
    m.ord_slot = reinterpret_cast<int *>(p); p += align16(sizeof(int) * T);
    m.slot_pos = reinterpret_cast<int *>(p); p += align16(sizeof(int) * T);
    m.ord_tent = p; p += align16(T);
    m.ord_tracked = p; p += align16(T);
    m.in_box = reinterpret_cast<float *>(p); p += align16(sizeof(float) * D * 4);
    m.in_score = reinterpret_cast<float *>(p);
    return m;
}

// ordered (stable) compaction of the indices i in [0,n) with flag(i) into dst;
// executed by warp 0, returns the count to every thread of the block via *cnt_sh
template <typename F>
__device__ inline void compact_warp0(int n, F flag, int *dst, int *cnt_sh) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
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
    __syncthreads();
}

// one association stage: rows = slots in sm.rows[0..R), cols = input dets in
// sm.dets[0..C).  ocm = velocity-direction term on (stage 1/2); boxes of the
// rows come from the Kalman mean (ocm) or the last observation (ocr).
__device__ void associate(const TrackState &st, const Params &pr, Smem &sm, long long sbase, int R, int C,
                          bool ocm, const DetView &dv, float *cost_global, int stage_code, int32_t *status_word) {
    if (R == 0 || C == 0) return;           // dists.size == 0 -> nothing matched (:263-265)
    // the matrix lives in shared memory when it fits (every solver step reads a row of it)
    float *cost = R * C <= sm.cost_cap ? sm.cost : cost_global;
    const double lap_limit = 2.0 * pr.lap_half;
    for (int r = threadIdx.x; r < R; r += kThreads) { sm.row_cnt[r] = 0; sm.row_adm[r] = -1; }
    for (int c = threadIdx.x; c < C; c += kThreads) sm.col_cnt[c] = 0;
    if (threadIdx.x == 0) { sm.adm_misc[0] = 0; sm.adm_misc[1] = 0; }
    // ---- per-row data --------------------------------------------------------
    for (int r = threadIdx.x; r < R; r += kThreads) {
        const long long g = sbase + sm.rows[r];
        if (ocm) {
            // track_bboxes = cxcyah_to_xyxy(float32(mean[:4])) (:217-223)
            const float cx = (float)st.mean[g * 8 + 0], cy = (float)st.mean[g * 8 + 1];
            const float a = (float)st.mean[g * 8 + 2], h = (float)st.mean[g * 8 + 3];
            const float hw = __fmul_rn(__fmul_rn(a, h), 0.5f), hh = __fmul_rn(h, 0.5f);
            sm.row_box[r * 4 + 0] = __fsub_rn(cx, hw);
            sm.row_box[r * 4 + 1] = __fsub_rn(cy, hh);
            sm.row_box[r * 4 + 2] = __fadd_rn(cx, hw);
            sm.row_box[r * 4 + 3] = __fadd_rn(cy, hh);
            const float v0 = st.velocity[g * 2], v1 = st.velocity[g * 2 + 1];
            sm.row_vel[r * 2] = v0;
            sm.row_vel[r * 2 + 1] = v1;
            const float *ko = k_step_obs(st, g, pr.vel_dt);
            sm.row_kc[r * 2] = __fmul_rn(__fadd_rn(ko[0], ko[2]), 0.5f);
            sm.row_kc[r * 2 + 1] = __fmul_rn(__fadd_rn(ko[1], ko[3]), 0.5f);
            sm.row_valid[r] = (__fadd_rn(v0, v1) != -2.f) && (sum4(ko) != -4.f);
        } else {
            for (int k = 0; k < 4; ++k) sm.row_box[r * 4 + k] = st.last_obs[g * 4 + k];
        }
    }
    for (int c = threadIdx.x; c < C; c += kThreads)
        for (int k = 0; k < 4; ++k) sm.det_box[c * 4 + k] = sm.in_box[sm.dets[c] * 4 + k];
    __syncthreads();
    // ---- cost matrix -----------------------------------------------------------
    const float kHalfPi = 1.57079637f, kPi = 3.14159274f;
    for (int e = threadIdx.x; e < R * C; e += kThreads) {
        const int r = e / C, c = e % C;
        const float *tb = sm.row_box + r * 4, *db = sm.det_box + c * 4;
        float iou = iou_pair(tb, db);
        if (pr.weight_iou) iou = __fmul_rn(iou, sm.in_score[sm.dets[c]]);
        float dist = __fsub_rn(1.f, iou);
        if (ocm) {
            const float dcx = __fmul_rn(__fadd_rn(db[0], db[2]), 0.5f), dcy = __fmul_rn(__fadd_rn(db[1], db[3]), 0.5f);
            float dy = __fsub_rn(dcy, sm.row_kc[r * 2 + 1]), dx = __fsub_rn(dcx, sm.row_kc[r * 2]);
            const float norm = __fadd_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(dy, dy), __fmul_rn(dx, dx))), 1e-6f);
            dy = __fdiv_rn(dy, norm);
            dx = __fdiv_rn(dx, norm);
            float cs = __fadd_rn(__fmul_rn(dy, sm.row_vel[r * 2]), __fmul_rn(dx, sm.row_vel[r * 2 + 1]));
            cs = cs < -1.f ? -1.f : (cs > 1.f ? 1.f : cs);
            const float ang = st_acos_f32(cs);
            const float na = __fdiv_rn(__fsub_rn(ang, kHalfPi), kPi);
            const float term = __fmul_rn(na, sm.row_valid[r] ? 1.f : 0.f);
            dist = __fadd_rn(dist, __fmul_rn(term, pr.vel_w));
        }
        cost[e] = dist;
        // admissible-edge statistics for the assignment fast path (lapjv.cuh)
        if ((double)dist < lap_limit) {
            atomicAdd(&sm.row_cnt[r], 1);
            atomicAdd(&sm.col_cnt[c], 1);
            sm.row_adm[r] = c;
        }
        if (dist != dist) sm.adm_misc[0] = 1;
        if (fabs((double)dist - lap_limit) <= 2.0e-9) sm.adm_misc[1] = 1;
    }
    __syncthreads();
    if ((g_skip >> 6) & 1)                   // diagnostics: mirror the matrix for st_debug_read_cost
        for (int e = threadIdx.x; e < R * C; e += kThreads) cost_global[e] = cost[e];
    // ---- assignment (warp 0) -----------------------------------------------------
    if (threadIdx.x < 32 && (g_skip >> (stage_code - 1) & 1)) {
        for (int c = threadIdx.x; c < C; c += 32) sm.lap.y[c] = -1;
    } else if (threadIdx.x < 32) {
        stlap::Problem P;
        P.cost = cost; P.R = R; P.C = C; P.n = R + C; P.half = pr.lap_half;
        P.dbg = st.debug + (sbase / pr.T) * ST_DEBUG_WORDS;
        P.why = nullptr; P.no_greedy = 0;
        P.row_cnt = sm.row_cnt; P.col_cnt = sm.col_cnt; P.row_adm = sm.row_adm; P.adm_misc = sm.adm_misc;
        P.mark(0, stage_code); P.mark(1, R); P.mark(2, C);
        // debug words 5/6: assignments solved / solved by the exact fallback (this handle's lifetime)
        if (threadIdx.x == 0) P.dbg[5] += 1;
        const bool ok = stlap::assign_warp(P, sm.lap, P.dbg + 6);
        if (!ok && threadIdx.x == 0) atomicOr(status_word, ST_STATUS_LAP_GUARD);
    }
    __syncthreads();
    // col -> row; rows >= R are the solver's dummy rows (= unmatched)
    for (int c = threadIdx.x; c < C; c += kThreads) {
        const int r = sm.lap.y[c];
        if (r >= 0 && r < R) {
            sm.det_slot[sm.dets[c]] = sm.rows[r];
            sm.det_stage[sm.dets[c]] = (unsigned char)stage_code;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
k4_track_step(TrackState st, Params pr, const int32_t *__restrict__ frame_ids,
              const float *__restrict__ det_boxes, const float *__restrict__ det_scores,
              const int64_t *__restrict__ det_labels, const float *__restrict__ det_scales,
              const float *__restrict__ det_depth, const int32_t *__restrict__ det_counts, int det_cap,
              float *cost_ws /* written and re-read in this kernel: no __restrict__ */, float *__restrict__ out_boxes, float *__restrict__ out_scores,
              int64_t *__restrict__ out_labels, float *__restrict__ out_scales,
              float *__restrict__ out_depth, int64_t *__restrict__ out_ids,
              int32_t *__restrict__ out_counts, int32_t *__restrict__ out_src /* may be null */,
              const float *__restrict__ cmc_warp /* [S,6] camera-motion affine, may be null */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int sh_cnt[4];
    __shared__ int sh_nlive;
    const int s = blockIdx.x;
    const int T = pr.T, D = pr.D;
    // debug words 8..14: SM cycles the last step spent in each phase (thread 0's clock)
    long long t_prev = clock64();
#define ST_PHASE(idx)                                                         \
    do {                                                                      \
        if (threadIdx.x == 0) {                                               \
            const long long t_now = clock64();                                \
            st.debug[s * ST_DEBUG_WORDS + (idx)] = (int)(t_now - t_prev);                 \
            t_prev = t_now;                                                   \
        }                                                                     \
    } while (0)
    Smem sm = carve(smem_raw, T, D);
    const long long sbase = (long long)s * T;
    const int frame = frame_ids[s];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *kfws = sm.kf + warp * stkf::kScratchDoubles;
    int32_t *order = st.order + sbase;
    int32_t *free_slots = st.free_slots + sbase;
    float *cost = cost_ws + (long long)s * T * D;
    DetView dv;
    dv.boxes = det_boxes + (long long)s * det_cap * 4;
    dv.scores = det_scores + (long long)s * det_cap;
    dv.labels = det_labels + (long long)s * det_cap;
    dv.scales = det_scales + (long long)s * det_cap;
    dv.depth = det_depth + (long long)s * det_cap;
    int nd_in = det_counts[s];
    if (nd_in > D) {
        if (threadIdx.x == 0) atomicOr(&st.status[s], ST_STATUS_DET_OVERFLOW);
        nd_in = D;
    }
    if (nd_in > det_cap) nd_in = det_cap;

    // frame 0: reset() (:385-387)
    if (frame == 0) {
        for (int i = threadIdx.x; i < T; i += kThreads) free_slots[i] = T - 1 - i;
        if (threadIdx.x == 0) { st.n_live[s] = 0; st.n_free[s] = T; st.next_id[s] = 0; }
    }
    __syncthreads();
    if (threadIdx.x == 0) sh_nlive = st.n_live[s];
    for (int i = threadIdx.x; i < nd_in; i += kThreads) {
        sm.det_slot[i] = -1;
        sm.det_stage[i] = 0;
        const float4 b = reinterpret_cast<const float4 *>(dv.boxes)[i];
        sm.in_box[i * 4 + 0] = b.x; sm.in_box[i * 4 + 1] = b.y; sm.in_box[i * 4 + 2] = b.z; sm.in_box[i * 4 + 3] = b.w;
        sm.in_score[i] = dv.scores[i];
    }
    __syncthreads();
    const int n_live = sh_nlive;
    for (int p = threadIdx.x; p < n_live; p += kThreads) {
        const int slot = order[p];
        sm.ord_slot[p] = slot;
        sm.slot_pos[slot] = p;
        sm.ord_tent[p] = st.tentative[sbase + slot];
        sm.ord_tracked[p] = st.tracked[sbase + slot];
    }
    __syncthreads();
    int n_out = 0;

    if (n_live == 0 || nd_in == 0) {
        // ---- init-only branch (:391-403): keep scores > init_track_thr, all new ids
        compact_warp0(nd_in, [&](int i) { return sm.in_score[i] > pr.init_track_thr; }, sm.out_src, &sh_cnt[0]);
        n_out = sh_cnt[0];
        for (int i = threadIdx.x; i < n_out; i += kThreads) sm.det_stage[sm.out_src[i]] = 4;
        __syncthreads();
    } else {
        // ---- detection filter (:412-421)
        compact_warp0(nd_in, [&](int i) {
            const float *b = sm.in_box + i * 4;
            const float area = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
            return sm.in_score[i] > pr.obj_score_thr && area > 100.f;
        }, sm.dets, &sh_cnt[0]);
        int C = sh_cnt[0];
        for (int c = threadIdx.x; c < C; c += kThreads) sm.det_stage[sm.dets[c]] = 4;   // leftover unless matched
        for (int p = threadIdx.x; p < n_live; p += kThreads) sm.slot_hit[p] = 0;
        // ---- confirmed tracks: predict (:431-441)
        compact_warp0(n_live, [&](int p) { return !sm.ord_tent[p]; }, sm.rows, &sh_cnt[1]);
        int R = sh_cnt[1];
        for (int r = threadIdx.x; r < R; r += kThreads) sm.rows[r] = sm.ord_slot[sm.rows[r]];   // position -> slot
        __syncthreads();
        // four tracks per warp (kalman.cuh, 8-lane groups, state in registers)
        for (int r0 = warp * 4; r0 < R; r0 += kWarps * 4) {
            const int r = r0 + (lane >> 3), gl = lane & 7;
            const bool active = r < R;
            const long long g = active ? sbase + sm.rows[r] : 0;
            stkf::Row8 row = stkf::load8(st.mean + g * 8, st.cov + g * 64, gl, active);
            if (active) {
                if (st.last_frame[g] != frame - 1 && gl == 7) row.m = 0.0;        // lost last frame: vh = 0
                if (st.tracked[g]) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) st.saved_cov[g * 64 + gl * 8 + j] = row.p[j];
                    st.saved_mean[g * 8 + gl] = row.m;
                    if (gl == 0) st.has_saved[g] = 1;
                }
            }
            if (!(g_skip & 32)) stkf::predict8(row, gl);
            stkf::store8(row, st.mean + g * 8, st.cov + g * 64, gl, active);
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
            for (int r = warp; r < R; r += kWarps) {
                const long long g = sbase + sm.rows[r];
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
                    st.mean[g * 8 + 0] = __dadd_rn(__dadd_rn(__dmul_rn(r00, m[0]), __dmul_rn(r01, m[1])), t0);
                    st.mean[g * 8 + 1] = __dadd_rn(__dadd_rn(__dmul_rn(r10, m[0]), __dmul_rn(r11, m[1])), t1);
                    st.mean[g * 8 + 3] = __dmul_rn(m[3], sc);
                    st.mean[g * 8 + 4] = __dadd_rn(__dmul_rn(r00, m[4]), __dmul_rn(r01, m[5]));
                    st.mean[g * 8 + 5] = __dadd_rn(__dmul_rn(r10, m[4]), __dmul_rn(r11, m[5]));
                    st.mean[g * 8 + 7] = __dmul_rn(m[7], sc);
                }
                __syncwarp();
            }
            __syncthreads();
        }
        ST_PHASE(8);
        // ---- stage 1: confirmed tracks vs filtered dets (:448-474)
        associate(st, pr, sm, sbase, R, C, true, dv, cost, 1, &st.status[s]);
        ST_PHASE(9);
        // ---- stage 2: tentative tracks vs leftovers (:476-503)
        compact_warp0(nd_in, [&](int i) { return sm.det_stage[i] == 4; }, sm.dets, &sh_cnt[0]);
        C = sh_cnt[0];
        compact_warp0(n_live, [&](int p) { return sm.ord_tent[p] != 0; }, sm.rows, &sh_cnt[1]);
        R = sh_cnt[1];
        for (int r = threadIdx.x; r < R; r += kThreads) sm.rows[r] = sm.ord_slot[sm.rows[r]];
        __syncthreads();
        associate(st, pr, sm, sbase, R, C, true, dv, cost, 2, &st.status[s]);
        ST_PHASE(10);
        // ---- which tracks are matched so far
        for (int i = threadIdx.x; i < nd_in; i += kThreads) {
            const int slot = sm.det_slot[i];
            if (slot >= 0) sm.slot_hit[sm.slot_pos[slot]] = 1;
        }
        __syncthreads();
        // ---- stage 3: OCR, every unmatched track's last observation (:505-569)
        compact_warp0(nd_in, [&](int i) { return sm.det_stage[i] == 4; }, sm.dets, &sh_cnt[0]);
        C = sh_cnt[0];
        compact_warp0(n_live, [&](int p) { return !sm.slot_hit[p]; }, sm.rows, &sh_cnt[1]);
        R = sh_cnt[1];
        for (int r = threadIdx.x; r < R; r += kThreads) sm.rows[r] = sm.ord_slot[sm.rows[r]];
        __syncthreads();
        associate(st, pr, sm, sbase, R, C, false, dv, cost, 3, &st.status[s]);
        ST_PHASE(11);
        for (int i = threadIdx.x; i < nd_in; i += kThreads) {
            const int slot = sm.det_slot[i];
            if (slot >= 0 && sm.det_stage[i] == 3) sm.slot_hit[sm.slot_pos[slot]] = 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) st.debug[s * ST_DEBUG_WORDS + 7] = 8;
        // ---- online smoothing of re-found lost tracks (:571-577, :319-343)
        for (int i = warp; i < nd_in; i += kWarps) {
            const int slot = sm.det_slot[i];
            if (slot < 0) continue;
            const long long g = sbase + slot;
            if (sm.ord_tracked[sm.slot_pos[slot]] || (g_skip & 8)) continue;      // tracked flags only change in the update below
            const int miss = st.miss_run[g];
            float last[4], shift[4];
            const float denom = (float)(miss + 1);
            for (int k = 0; k < 4; ++k) {
                last[k] = st.last_obs[g * 4 + k];
                shift[k] = __fdiv_rn(__fsub_rn(dv.boxes[i * 4 + k], last[k]), denom);
            }
            // The chain of `miss` virtual updates runs on the register-resident filter: state
            // loaded once, stored once (it was a global load + store of the 8x8 covariance per
            // virtual update: 90k cycles for a track lost for 29 frames, the longest path of the
            // whole kernel).  The four 8-lane groups of the warp compute the same track.
            const int gl = lane & 7;
            const bool saved = st.has_saved[g] != 0;
            stkf::Row8 row = stkf::load8(saved ? st.saved_mean + g * 8 : st.mean + g * 8,
                                         saved ? st.saved_cov + g * 64 : st.cov + g * 64, gl, true);
            for (int t = 0; t < miss; ++t) {
                float vb[4], zc[4];
                double z[4];
                for (int k = 0; k < 4; ++k) vb[k] = __fadd_rn(last[k], __fmul_rn((float)(t + 1), shift[k]));
                stkf::xyxy_to_cxcyah(vb, zc);
                for (int k = 0; k < 4; ++k) z[k] = (double)zc[k];
                stkf::update8(row, z, gl);
            }
            stkf::store8(row, st.mean + g * 8, st.cov + g * 64, gl, lane < 8);
        }
        __syncthreads();
        if (threadIdx.x == 0) st.debug[s * ST_DEBUG_WORDS + 7] = 9;
        // ---- unmatched tracks: tracked = False, obs.append(None) (:579-582)
        for (int p = threadIdx.x; p < n_live; p += kThreads)
            if (!sm.slot_hit[p]) {
                const long long g = sbase + sm.ord_slot[p];
                st.tracked[g] = 0;
                push_obs(st, g, nullptr);
            }
        // ---- output order: stage 1, 2, 3 matches then leftovers (:584-596)
        if (threadIdx.x < 32) {
            int total = 0;
            for (int code = 1; code <= 4; ++code)
                for (int i0 = 0; i0 < nd_in; i0 += 32) {
                    const int i = i0 + lane;
                    const bool f = i < nd_in && sm.det_stage[i] == code;
                    const unsigned b = __ballot_sync(0xffffffffu, f);
                    if (f) sm.out_src[total + __popc(b & ((1u << lane) - 1u))] = i;
                    total += __popc(b);
                }
            if (lane == 0) sh_cnt[0] = total;
        }
        __syncthreads();
        n_out = sh_cnt[0];
    }

    if (threadIdx.x == 0) st.debug[s * ST_DEBUG_WORDS + 7] = 10;
    ST_PHASE(12);
    // ---- ids + BaseTracker.update (:598-605, base_tracker.py:54-91) ------------
    // new tracks are exactly the outputs without a slot; they are last in order
    __syncthreads();
    if (threadIdx.x == 0) {
        int first_new = n_out;
        for (int k = 0; k < n_out; ++k)
            if (sm.det_slot[sm.out_src[k]] < 0) { first_new = k; break; }
        sh_cnt[1] = first_new;
    }
    __syncthreads();
    const int first_new = sh_cnt[1];
    const int n_new = n_out - first_new;
    const long long id0 = st.next_id[s];
    const int n_free = st.n_free[s];
    int n_create = n_new;
    if (n_live + n_new > T) {
        n_create = T - n_live;
        if (threadIdx.x == 0) atomicOr(&st.status[s], ST_STATUS_TRACK_OVERFLOW);
    }
    __syncthreads();
    // four output rows per warp: 8-lane groups run the Kalman update / track birth of their row
    for (int k0 = warp * 4; k0 < n_out; k0 += kWarps * 4) {
        const int k = k0 + (lane >> 3), gl = lane & 7;
        const bool active = k < n_out;
        const int i = active ? sm.out_src[k] : 0;
        const float *box = dv.boxes + i * 4;
        float zc[4];
        stkf::xyxy_to_cxcyah(box, zc);
        int slot = active ? sm.det_slot[i] : -1;
        long long id = 0;
        const bool upd = active && slot >= 0;
        {
            // update_track (:128-146 + kalman_tracker_base.py:65-75); every lane runs the shuffles
            const long long g = upd ? sbase + slot : 0;
            double z[4];
            for (int q = 0; q < 4; ++q) z[q] = (double)zc[q];
            stkf::Row8 row = stkf::load8(st.mean + g * 8, st.cov + g * 64, gl, upd);
            if (!(g_skip & 16)) stkf::update8(row, z, gl);
            stkf::store8(row, st.mean + g * 8, st.cov + g * 64, gl, upd);
            if (upd) id = st.id[g];
            if (upd && gl == 0) {
                st.hits[g] += 1;
                st.last_frame[g] = frame;
                if (st.tentative[g] && st.hits[g] >= pr.num_tentatives) st.tentative[g] = 0;
                st.tracked[g] = 1;
                push_obs(st, g, box);
                vel_direction(k_step_obs(st, g, pr.vel_dt), box, st.velocity + g * 2);
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
                    st.num_obs[g] = 0;
                    st.miss_run[g] = 0;
                    push_obs(st, g, box);
                    st.velocity[g * 2] = -1.f;
                    st.velocity[g * 2 + 1] = -1.f;
                    order[n_live + r] = slot;
                }
            }
        }
        if (active && gl == 0) {
            const long long o = (long long)s * det_cap + k;
            if (slot >= 0) {
                const long long g = sbase + slot;
                for (int q = 0; q < 4; ++q) st.last_box[g * 4 + q] = box[q];
                st.last_score[g] = dv.scores[i];
                st.last_scale[g] = dv.scales[i];
                st.last_depth[g] = dv.depth[i];
                st.last_label[g] = dv.labels[i];
            }
            for (int q = 0; q < 4; ++q) out_boxes[o * 4 + q] = box[q];
            out_scores[o] = dv.scores[i];
            out_labels[o] = dv.labels[i];
            out_scales[o] = dv.scales[i];
            out_depth[o] = dv.depth[i];
            out_ids[o] = id;
            if (out_src) out_src[o] = i;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) st.debug[s * ST_DEBUG_WORDS + 7] = 11;
    ST_PHASE(13);
    // ---- pop_invalid_tracks (kalman_tracker_base.py:77-88) -------------------------
    const int n_now = n_live + n_create;
    if (threadIdx.x < 32) {
        int kept = 0, freed = 0;
        for (int p0 = 0; p0 < n_now; p0 += 32) {
            const int p = p0 + lane;
            int slot = -1;
            bool dead = false;
            if (p < n_now) {
                slot = order[p];
                const long long g = sbase + slot;
                const int lf = st.last_frame[g];
                dead = (frame - lf >= pr.retain) || (st.tentative[g] && lf != frame);
            }
            const unsigned alive_b = __ballot_sync(0xffffffffu, p < n_now && !dead);
            const unsigned dead_b = __ballot_sync(0xffffffffu, p < n_now && dead);
            __syncwarp();
            if (p < n_now && !dead) order[kept + __popc(alive_b & ((1u << lane) - 1u))] = slot;
            if (p < n_now && dead)
                free_slots[(n_free - n_create) + freed + __popc(dead_b & ((1u << lane) - 1u))] = slot;
            kept += __popc(alive_b);
            freed += __popc(dead_b);
            __syncwarp();
        }
        if (lane == 0) {
            st.n_live[s] = kept;
            st.n_free[s] = n_free - n_create + freed;
            st.next_id[s] = id0 + n_new;
            out_counts[s] = n_out;
            st.debug[s * ST_DEBUG_WORDS + 7] = 12;
        }
    }
}

__global__ void k4_reset(TrackState st, int T, const uint8_t *__restrict__ mask) {
    const int s = blockIdx.x;
    if (mask && !mask[s]) return;
    for (int i = threadIdx.x; i < T; i += blockDim.x) st.free_slots[(long long)s * T + i] = T - 1 - i;
    if (threadIdx.x == 0) { st.n_live[s] = 0; st.n_free[s] = T; st.next_id[s] = 0; }
}

__global__ void __launch_bounds__(32)
k5_lapjv_batch(int R, int C, int n_alloc /* nodes the scratch is sized for: the handle's max_tracks + max_dets, as in the
                                             tracker kernel, so that the sparse path finds the same arena */,
               const float *__restrict__ cost, double half, int mode,
               int32_t *__restrict__ row2col, int32_t *__restrict__ col2row, int32_t *__restrict__ status,
               int32_t *__restrict__ fallbacks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x, lane = threadIdx.x;
    stlap::Scratch sc = stlap::carve(smem_raw, n_alloc);
    stlap::Problem P;
    P.cost = cost + (long long)b * R * C; P.R = R; P.C = C; P.n = R + C; P.half = half;
    P.dbg = nullptr;
    P.why = nullptr; P.no_greedy = 0;
    int *stats = reinterpret_cast<int *>(smem_raw + stlap::scratch_bytes(n_alloc));
    P.row_cnt = stats; P.row_adm = stats + R; P.col_cnt = stats + 2 * R; P.adm_misc = stats + 2 * R + C;
    if (R > 0 && C > 0) {
        stlap::count_admissible(P);
        bool ok;
        if (mode == 1) ok = stlap::solve_warp(P, sc);                 // exact emulation only
        else if (mode == 2 && R <= 64 && C <= 64) {
            // component path of the round-2 tracker kernel, on one warp
            stlap::Masks mk = stlap::carve_masks(smem_raw + ((stlap::scratch_bytes(n_alloc) + sizeof(int) * (2 * R + C + 4) + 15) & ~(size_t)15));
            stlap::build_masks(P, mk);
            stlap::prepare_components(P, sc, mk);
            bool bad = *mk.fallback != 0;
            if (!bad) bad = !stlap::solve_big_components(P, sc, mk, reinterpret_cast<unsigned char *>(mk.fallback + 3));
            for (int k = 0; !bad && k < *mk.n_comp; ++k)
                bad = !stlap::enumerate_component(P, sc, mk.comp_rows[k], mk.comp_cols[k]);
            ok = true;
            if (bad) {
                if (lane == 0) atomicAdd(fallbacks, 1);
                __syncwarp();
                ok = stlap::assign_warp(P, sc, nullptr);
            }
        } else if (mode == 2) {
            // crowded / wide problems: the sparse certified solver of the round-2 tracker kernel
            __shared__ int hdr[4];
            const int verdict = stlap::sparse_assign_block(P, sc, (size_t)24 * n_alloc, hdr, nullptr);
            ok = true;
            if (verdict != 0) {
                if (lane == 0) atomicAdd(fallbacks, 1);
                __syncwarp();
                ok = (verdict == 1 || verdict == 6 || verdict == 7) ? stlap::solve_warp(P, sc) : stlap::assign_warp(P, sc, nullptr);
            }
        } else {
            ok = R + C <= 512 && stlap::fast_assign_warp(P, sc);       // certified fast path
            if (!ok) {
                if (lane == 0) atomicAdd(fallbacks, 1);
                __syncwarp();
                ok = stlap::solve_warp(P, sc);
            }
        }
        if (!ok && lane == 0) atomicOr(status, ST_STATUS_LAP_GUARD);
        __syncwarp();
    }
    for (int i = lane; i < R; i += 32) {
        int j = (R > 0 && C > 0) ? sc.x[i] : -1;
        row2col[(long long)b * R + i] = (j >= 0 && j < C) ? j : -1;
    }
    for (int j = lane; j < C; j += 32) {
        int i = (R > 0 && C > 0) ? sc.y[j] : -1;
        col2row[(long long)b * C + j] = (i >= 0 && i < R) ? i : -1;
    }
}

__global__ void k4_kalman_batch(int op, int n, double *mean, double *cov, const float *z,
                                const double *__restrict__ noise_scale /* [n] 1 - bbox_score, or null */) {
    // 4 warps per CTA, 4 tracks per warp (the 8-lane-group filter of kalman.cuh)
    const int lane = threadIdx.x & 31, gl = lane & 7;
    const int t = (blockIdx.x * 4 + (threadIdx.x >> 5)) * 4 + (lane >> 3);
    const bool active = t < n;
    const long long g = active ? t : 0;
    if (op == 0) {
        if (active) stkf::initiate8(mean + g * 8, cov + g * 64, z + g * 4, gl);
        return;
    }
    stkf::Row8 row = stkf::load8(mean + g * 8, cov + g * 64, gl, active);
    if (op == 1) stkf::predict8(row, gl);
    else {
        double zz[4];
        for (int k = 0; k < 4; ++k) zz[k] = active ? (double)z[g * 4 + k] : 0.0;
        stkf::update8(row, zz, gl, (active && noise_scale) ? noise_scale[g] : 1.0);
    }
    stkf::store8(row, mean + g * 8, cov + g * 64, gl, active);
}

// KalmanFilter.project (kalman_filter.py:121-152) and gating_distance (:190-234) for n states,
// one thread per state: S = H P H^T + diag((h/20)^2, (h/20)^2, 1e-1^2, (h/20)^2) is the upper-left
// 4x4 block of P plus the diagonal (H selects, so the products are exact); the squared Mahalanobis
// distance of every measurement through the lower Cholesky factor and a forward substitution.
__global__ void k4_kalman_project(int n, const double *__restrict__ mean, const double *__restrict__ cov,
                                  double *__restrict__ pmean, double *__restrict__ pcov, int m,
                                  const double *__restrict__ meas, int only_position, double *__restrict__ dist,
                                  const double *__restrict__ noise_scale /* [n] 1 - bbox_score (use_nsa), or null */) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double *mu = mean + (long long)t * 8, *P = cov + (long long)t * 64;
    const double ns = noise_scale ? noise_scale[t] : 1.0;          // x * 1.0 is exact
    const double sp = __dmul_rn(ns, __dmul_rn(1.0 / 20, mu[3])), sa = __dmul_rn(ns, 1e-1);
    const double r[4] = {__dmul_rn(sp, sp), __dmul_rn(sp, sp), __dmul_rn(sa, sa), __dmul_rn(sp, sp)};
    double A[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) A[i][j] = __dadd_rn(P[i * 8 + j], i == j ? r[i] : 0.0);
    if (pmean)
        for (int i = 0; i < 4; ++i) pmean[(long long)t * 4 + i] = mu[i];
    if (pcov)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) pcov[(long long)t * 16 + i * 4 + j] = A[i][j];
    if (!dist) return;
    const int nd = only_position ? 2 : 4;
    double L[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < nd) {
            double acc = A[j][j];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < j) acc = __dsub_rn(acc, __dmul_rn(L[j][k], L[j][k]));
            L[j][j] = sqrt(acc);                          // not positive definite -> NaN (NumPy raises LinAlgError)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i > j && i < nd) {
                    double a = A[i][j];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < j) a = __dsub_rn(a, __dmul_rn(L[i][k], L[j][k]));
                    L[i][j] = __ddiv_rn(a, L[j][j]);
                }
        }
    }
    for (int q = 0; q < m; ++q) {
        double z[4], acc2 = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < nd) {
                double a = __dsub_rn(meas[(long long)q * 4 + i], mu[i]);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < i) a = __dsub_rn(a, __dmul_rn(L[i][k], z[k]));
                z[i] = __ddiv_rn(a, L[i][i]);
                acc2 = __dadd_rn(acc2, __dmul_rn(z[i], z[i]));
            }
        dist[(long long)t * m + q] = acc2;
    }
}

template <typename Tp>
void take(Tp *&ptr, unsigned char *&cursor, size_t count) {
    ptr = reinterpret_cast<Tp *>(cursor);
    cursor += (sizeof(Tp) * count + 255) & ~(size_t)255;
}

// lays out every device array of the handle in one allocation; with base ==
// nullptr only measures
size_t layout(st_handle *h, unsigned char *base) {
    const size_t S = h->cfg.max_sequences, T = h->cfg.max_tracks, D = h->cfg.max_dets;
    const size_t N = h->cfg.max_candidates;
    unsigned char *c = base;
    TrackState &st = h->st;
    take(st.n_live, c, S); take(st.next_id, c, S); take(st.order, c, S * T); take(st.free_slots, c, S * T);
    take(st.n_free, c, S); take(st.status, c, S); take(st.debug, c, S * ST_DEBUG_WORDS);
    take(st.id, c, S * T); take(st.hits, c, S * T); take(st.last_frame, c, S * T);
    take(st.tentative, c, S * T); take(st.tracked, c, S * T); take(st.has_saved, c, S * T);
    take(st.mean, c, S * T * 8); take(st.cov, c, S * T * 64);
    take(st.saved_mean, c, S * T * 8); take(st.saved_cov, c, S * T * 64);
    take(st.velocity, c, S * T * 2); take(st.obs_ring, c, S * T * ST_RING * 4);
    take(st.obs_valid, c, S * T * ST_RING); take(st.num_obs, c, S * T);
    take(st.last_obs, c, S * T * 4); take(st.miss_run, c, S * T);
    take(st.last_box, c, S * T * 4); take(st.last_score, c, S * T); take(st.last_scale, c, S * T);
    take(st.last_depth, c, S * T); take(st.last_label, c, S * T);
    take(h->cand_count, c, S); take(h->cand_score, c, S * N); take(h->cand_prior, c, S * N);
    take(h->cand_label, c, S * N);
    take(h->cost, c, S * T * D * 3);
    return (size_t)(c - base);
}

}  // namespace

extern "C" int st_abi_version(void) { return ST_ABI_VERSION; }

extern "C" const char *st_error_string(int code) {
    switch (code) {
        case ST_OK: return "ok";
        case ST_ERR_INVALID_ARG: return "invalid argument";
        case ST_ERR_CUDA: return "CUDA runtime error";
        case ST_ERR_NO_DEVICE: return "no CUDA device";
        case ST_ERR_CAPACITY: return "capacity exceeded";
        case ST_ERR_ALLOC: return "device allocation failed";
        default: return "unknown error";
    }
}

extern "C" void st_default_config(st_config *cfg) {
    if (!cfg) return;
    cfg->max_sequences = 1; cfg->max_tracks = 256; cfg->max_dets = 256; cfg->max_candidates = 2048;
    // constructor defaults of the reference classes (see the header)
    cfg->obj_score_thr = 0.3f; cfg->init_track_thr = 0.7f; cfg->weight_iou_with_det_scores = 1;
    cfg->match_iou_thr = 0.3; cfg->num_tentatives = 3; cfg->vel_consist_weight = 0.2f;
    cfg->vel_delta_t = 3; cfg->num_frames_retain = 10;
    cfg->baseline = 0.25f; cfg->focal_length = 640.f; cfg->score_thr = 0.01f; cfg->nms_iou_thr = 0.5f;
}

extern "C" int st_create(const st_config *cfg, st_handle **out) {
    if (!cfg || !out) return ST_ERR_INVALID_ARG;
    if (cfg->max_sequences < 1 || cfg->max_tracks < 1 || cfg->max_tracks > 1024 || cfg->max_dets < 1 ||
        cfg->max_dets > 1024 || cfg->max_candidates < 1 || cfg->vel_delta_t < 0 ||
        cfg->vel_delta_t >= ST_RING || cfg->num_tentatives < 0)
        return ST_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ST_ERR_NO_DEVICE;
    st_handle *h = new (std::nothrow) st_handle();
    if (!h) return ST_ERR_ALLOC;
    h->cfg = *cfg;
    cudaGetDevice(&h->device);
    h->all_bytes = layout(h, nullptr);
    if (cudaMalloc(&h->all_mem, h->all_bytes) != cudaSuccess) { delete h; return ST_ERR_ALLOC; }
    layout(h, static_cast<unsigned char *>(h->all_mem));
    if (cudaMemset(h->all_mem, 0, h->all_bytes) != cudaSuccess) { cudaFree(h->all_mem); delete h; return ST_ERR_CUDA; }
    const size_t smem = smem_bytes(cfg->max_tracks, cfg->max_dets);
    if (st_k2_configure(cfg->max_candidates) != ST_OK) { cudaFree(h->all_mem); delete h; return ST_ERR_CAPACITY; }
    if (cudaFuncSetAttribute(k4_track_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(k5_lapjv_batch, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(stlap::scratch_bytes(cfg->max_tracks + cfg->max_dets) +
                                   sizeof(int) * (2 * cfg->max_tracks + cfg->max_dets + 4) + 32 +
                                   stlap::masks_bytes() + stlap::sub_bytes())) != cudaSuccess) {
        cudaFree(h->all_mem); delete h; return ST_ERR_CAPACITY;
    }
    {
        const char *e = getenv("ST_K4_V1");
        h->k4_version = (e && atoi(e) != 0) || st_k4v2_configure(cfg->max_tracks, cfg->max_dets) != ST_OK ? 1 : 2;
        cudaGetLastError();
    }
    k4_reset<<<cfg->max_sequences, 128>>>(h->st, cfg->max_tracks, nullptr);
    if (cudaDeviceSynchronize() != cudaSuccess) { cudaFree(h->all_mem); delete h; return ST_ERR_CUDA; }
    *out = h;
    return ST_OK;
}

extern "C" int st_destroy(st_handle *h) {
    if (!h) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    cudaFree(h->all_mem);
    delete h;
    return ST_OK;
}

extern "C" int st_reset(st_handle *h, const uint8_t *seq_mask, void *stream) {
    if (!h) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    k4_reset<<<h->cfg.max_sequences, 128, 0, (cudaStream_t)stream>>>(h->st, h->cfg.max_tracks, seq_mask);
    return st_last_launch();
}

// ---- checkpoint / resume ----------------------------------------------------------------
// Every device array of a handle lives in one allocation (layout()), so a snapshot is that
// allocation behind a small header; loading checks that the capacities that determine the
// layout are the same.  The thresholds of the loading handle stay its own.
namespace {
struct SnapshotHeader {
    uint32_t magic, abi;
    int32_t max_sequences, max_tracks, max_dets, max_candidates;
    uint64_t payload;
};
constexpr uint32_t kSnapshotMagic = 0x4b525453u;       // "STRK"
}  // namespace

extern "C" long long st_snapshot_bytes(st_handle *h) {
    return h ? (long long)(sizeof(SnapshotHeader) + h->all_bytes) : (long long)ST_ERR_INVALID_ARG;
}

extern "C" int st_snapshot_save(st_handle *h, void *buf_host, long long bytes, void *stream) {
    if (!h || !buf_host || bytes < st_snapshot_bytes(h)) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    SnapshotHeader hd{kSnapshotMagic, (uint32_t)ST_ABI_VERSION, h->cfg.max_sequences, h->cfg.max_tracks,
                      h->cfg.max_dets, h->cfg.max_candidates, (uint64_t)h->all_bytes};
    memcpy(buf_host, &hd, sizeof(hd));
    cudaStream_t st = (cudaStream_t)stream;
    ST_CUDA_OK(cudaMemcpyAsync(static_cast<unsigned char *>(buf_host) + sizeof(hd), h->all_mem, h->all_bytes,
                               cudaMemcpyDeviceToHost, st));
    ST_CUDA_OK(cudaStreamSynchronize(st));
    return ST_OK;
}

extern "C" int st_snapshot_load(st_handle *h, const void *buf_host, long long bytes, void *stream) {
    if (!h || !buf_host || bytes < (long long)sizeof(SnapshotHeader)) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    SnapshotHeader hd;
    memcpy(&hd, buf_host, sizeof(hd));
    if (hd.magic != kSnapshotMagic || hd.abi != (uint32_t)ST_ABI_VERSION || hd.max_sequences != h->cfg.max_sequences ||
        hd.max_tracks != h->cfg.max_tracks || hd.max_dets != h->cfg.max_dets ||
        hd.max_candidates != h->cfg.max_candidates || hd.payload != (uint64_t)h->all_bytes ||
        bytes < (long long)(sizeof(hd) + hd.payload))
        return ST_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    ST_CUDA_OK(cudaMemcpyAsync(h->all_mem, static_cast<const unsigned char *>(buf_host) + sizeof(hd), h->all_bytes,
                               cudaMemcpyHostToDevice, st));
    ST_CUDA_OK(cudaStreamSynchronize(st));
    return ST_OK;
}

extern "C" int st_read_status(st_handle *h, int32_t *status_host, int S) {
    if (!h || !status_host || S < 0 || S > h->cfg.max_sequences) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    // the steps run on non-blocking streams the legacy default stream does not order against:
    // wait for everything in flight on the device, then read
    ST_CUDA_OK(cudaDeviceSynchronize());
    ST_CUDA_OK(cudaMemcpy(status_host, h->st.status, sizeof(int32_t) * S, cudaMemcpyDeviceToHost));
    return ST_OK;
}

extern "C" int st_read_live_counts(st_handle *h, int32_t *live_host, int S, void *stream) {
    if (!h || !live_host || S < 0 || S > h->cfg.max_sequences) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    ST_CUDA_OK(cudaMemcpyAsync(live_host, h->st.n_live, sizeof(int32_t) * S, cudaMemcpyDeviceToHost, st));
    ST_CUDA_OK(cudaStreamSynchronize(st));
    return ST_OK;
}

extern "C" int st_read_debug(st_handle *h, int32_t *debug_host, int S) {
    if (!h || !debug_host || S < 0 || S > h->cfg.max_sequences) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    // a side stream so that the words can be read while a kernel is still running
    cudaStream_t side;
    ST_CUDA_OK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    cudaError_t e = cudaMemcpyAsync(debug_host, h->st.debug, sizeof(int32_t) * ST_DEBUG_WORDS * S, cudaMemcpyDeviceToHost, side);
    if (e == cudaSuccess) e = cudaStreamSynchronize(side);
    cudaStreamDestroy(side);
    return e == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

extern "C" int st_debug_read_cost(st_handle *h, int seq, float *cost_host, int count) {
    if (!h || !cost_host || seq < 0 || seq >= h->cfg.max_sequences) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    ST_CUDA_OK(cudaDeviceSynchronize());
    // round-2 kernel: the third slot of the sequence's workspace (a stage-1 matrix the certified paths refused,
    // kept when diagnostics bit 9 is set); round-1 kernel: the mirror of the last matrix (bit 6)
    const size_t slot = h->k4_version == 2 ? (size_t)seq * 3 + 2 : (size_t)seq;
    ST_CUDA_OK(cudaMemcpy(cost_host, h->cost + slot * h->cfg.max_tracks * h->cfg.max_dets,
                          sizeof(float) * count, cudaMemcpyDeviceToHost));
    return ST_OK;
}

extern "C" int st_debug_set_skip(int mask) {
    ST_CUDA_OK(cudaMemcpyToSymbol(g_skip, &mask, sizeof(int)));
    return st_k4v2_set_skip(mask);
}

extern "C" int st_clear_status(st_handle *h, void *stream) {
    if (!h) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    ST_CUDA_OK(cudaMemsetAsync(h->st.status, 0, sizeof(int32_t) * h->cfg.max_sequences, (cudaStream_t)stream));
    return ST_OK;
}

extern "C" int st_track_step(st_handle *h, int S, const int32_t *frame_ids, const float *det_boxes,
                             const float *det_scores, const int64_t *det_labels, const float *det_scales,
                             const float *det_depth, const int32_t *det_counts, int det_cap,
                             float *out_boxes, float *out_scores, int64_t *out_labels, float *out_scales,
                             float *out_depth, int64_t *out_ids, int32_t *out_counts, int32_t *out_src,
                             const float *warp, void *stream) {
    if (!h || S < 0 || S > h->cfg.max_sequences || !frame_ids || !det_boxes || !det_scores || !det_labels ||
        !det_scales || !det_depth || !det_counts || det_cap < 1 || !out_boxes || !out_scores || !out_labels ||
        !out_scales || !out_depth || !out_ids || !out_counts)
        return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    if (S == 0) return ST_OK;
    Params pr;
    pr.T = h->cfg.max_tracks; pr.D = h->cfg.max_dets;
    pr.obj_score_thr = h->cfg.obj_score_thr; pr.init_track_thr = h->cfg.init_track_thr;
    pr.vel_w = h->cfg.vel_consist_weight; pr.weight_iou = h->cfg.weight_iou_with_det_scores;
    pr.num_tentatives = h->cfg.num_tentatives; pr.vel_dt = h->cfg.vel_delta_t;
    pr.retain = h->cfg.num_frames_retain;
    // lap.lapjv(cost_limit = 1 - match_iou_thr): Python float arithmetic on the config value
    pr.lap_half = (1.0 - h->cfg.match_iou_thr) / 2.0;
    {
        // a pair with IoU 0 costs 1 + w (angle - pi/2) / pi: above limit + 2e-4 as soon as the angle exceeds
        // a_req, i.e. its cosine is below cos(a_req)
        const double L = 2.0 * pr.lap_half, w = (double)pr.vel_w, pi = 3.14159265358979323846;
        pr.screen_cos = -2.f;
        pr.placeholder = (float)(L + 1e-3);
        if (w > 0.0 && w < 1e6) {
            const double a_req = pi / 2 + pi * (L + 2e-4 - 1.0) / w;
            if (a_req <= 0.0) pr.screen_cos = 2.f;
            else if (a_req < pi) pr.screen_cos = (float)(cos(a_req) - 1e-4);
        }
    }
    if (h->k4_version == 2)
        return st_k4v2_launch(h, pr, S, frame_ids, det_boxes, det_scores, det_labels, det_scales, det_depth, det_counts,
                              det_cap, out_boxes, out_scores, out_labels, out_scales, out_depth, out_ids, out_counts,
                              out_src, warp, (cudaStream_t)stream);
    const size_t smem = smem_bytes(pr.T, pr.D);
    k4_track_step<<<S, kThreads, smem, (cudaStream_t)stream>>>(
        h->st, pr, frame_ids, det_boxes, det_scores, det_labels, det_scales, det_depth, det_counts, det_cap,
        h->cost, out_boxes, out_scores, out_labels, out_scales, out_depth, out_ids, out_counts, out_src, warp);
    return st_last_launch();
}

extern "C" int st_export_state(st_handle *h, int seq, int64_t *ids_host, double *mean_host, double *cov_host,
                               float *velocity_host, uint8_t *flags_host, int32_t *hits_host,
                               int32_t *last_frame_host, int64_t *num_tracks_host) {
    if (!h || seq < 0 || seq >= h->cfg.max_sequences) return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    ST_CUDA_OK(cudaDeviceSynchronize());
    const int T = h->cfg.max_tracks;
    int n_live = 0;
    ST_CUDA_OK(cudaMemcpy(&n_live, h->st.n_live + seq, sizeof(int), cudaMemcpyDeviceToHost));
    if (num_tracks_host)
        ST_CUDA_OK(cudaMemcpy(num_tracks_host, h->st.next_id + seq, sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (n_live < 0 || n_live > T) return ST_ERR_CUDA;
    int32_t *order = new int32_t[T];
    cudaMemcpy(order, h->st.order + (size_t)seq * T, sizeof(int32_t) * T, cudaMemcpyDeviceToHost);
    int rc = ST_OK;
    for (int p = 0; p < n_live && rc == ST_OK; ++p) {
        const size_t g = (size_t)seq * T + order[p];
        uint8_t tent = 0, trk = 0;
        cudaError_t e = cudaSuccess;
        if (ids_host) e = cudaMemcpy(ids_host + p, h->st.id + g, sizeof(int64_t), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && mean_host) e = cudaMemcpy(mean_host + p * 8, h->st.mean + g * 8, sizeof(double) * 8, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && cov_host) e = cudaMemcpy(cov_host + p * 64, h->st.cov + g * 64, sizeof(double) * 64, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && velocity_host) e = cudaMemcpy(velocity_host + p * 2, h->st.velocity + g * 2, sizeof(float) * 2, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && hits_host) e = cudaMemcpy(hits_host + p, h->st.hits + g, sizeof(int32_t), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && last_frame_host) e = cudaMemcpy(last_frame_host + p, h->st.last_frame + g, sizeof(int32_t), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && flags_host) {
            e = cudaMemcpy(&tent, h->st.tentative + g, 1, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) e = cudaMemcpy(&trk, h->st.tracked + g, 1, cudaMemcpyDeviceToHost);
            flags_host[p] = (uint8_t)((tent ? 1 : 0) | (trk ? 2 : 0));
        }
        if (e != cudaSuccess) rc = ST_ERR_CUDA;
    }
    delete[] order;
    return rc == ST_OK ? n_live : rc;
}

extern "C" int st_kalman_batch(int op, int n, double *mean, double *cov, const float *z, void *stream) {
    return st_kalman_batch_scored(op, n, mean, cov, z, nullptr, stream);
}

extern "C" int st_kalman_batch_scored(int op, int n, double *mean, double *cov, const float *z,
                                      const double *noise_scale, void *stream) {
    if (op < 0 || op > 2 || n < 0 || !mean || !cov || (op != 1 && !z)) return ST_ERR_INVALID_ARG;
    if (n == 0) return ST_OK;
    k4_kalman_batch<<<(n + 15) / 16, 128, 0, (cudaStream_t)stream>>>(op, n, mean, cov, z, noise_scale);
    return st_last_launch();
}

extern "C" int st_kalman_project(int n, const double *mean, const double *cov, double *pmean, double *pcov,
                                 void *stream) {
    return st_kalman_project_scored(n, mean, cov, nullptr, pmean, pcov, stream);
}

extern "C" int st_kalman_project_scored(int n, const double *mean, const double *cov, const double *noise_scale,
                                        double *pmean, double *pcov, void *stream) {
    if (n < 0 || !mean || !cov || (!pmean && !pcov)) return ST_ERR_INVALID_ARG;
    if (n == 0) return ST_OK;
    k4_kalman_project<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, mean, cov, pmean, pcov, 0, nullptr, 0,
                                                                       nullptr, noise_scale);
    return st_last_launch();
}

extern "C" int st_kalman_gating(int n, const double *mean, const double *cov, int m, const double *meas,
                                int only_position, double *dist, void *stream) {
    if (n < 0 || m < 0 || !mean || !cov || !dist || (m > 0 && !meas)) return ST_ERR_INVALID_ARG;
    if (n == 0 || m == 0) return ST_OK;
    k4_kalman_project<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, mean, cov, nullptr, nullptr, m, meas,
                                                                       only_position != 0, dist, nullptr);
    return st_last_launch();
}

extern "C" int st_lapjv_batch(st_handle *h, int B, int R, int C, const float *cost, double cost_limit,
                              int mode, int32_t *row2col, int32_t *col2row, void *stream) {
    if (!h || B < 0 || R < 0 || C < 0 || R + C > h->cfg.max_tracks + h->cfg.max_dets || !row2col || !col2row ||
        (R > 0 && C > 0 && !cost))
        return ST_ERR_INVALID_ARG;
    StDeviceGuard guard(h->device);
    if (B == 0 || R + C == 0) return ST_OK;
    const int n_alloc = h->cfg.max_tracks + h->cfg.max_dets;
    k5_lapjv_batch<<<B, 32, stlap::scratch_bytes(n_alloc) + sizeof(int) * (2 * R + C + 4) + 32 + stlap::masks_bytes() +
                            stlap::sub_bytes(),
                     (cudaStream_t)stream>>>(
        R, C, n_alloc, cost, cost_limit / 2.0, mode, row2col, col2row, h->st.status, h->st.debug + 6);
    return st_last_launch();
}
