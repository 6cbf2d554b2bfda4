// Shared pieces of the tracker kernels (k4_track.cu: the round-1 kernel, k4_track2.cu: the
// round-2 kernel): launch parameters and the per-track helpers that restate
//   mmtrack/models/trackers/ocsort_tracker_disparity.py:140-185 (observations, velocity direction)
//   mmdet bbox_overlaps
#pragma once
#include "st_common.cuh"

namespace stk4 {

// diagnostics only (st_debug_set_skip): bit0..2 skip the assignment of stage
// 1..3, bit3 skip online smoothing, bit4 skip Kalman update, bit5 skip predict.  One copy per
// translation unit (no relocatable device code); st_debug_set_skip sets them all.
static __device__ int g_skip = 0;

struct Params {
    int T, D;                    // capacities
    float obj_score_thr, init_track_thr, vel_w;
    int weight_iou, num_tentatives, vel_dt, retain;
    double lap_half;             // (1 - match_iou_thr) / 2
    // screen of the round-2 kernel for pairs without overlap (k4_track2.cu ocm_cost): a pair whose
    // direction cosine is below screen_cos cannot cost less than the limit + 1e-4; such entries get
    // `placeholder` (-2 = screen off)
    float screen_cos, placeholder;
};

__host__ __device__ inline size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }


__device__ __forceinline__ float sum4(const float *b) {      // torch .sum() of 4 float32
    return __fadd_rn(__fadd_rn(__fadd_rn(b[0], b[1]), b[2]), b[3]);
}

// k_step_observation (:173-185): the observation vel_delta_t steps before the
// newest one if it exists and is not None, else the last non-None observation
__device__ inline const float *k_step_obs(const TrackState &st, long long g, int vel_dt) {
    const int n = st.num_obs[g];
    if (n > vel_dt) {
        const int r = (n - 1 - vel_dt) % ST_RING;
        if (st.obs_valid[g * ST_RING + r]) return st.obs_ring + (g * ST_RING + r) * 4;
    }
    return st.last_obs + g * 4;
}

__device__ inline void push_obs(const TrackState &st, long long g, const float *box) {
    const int n = st.num_obs[g];
    const int r = n % ST_RING;
    if (box) {
        for (int k = 0; k < 4; ++k) { st.obs_ring[(g * ST_RING + r) * 4 + k] = box[k]; st.last_obs[g * 4 + k] = box[k]; }
        st.obs_valid[g * ST_RING + r] = 1;
        st.miss_run[g] = 0;
    } else {
        st.obs_valid[g * ST_RING + r] = 0;
        st.miss_run[g] += 1;
    }
    st.num_obs[g] = n + 1;
}

// vel_direction (:148-156)
__device__ inline void vel_direction(const float *b1, const float *b2, float *out) {
    if (sum4(b1) < 0.f || sum4(b2) < 0.f) { out[0] = -1.f; out[1] = -1.f; return; }
    const float cx1 = __fmul_rn(__fadd_rn(b1[0], b1[2]), 0.5f), cy1 = __fmul_rn(__fadd_rn(b1[1], b1[3]), 0.5f);
    const float cx2 = __fmul_rn(__fadd_rn(b2[0], b2[2]), 0.5f), cy2 = __fmul_rn(__fadd_rn(b2[1], b2[3]), 0.5f);
    const float sy = __fsub_rn(cy2, cy1), sx = __fsub_rn(cx2, cx1);
    const float norm = __fadd_rn(__fsqrt_rn(__fadd_rn(__fmul_rn(sy, sy), __fmul_rn(sx, sx))), 1e-6f);
    out[0] = __fdiv_rn(sy, norm);
    out[1] = __fdiv_rn(sx, norm);
}

// mmdet bbox_overlaps(mode='iou', eps=1e-6) for one pair, float32
__device__ __forceinline__ float iou_pair(const float *a, const float *b) {
    const float a1 = __fmul_rn(__fsub_rn(a[2], a[0]), __fsub_rn(a[3], a[1]));
    const float a2 = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
    float w = __fsub_rn(fminf(a[2], b[2]), fmaxf(a[0], b[0]));
    float h = __fsub_rn(fminf(a[3], b[3]), fmaxf(a[1], b[1]));
    w = w < 0.f ? 0.f : w;           // clamp(min=0) keeps NaN
    h = h < 0.f ? 0.f : h;
    const float inter = __fmul_rn(w, h);
    float uni = __fsub_rn(__fadd_rn(a1, a2), inter);
    uni = uni < 1e-6f ? 1e-6f : uni; // torch.max(union, eps) (NaN stays NaN)
    return __fdiv_rn(inter, uni);
}

struct DetView {
    const float *boxes, *scores, *scales, *depth;
    const int64_t *labels;
};

}  // namespace stk4

// round-2 tracker kernel (k4_track2.cu)
int st_k4v2_configure(int max_tracks, int max_dets);
int st_k4v2_launch(st_handle *h, const stk4::Params &pr, int S, const int32_t *frame_ids, const float *det_boxes,
                   const float *det_scores, const int64_t *det_labels, const float *det_scales, const float *det_depth,
                   const int32_t *det_counts, int det_cap, float *out_boxes, float *out_scores, int64_t *out_labels,
                   float *out_scales, float *out_depth, int64_t *out_ids, int32_t *out_counts, int32_t *out_src,
                   const float *warp, cudaStream_t stream);
