// Shared declarations for the sm_100a kernels behind include/stereotrack_b200.h.
// Built with -fmad=false: float32/float64 arithmetic must round exactly as the
// reference's un-fused NumPy/torch expressions do (see DESIGN.md "Numerics").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/stereotrack_b200.h"

#define ST_RING 16            // observation ring capacity (vel_delta_t <= 15)
#define ST_MAX_LEVELS 8
#define ST_DEBUG_WORDS 32     // diagnostic words per sequence (st_read_debug)

#define ST_CUDA_OK(expr)                              \
    do {                                              \
        cudaError_t e__ = (expr);                     \
        if (e__ != cudaSuccess) return ST_ERR_CUDA;   \
    } while (0)

int st_k2_configure(int max_candidates);    // k2_decode_nms.cu

static inline int st_last_launch() {
    return cudaGetLastError() == cudaSuccess ? ST_OK : ST_ERR_CUDA;
}

// Per-sequence / per-slot track state, structure of arrays.  A slot index is
// seq * max_tracks + slot.  `order` lists the live slots of a sequence in
// ascending-id order (= the reference's dict insertion order).
struct TrackState {
    // per sequence
    int32_t *n_live;        // [S]
    int64_t *next_id;       // [S]  (self.num_tracks)
    int32_t *order;         // [S, T] live slots in dict order
    int32_t *free_slots;    // [S, T] stack of unused slots
    int32_t *n_free;        // [S]
    int32_t *status;        // [S] sticky ST_STATUS_* bits
    int32_t *debug;         // [S, ST_DEBUG_WORDS] diagnostics of the last track step (st_read_debug)
    // per slot
    int64_t *id;            // [S*T]
    int32_t *hits;          // len(track.bboxes)
    int32_t *last_frame;    // track.frame_ids[-1]
    uint8_t *tentative;
    uint8_t *tracked;
    uint8_t *has_saved;
    double *mean;           // [S*T, 8]
    double *cov;            // [S*T, 64]
    double *saved_mean;     // [S*T, 8]
    double *saved_cov;      // [S*T, 64]
    float *velocity;        // [S*T, 2]  (-1,-1) sentinel
    float *obs_ring;        // [S*T, ST_RING, 4]
    uint8_t *obs_valid;     // [S*T, ST_RING]   0 = None
    int32_t *num_obs;       // len(track.obs)
    float *last_obs;        // [S*T, 4] last non-None observation
    int32_t *miss_run;      // trailing None count
    float *last_box;        // [S*T, 4] memo: bboxes[-1]
    float *last_score;      // memo
    float *last_scale;
    float *last_depth;
    int64_t *last_label;
};

struct st_handle {
    st_config cfg;
    int device;
    TrackState st;
    // K2 workspace
    int32_t *cand_count;    // [S]
    float *cand_score;      // [S, N]
    int32_t *cand_prior;    // [S, N]
    int32_t *cand_label;    // [S, N]
    // K4/K5 workspace
    float *cost;            // [S, T*D]
    void *all_mem;          // single allocation backing everything above
    size_t all_bytes;
};

// ---- exactly rounded float32 transcendentals (DESIGN.md "Numerics") ---------
__device__ __forceinline__ float st_exp_f32(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float st_sigmoid_f32(float x) {
    return (float)(1.0 / (1.0 + exp(-(double)x)));
}
__device__ __forceinline__ float st_acos_f32(float x) { return (float)acos((double)x); }
