// Shared declarations for the sm_100a kernels behind include/stereotrack_b200.h.
// Built with -fmad=false: float32/float64 arithmetic must round exactly as the
// reference's un-fused NumPy/torch expressions do (see DESIGN.md "Numerics").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
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
int st_k4v2_set_skip(int mask);             // k4_track2.cu

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
    float *cost;            // [S, 3, T*D] all-pairs OCM / OCR costs and the dense matrix of a stage, when they
                            // do not fit the tracker kernel's shared memory
    int k4_version;         // 2 = k4_track2.cu (default), 1 = k4_track.cu (ST_K4_V1=1, or capacities beyond v2)
    void *all_mem;          // single allocation backing everything above
    size_t all_bytes;
};

// Per-frame INPUT pointers of the frame pipeline, resident on the device (one entry per ring
// slot).  The kernels of a captured frame read their inputs through it, so a CUDA graph
// is replayed on new inputs by updating this table (one tiny kernel) instead of patching
// kernel-node parameters.  nullptr table = use the pointer passed to the kernel.
struct StFrameInputs {
    const uint16_t *raw;
    const float *cls[ST_MAX_LEVELS], *bbox[ST_MAX_LEVELS], *obj[ST_MAX_LEVELS];
};
// K2's candidate scratch (one set per frame in flight)
struct StK2Scratch {
    int32_t *count;         // [S]
    float *score;           // [S, N]
    int32_t *prior;         // [S, N]
    int32_t *label;         // [S, N]
};
// internal forms of the public entry points that take the table / scratch (k1, k2, k3 raw)
int st_disp2depth_tab(const uint16_t *raw, const StFrameInputs *tab, int S, int H, int W, int Hr, int Wr, int Hp,
                      int Wp, float *depth, float *disp3, uint8_t *mask, float baseline, float focal_length,
                      void *stream);
int st_decode_nms_ws(st_handle *h, const float *const *cls_host, const float *const *bbox_host,
                     const float *const *obj_host, const StFrameInputs *tab, const StK2Scratch *ws, int num_levels,
                     const int32_t *level_h_host, const int32_t *level_w_host, const int32_t *strides_host, int S,
                     int num_classes, float scale_w, float scale_h, int ori_h, int ori_w, int det_cap, float *boxes,
                     float *scores, int64_t *labels, int32_t *prior_idx, int32_t *keep_idx, int32_t *counts,
                     void *stream);
int st_box_depth_raw_tab(const uint16_t *raw, const StFrameInputs *tab, int S, int H, int W, int Hp, int Wp,
                         float baseline, float focal_length, const float *boxes, const int32_t *counts, int cap,
                         float *d, float *scale, float *inflated, int32_t *workspace, const float *prev_boxes,
                         const int32_t *prev_index, const float *prev_d, void *stream);

// ---- exactly rounded float32 transcendentals (DESIGN.md "Numerics") ---------
__device__ __forceinline__ float st_exp_f32(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float st_sigmoid_f32(float x) {
    return (float)(1.0 / (1.0 + exp(-(double)x)));
}
__device__ __forceinline__ float st_acos_f32(float x) { return (float)acos((double)x); }

// Depth of one raw u16 disparity code, exactly as K1 writes it into the map: 65535 -> disparity 0
// (loading_disparity.py:129-134), disp = code / 16, Z = reciprocal(disp + 1e-6) * (B * f)
// (torch's evaluation of ocsort_disparity.py:132-134).  K1's FMA-only reciprocal is proven equal
// to the IEEE one on all 65 536 codes (tests/test_k1_reciprocal.py), so __frcp_rn is the same value.
__device__ __forceinline__ float st_code_depth(uint32_t code, float bf) {
    const float disp = code == 65535u ? 0.0f : __fmul_rn((float)code, 0.0625f);
    return __fmul_rn(__frcp_rn(__fadd_rn(disp, 1e-6f)), bf);
}
__host__ __device__ inline float st_bf(float baseline, float focal_length) {
    return (float)((double)baseline * (double)focal_length);
}

// ---- per-device library state ------------------------------------------------------------
// Everything the entry points set up lazily (SM count, shared-memory opt-ins, occupancy-derived
// grids, the event ring and the auxiliary streams K3 forks its large boxes onto) lives in one
// context per CUDA device, created under a mutex at first use on that device: two devices in one
// process, or an engine on a non-current device, each get their own.  st_ctx() returns the
// context of the CURRENT device; entry points that take a handle make the handle's device
// current for the duration of the call (StDeviceGuard).
struct StDeviceCtx;
StDeviceCtx *st_ctx();                       // k3_box_depth.cu; nullptr on a CUDA error
int st_ctx_sms();                            // SM count of the current device
// fork / join helper: the auxiliary stream paired with `caller` (created at first use, same
// priority) and two events of the device's ring
int st_ctx_aux(cudaStream_t caller, cudaStream_t *aux, cudaEvent_t *ev_fork, cudaEvent_t *ev_join);
int st_ctx_set_aux(cudaStream_t caller, cudaStream_t aux);
void st_ctx_drop_aux(cudaStream_t caller);

// Lazily initialised per-device values of one translation unit (occupancy-derived grids,
// shared-memory opt-ins): init(T&) -> bool runs once per device under the mutex; a failed
// init leaves nothing behind and is retried by the next call.
template <typename T>
struct StPerDevice {
    T val[64];
    bool ready[64] = {};
    std::mutex mu;
    template <typename Init>
    T *get(Init init) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
        std::lock_guard<std::mutex> lock(mu);
        if (!ready[dev]) {
            T t{};
            if (!init(t)) return nullptr;
            val[dev] = t;
            ready[dev] = true;
        }
        return &val[dev];
    }
};

struct StDeviceGuard {
    int prev, want;
    explicit StDeviceGuard(int device) : prev(-1), want(device) {
        cudaGetDevice(&prev);
        if (prev != want) cudaSetDevice(want);
    }
    ~StDeviceGuard() { if (prev != want && prev >= 0) cudaSetDevice(prev); }
};
