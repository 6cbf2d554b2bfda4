/* stereotrack_b200.h - C ABI of the B200-native depth-guided tracking path.
 *
 * Drop-in boundary for the per-frame path of Superjie13/StereoTracking that
 * runs after the detector backbone and around the disparity input
 * (mmtrack/models/mot/ocsort_disparity.py:82-109 and everything it calls).
 * The reference is pure Python, so there is no FFI in the reference to bind;
 * these entry points are what the Python boundary classes in
 * stereotracking_b200/ (same registry names and call signatures as the
 * reference) call through ctypes.  INTEGRATION.md shows the reference-side
 * binding.
 *
 * Conventions
 *  - every data pointer is a DEVICE pointer into caller-owned memory unless the
 *    parameter name ends in _host;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *    never allocates or frees caller memory and never synchronises (st_create /
 *    st_destroy / st_export_state / st_read_* / st_snapshot_* excepted).  The per-frame
 *    calls consist of kernel launches, memsets and event fork/join on streams the
 *    library owns, i.e. the stream-capturable subset; st_pipeline_* can replay a frame
 *    as a CUDA graph (st_pipeline_use_graphs);
 *  - a handle belongs to the CUDA device that was current at st_create; every call that
 *    takes a handle makes that device current for its duration.  Lazily initialised
 *    library state (occupancy-derived grids, auxiliary streams, events) is kept per
 *    device and is safe to touch from several host threads;
 *  - return value: 0 = ok, < 0 = error (st_error_string); nothing throws;
 *  - S = sequences processed in lockstep (one frame each per call); per-sequence
 *    arrays are padded to a fixed capacity with a separate count array;
 *  - overflow of a capacity never truncates silently: the per-sequence bit is
 *    raised in the handle's status words (st_read_status).
 */
#ifndef STEREOTRACK_B200_H
#define STEREOTRACK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ST_ABI_VERSION 7

enum {
    ST_OK = 0,
    ST_ERR_INVALID_ARG = -1,
    ST_ERR_CUDA = -2,
    ST_ERR_NO_DEVICE = -3,
    ST_ERR_CAPACITY = -4,
    ST_ERR_ALLOC = -5
};

/* per-sequence status bits written by the kernels (sticky until st_clear_status) */
enum {
    ST_STATUS_CAND_OVERFLOW = 1,   /* more post-threshold candidates than max_candidates */
    ST_STATUS_DET_OVERFLOW = 2,    /* more detections than the tracker's max_dets        */
    ST_STATUS_TRACK_OVERFLOW = 4,  /* more live tracks than max_tracks                   */
    ST_STATUS_LAP_GUARD = 8        /* assignment solver hit its iteration guard          */
};

/* Tracker / model configuration.  Field names and defaults follow the
 * reference constructors: OCSORTTracker_Disparity
 * (mmtrack/models/trackers/ocsort_tracker_disparity.py:42-52), KalmanTrackerBase
 * (kalman_tracker_base.py:33-47), BaseTracker (base_tracker.py:22-32),
 * OCSORT_Disparity (mmtrack/models/mot/ocsort_disparity.py:30-40) and the
 * detector test_cfg (configs/stereo_tracking/ocsort/
 * yolox_s_mmyolo_mot_airdrone_disp.py:42). */
typedef struct st_config {
    int32_t max_sequences;      /* S capacity of the handle                    */
    int32_t max_tracks;         /* live tracks per sequence (<= 1024)          */
    int32_t max_dets;           /* detections per sequence per frame (<= 1024) */
    int32_t max_candidates;     /* post-threshold candidates per frame         */
    float obj_score_thr;        /* 0.3  */
    float init_track_thr;       /* 0.7  */
    int32_t weight_iou_with_det_scores; /* 0/1 */
    double match_iou_thr;       /* 0.1 in the shipped config (ctor default 0.3); double:
                                 * cost_limit = 1 - match_iou_thr is Python float math */
    int32_t num_tentatives;     /* 3    */
    float vel_consist_weight;   /* 0.2  */
    int32_t vel_delta_t;        /* 3 (<= 15) */
    int32_t num_frames_retain;  /* 30 in the shipped config (ctor default 10)   */
    float baseline;             /* 0.25 m */
    float focal_length;         /* 640 px */
    float score_thr;            /* 0.01 */
    float nms_iou_thr;          /* 0.5  */
} st_config;

typedef struct st_handle st_handle;   /* opaque: owns the SoA track state of S x max_tracks slots */

int st_abi_version(void);
const char *st_error_string(int code);
void st_default_config(st_config *cfg);

/* Replaces the Python object state of OCSORTTracker_Disparity
 * (base_tracker.py:34-37 `reset`, self.tracks / self.num_tracks). */
int st_create(const st_config *cfg, st_handle **out);
int st_destroy(st_handle *h);
/* reset every sequence whose byte in seq_mask is non-zero (NULL = all) */
int st_reset(st_handle *h, const uint8_t *seq_mask, void *stream);
int st_read_status(st_handle *h, int32_t *status_host, int S);   /* waits for ALL work on the device, then reads */
/* number of live tracks of every sequence (len(self.tracks), base_tracker.py:39-41 `empty`):
 * one copy on `stream`, which is synchronised */
int st_read_live_counts(st_handle *h, int32_t *live_host, int S, void *stream);
int st_clear_status(st_handle *h, void *stream);
/* Diagnostics: 32 words per sequence written by st_track_step: 0-2 stage / rows / columns of
 * the last association, 3-4 active rows and rows*1000+columns of the last stage-1 fast solve,
 * 5 assignments solved, 6 solved by the exact fallback, 7 phase, 8-13 SM cycles of the last
 * step per phase (filter+predict, stage 1, 2, 3, smoothing+ordering, update), 14-15 cycles of
 * the search and of the uniqueness certificate of the last stage-1 solve, 16-19 cycles of the
 * certificate's edge scan / reach rounds / cycle rounds and its arc count.  Copied on a side
 * stream, so it can be called while a step is still running. */
int st_read_debug(st_handle *h, int32_t *debug_host, int S);
/* Diagnostics: disable parts of st_track_step (bit0..2 assignment of stage 1..3,
 * bit3 online smoothing, bit4 Kalman update, bit5 Kalman predict). 0 = normal. */
int st_debug_set_skip(int mask);
/* Diagnostics: the cost matrix of the last association of sequence `seq` (row-major R x C). */
int st_debug_read_cost(st_handle *h, int seq, float *cost_host, int count);

/* K1 - replaces loading_disparity.py:82-86,:129-134 (mask, 65535->0, /16),
 * transforms_disparity.py:52-70 (nearest resize to Hr x Wr), :177-194 (zero pad
 * to Hp x Wp) and ocsort_disparity.py:132-134 (Z = B*f/(d+1e-6)).
 * raw [S,H,W] u16 -> depth [S,Hp,Wp] f32; optional disp3 [S,3,Hp,Wp] f32 (the
 * detector's disp_postp input) and mask [S,Hp,Wp] u8. */
int st_disp2depth(const uint16_t *raw, int S, int H, int W, int Hr, int Wr, int Hp, int Wp,
                  float *depth, float *disp3, uint8_t *mask,
                  float baseline, float focal_length, void *stream);

/* K1b - the RGB half of the detector input from the decoded u8 frame, for the test pipeline of
 * configs/stereo_tracking/ocsort/yolox_s_mmyolo_mot_airdrone_disp.py:104-116 at resize factor 1:
 * Pad_Disparity._pad_img (transforms_disparity.py:243; bottom/right pad of the u8 image with
 * pad_val['img']), PackTrackInputs_Disparity HWC -> CHW (formatting_disparity.py:166-176) and
 * TrackDataPreprocessor_Disparity_V1.forward (data_preprocessor_disparity_v1.py:38-52: optional
 * channel swap, .float(), optional (img - mean) / std, stack_batch to the padded size).
 * img [S,H,W,3] u8 (device) -> out [S,3,Hp,Wp] f32 (device).  pad_val_host [3] u8 in INPUT channel
 * order (NULL = 0); mean_host / std_host [3] f32 in OUTPUT channel order, both or neither
 * (NULL = no normalisation, as in the shipped config); swap_rb != 0 = bgr_to_rgb / rgb_to_bgr. */
int st_pack_image(const uint8_t *img, int S, int H, int W, int Hp, int Wp,
                  const uint8_t *pad_val_host, const float *mean_host, const float *std_host,
                  int swap_rb, float *out, void *stream);
/* The same with the image branch of Resize_Disparity in front, for a resize factor != 1
 * (transforms_disparity.py:128 -> mmdet Resize._resize_img -> mmcv.imrescale, 'bilinear', cv2 backend =
 * cv2.resize INTER_LINEAR on the u8 image; a downscale by exactly 2 is OpenCV's 2 x 2 box average):
 * img [S,H,W,3] u8 -> bilinear to Hr x Wr (bit-exact, 11-bit fixed-point taps) -> pad to Hp x Wp ->
 * out [S,3,Hp,Wp] f32.  Hr == H and Wr == W is exactly the entry above. */
int st_resize_pack_image(const uint8_t *img, int S, int H, int W, int Hr, int Wr, int Hp, int Wp,
                         const uint8_t *pad_val_host, const float *mean_host, const float *std_host,
                         int swap_rb, float *out, void *stream);

/* disp2depth of OCSORT_Disparity (ocsort_disparity.py:132-134) on an already
 * post-processed float disparity (`disp_postp[:, 0:1]`), n contiguous elements. */
int st_disp_f32_to_depth(const float *disp, long long n, float *depth, float baseline,
                         float focal_length, void *stream);

/* K2 - replaces the detector-head post-process invoked at
 * mmtrack/models/detectors/yolo_detector_disparity_v1.py:121-122 (mmyolo
 * YOLOv5Head.predict_by_feat + YOLOXBBoxCoder.decode, mmdet
 * filter_scores_and_topk, mmcv batched_nms).  Level l has maps cls[l]
 * [S,num_classes,h_l,w_l], bbox[l] [S,4,h_l,w_l], obj[l] [S,1,h_l,w_l] (NCHW f32).
 * Outputs are padded to det_cap rows per sequence, in descending score order:
 * boxes [S,det_cap,4], scores [S,det_cap], labels [S,det_cap] i64, prior_idx
 * [S,det_cap] i32 (flattened prior index), keep_idx [S,det_cap] i32 (position
 * in the score-sorted candidate list, what mmcv nms returns), counts [S] i32. */
int st_decode_nms(st_handle *h, const float *const *cls_host, const float *const *bbox_host,
                  const float *const *obj_host, int num_levels, const int32_t *level_h_host,
                  const int32_t *level_w_host, const int32_t *strides_host, int S, int num_classes,
                  float scale_w, float scale_h, int ori_h, int ori_w, int det_cap,
                  float *boxes, float *scores, int64_t *labels, int32_t *prior_idx,
                  int32_t *keep_idx, int32_t *counts, void *stream);

/* K3 - replaces OCSORT_Disparity.extract_depth + bbox_postp_depth + scale_bbox
 * (ocsort_disparity.py:113-130,:136-175; trackers/utils.py:58-73).
 * depth [S,Hp,Wp]; boxes [S,cap,4] xyxy; counts [S].  Outputs d [S,cap] (value,
 * -1, or NaN exactly as the reference), scale [S,cap], inflated [S,cap,4]
 * (may be NULL). */
int st_box_depth(const float *depth, int S, int Hp, int Wp, const float *boxes,
                 const int32_t *counts, int cap, float *d, float *scale, float *inflated,
                 int32_t *workspace /* [8 + 6*S*cap] i32 scratch (worklists of large boxes) */,
                 void *stream);
/* Same, for the second call of a frame (ocsort_disparity.py:100-101, on the deflated
 * track boxes): row r of `boxes` came from row prev_index[r] of prev_boxes (-1 = none),
 * whose statistic prev_d was computed by an earlier call on the same depth map.  The
 * statistic depends on a box only through its truncated corners (:141), so rows whose
 * truncated corners equal those of their source row take prev_d; all others are
 * computed.  Results are identical to st_box_depth. */
int st_box_depth_reuse(const float *depth, int S, int Hp, int Wp, const float *boxes,
                       const int32_t *counts, int cap, float *d, float *scale, float *inflated,
                       int32_t *workspace, const float *prev_boxes, const int32_t *prev_index,
                       const float *prev_d, void *stream);

/* K3 on the RAW u16 frame, no depth map in between: the same call as st_box_depth_reuse on the map
 * st_disp2depth(raw, S, H, W, H, W, Hp, Wp, ..., baseline, focal_length) would write (resize
 * factor 1, i.e. Hr == H and Wr == W as in every shipped config), with bit-identical results.
 * Replaces ocsort_disparity.py:132-134 + :136-175 for the boxes only: depth is monotone in the raw
 * code, so the order statistics are taken on 16-bit codes (shared-memory histograms) and only the
 * window sum evaluates Z = B f / (d + 1e-6).  prev_* as in st_box_depth_reuse (NULL = none). */
int st_box_depth_raw(const uint16_t *raw, int S, int H, int W, int Hp, int Wp, float baseline,
                     float focal_length, const float *boxes, const int32_t *counts, int cap,
                     float *d, float *scale, float *inflated, int32_t *workspace,
                     const float *prev_boxes, const int32_t *prev_index, const float *prev_d,
                     void *stream);

/* The ablation reducers of mmtrack/models/mot/depth_extraction_comparison.py (decorators that
 * replace extract_depth in comparison runs): same patch, validity filter and -1 rule, another
 * statistic, scale = clip(d*d/400, 1, 3).  ST_DEPTH_WINDOW is st_box_depth (scale = clip(d*d, 1, 3)).
 * ST_DEPTH_CENTER reads the unfiltered map at the integer box centre with NumPy indexing; where
 * the reference raises IndexError (:141, centre outside the map) d = NaN, scale = 1 and oob[s]
 * (may be NULL; [S] i32, zeroed by the call) counts such boxes of sequence s.  `workspace` is
 * only used by ST_DEPTH_WINDOW (may be NULL otherwise). */
enum st_depth_mode {
    ST_DEPTH_WINDOW = 0,          /* ocsort_disparity.py:136-175 */
    ST_DEPTH_TRUNCATED_MEAN = 1,  /* depth_extraction_comparison.py:5-46 */
    ST_DEPTH_MEAN = 2,            /* :49-80 */
    ST_DEPTH_MEDIAN = 3,          /* :83-115 */
    ST_DEPTH_CENTER = 4           /* :117-150 */
};
int st_box_depth_mode(const float *depth, int S, int Hp, int Wp, const float *boxes,
                      const int32_t *counts, int cap, int mode, float *d, float *scale,
                      float *inflated, int32_t *oob, int32_t *workspace, void *stream);

/* scale_bbox(boxes, 1/scales) (ocsort_disparity.py:96-97) or scale_bbox(boxes,
 * scales) when invert == 0. */
int st_scale_boxes(const float *boxes, const float *scales, const int32_t *counts, int S, int cap,
                   int invert, float *out, void *stream);

/* K4+K5 - replaces OCSORTTracker_Disparity.track
 * (ocsort_tracker_disparity.py:345-618) including KalmanFilter
 * predict/update (motion/kalman_filter.py:87-189), the three lap.lapjv
 * associations and BaseTracker.update / pop_invalid_tracks.
 * Inputs = pred_det_instances fields (boxes already inflated), padded to
 * det_cap rows, det_counts [S]; frame_ids [S] i32.  Outputs = the returned
 * InstanceData fields, padded to det_cap rows, out_counts [S]. */
int st_track_step(st_handle *h, int S, const int32_t *frame_ids, const float *det_boxes,
                  const float *det_scores, const int64_t *det_labels, const float *det_scales,
                  const float *det_depth, const int32_t *det_counts, int det_cap,
                  float *out_boxes, float *out_scores, int64_t *out_labels, float *out_scales,
                  float *out_depth, int64_t *out_ids, int32_t *out_counts,
                  int32_t *out_src /* [S,det_cap] input row of every output row, may be NULL */,
                  const float *warp /* [S,6] row-major 2x3 camera-motion affine per sequence applied to
                                       the confirmed tracks after predict (cmc=dict(method='glme_affine'),
                                       ocsort_tracker_disparity.py:428-445, trackers/gmc.py:20-45); a NaN
                                       first entry = no warp for that sequence; NULL = cmc disabled */,
                  void *stream);

/* All buffers of one lockstep frame (device pointers; host arrays where the name
 * says so).  Capacities: every [S,cap,...] array has cap = st_config.max_dets. */
typedef struct st_frame_io {
    /* inputs */
    const int32_t *frame_ids;          /* [S] */
    const uint16_t *raw;               /* [S,H,W] or NULL when depth_map is already filled */
    int32_t H, W, Hr, Wr, Hp, Wp;
    const float *const *cls;           /* host arrays of num_levels device pointers, or NULL */
    const float *const *bbox;          /*   when det_* already hold the detector output      */
    const float *const *obj;
    const int32_t *level_h, *level_w, *strides;   /* host */
    int32_t num_levels, num_classes, ori_h, ori_w;
    float scale_w, scale_h;
    /* intermediates */
    float *depth_map;                  /* [S,Hp,Wp]; may be NULL when depth_from_raw != 0 */
    float *disp3;                      /* [S,3,Hp,Wp] or NULL */
    uint8_t *mask;                     /* [S,Hp,Wp] or NULL */
    float *det_boxes; float *det_scores; int64_t *det_labels; int32_t *det_prior; int32_t *det_keep;
    int32_t *det_counts; float *det_depth; float *det_scales; float *det_inflated;
    float *trk_boxes; float *trk_depth; float *out_scale2;
    int32_t *out_src;                  /* [S,cap] detection row of every track row */
    int32_t *k3_workspace;             /* [8 + 6*S*cap] i32 scratch of st_box_depth */
    /* outputs = pred_track_instances fields */
    float *out_boxes; float *out_scores; int64_t *out_labels; float *out_scales; float *out_depth;
    int64_t *out_ids; int32_t *out_counts;
    const float *warp;                 /* [S,6] camera-motion affine of this frame or NULL (st_track_step) */
    int32_t depth_from_raw;            /* != 0: both K3 calls gather from `raw` (st_box_depth_raw; needs raw,
                                        * Hr == H, Wr == W); K1 then only runs when the caller wants the map
                                        * (depth_map != NULL), so the 4 Hp Wp byte write leaves the path */
} st_frame_io;

/* The whole per-frame path for S sequences in one call (K1 -> K2 -> K3 -> K4/K5
 * -> deflate -> K3); replaces ocsort_disparity.py:79-109 after the backbone. */
int st_frame_step(st_handle *h, int S, const st_frame_io *io, void *stream);

/* Software pipeline over consecutive frames of the same S sequences.  Only the
 * tracker state couples frame f+1 to frame f (ocsort_tracker_disparity.py:345-618);
 * K1, K2 and K3 on the detections of later frames and the deflate + K3 on the tracks
 * of earlier frames run concurrently with it on internal streams: a "pre" segment per
 * ring slot (K1 | K2 -> K3 on detections), the tracker on one stream in frame order, a
 * "post" segment (deflate + K3 on tracks).  Results equal st_frame_step.  The caller owns `depth` complete st_frame_io buffer sets and
 * passes set (k mod depth) with the k-th submit; inputs must stay valid until that
 * frame is done.  inputs_ready / frame_done are cudaEvent_t (may be NULL): the first
 * stage waits for inputs_ready, frame_done is recorded when every output of the
 * frame is complete. */
typedef struct st_pipeline st_pipeline;
int st_pipeline_create(st_handle *h, int depth, st_pipeline **out);
int st_pipeline_destroy(st_pipeline *p);
int st_pipeline_submit(st_pipeline *p, int S, const st_frame_io *io, void *inputs_ready, void *frame_done);
/* Replay frames as CUDA graphs: from a ring slot's second use on, its pre
 * and post segments are stream-captured once and relaunched with cudaGraphLaunch; the tracker
 * kernel stays a plain launch between them (it carries the cross-frame dependency in stream
 * order).  The frame's input pointers (raw, head maps) are read by the kernels from a per-slot
 * device table that follows st_frame_io, so replaying on new inputs needs no graph update; a
 * change of any other st_frame_io field - buffers, sizes - re-captures the slot.  Identical
 * results; st_pipeline_stats reports how many frames were replayed and how many captures ran. */
int st_pipeline_use_graphs(st_pipeline *p, int enable);
int st_pipeline_stats(st_pipeline *p, long long *graph_frames, long long *captures);
/* SM partition of the pipeline (environment ST_PIPE_PARTITION=1 at st_pipeline_create, while max_sequences is at
 * most ~55 % of the SMs): the tracker stream runs in a green context that owns tracker_sms SMs - one tracker CTA
 * per sequence, one SM each -, every other stream of the pipeline in a second one with the remaining SMs, so the
 * latency-bound tracker never shares an SM with the streaming stages of the neighbouring frames.  Both are 0
 * when the pipeline is not partitioned (switched off, driver without green contexts, too many sequences). */
int st_pipeline_partition(st_pipeline *p, int *tracker_sms, int *other_sms);
/* the pipeline's streams start after the work already queued on `stream` */
int st_pipeline_fork(st_pipeline *p, void *stream);
/* `stream` waits for every frame submitted so far */
int st_pipeline_join(st_pipeline *p, void *stream);
/* Timeline (tracing aid; the reference has no per-stage timers, utils/benchmark.py:195-231
 * times whole test_steps): time-stamp begin and end of the five stages of the next n_frames
 * submitted frames; read_trace synchronises and fills ms_host[frame][stage][begin,end] in
 * milliseconds since tracing was switched on, returning the number of traced frames. */
int st_pipeline_trace(st_pipeline *p, int n_frames);
int st_pipeline_read_trace(st_pipeline *p, float *ms_host, int max_frames);

/* Debug / parity / checkpoint: copies the live tracks of sequence `seq` in
 * dict order to HOST arrays sized max_tracks (synchronises).  mean [T,8] f64,
 * cov [T,64] f64, velocity [T,2] f32, flags [T] u8 (bit0 tentative, bit1
 * tracked), hits/last_frame [T] i32.  Returns the number of live tracks. */
int st_export_state(st_handle *h, int seq, int64_t *ids_host, double *mean_host, double *cov_host,
                    float *velocity_host, uint8_t *flags_host, int32_t *hits_host,
                    int32_t *last_frame_host, int64_t *num_tracks_host);

/* Checkpoint / resume of the whole tracker state of a handle (the reference never checkpoints
 * its tracker: it restarts at frame_id == 0, ocsort_tracker_disparity.py:385-387).  A snapshot
 * is a HOST buffer of st_snapshot_bytes(h) bytes; save and load synchronise `stream`; no step
 * may be in flight on another stream.  Loading needs a handle created with the same
 * max_sequences / max_tracks / max_dets / max_candidates (else ST_ERR_INVALID_ARG); the
 * tracker thresholds of the loading handle stay its own. */
long long st_snapshot_bytes(st_handle *h);
int st_snapshot_save(st_handle *h, void *buf_host, long long bytes, void *stream);
int st_snapshot_load(st_handle *h, const void *buf_host, long long bytes, void *stream);

/* Batched Kalman primitives on caller arrays (KalmanFilter.initiate/predict/
 * update of motion/kalman_filter.py:60-189), one warp per track: mean [n,8],
 * cov [n,64] f64 in place, z [n,4] f32 (cx,cy,a,h). op: 0 initiate, 1 predict,
 * 2 update. */
int st_kalman_batch(int op, int n, double *mean, double *cov, const float *z, void *stream);
/* KalmanFilter.project (kalman_filter.py:121-152, use_nsa=False): pmean [n,4], pcov [n,16] f64
 * (either may be NULL). */
int st_kalman_project(int n, const double *mean, const double *cov, double *pmean, double *pcov,
                      void *stream);
/* KalmanFilter(use_nsa=True) (kalman_filter.py:38,:144-145): project / update with every
 * measurement std scaled by noise_scale[i] = 1 - bbox_score of state i ([n] f64; NULL = 1, the
 * plain calls above).  The OC-SORT trackers never pass a bbox_score (kalman_tracker_base.py:74-75,
 * ocsort_tracker_disparity.py:342-343), so inside st_track_step use_nsa is the identity. */
int st_kalman_project_scored(int n, const double *mean, const double *cov, const double *noise_scale,
                             double *pmean, double *pcov, void *stream);
int st_kalman_batch_scored(int op, int n, double *mean, double *cov, const float *z,
                           const double *noise_scale, void *stream);
/* KalmanFilter.gating_distance (kalman_filter.py:190-234) of n states against the same m
 * measurements meas [m,4] f64 (cx,cy,a,h): dist [n,m] f64 squared Mahalanobis distances;
 * only_position != 0 uses (cx,cy) only.  A projected covariance that is not positive definite
 * gives NaN (NumPy raises LinAlgError). */
int st_kalman_gating(int n, const double *mean, const double *cov, int m, const double *meas,
                     int only_position, double *dist, void *stream);

/* Batched association (ocm/ocr cost is built by st_track_step; this entry
 * exposes the assignment itself for parity tests): cost [B,R,C] f32 ->
 * row2col [B,R], col2row [B,C] i32 (-1 = unmatched), lap.lapjv(extend_cost=True,
 * cost_limit) semantics.  Debug word 6 of sequence 0 counts exact-fallback solves.  Uses the handle's scratch; R,C <= max_tracks/max_dets. */
int st_lapjv_batch(st_handle *h, int B, int R, int C, const float *cost, double cost_limit,
                   int mode /* 0: certified fast path + exact fallback, 1: exact emulation only */,
                   int32_t *row2col, int32_t *col2row, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* STEREOTRACK_B200_H */
