// One lockstep frame of the whole path in a single C call: the launches of
// K1 -> K2 -> K3 -> K4/K5 -> deflate -> K3 in order on one stream (no host work
// between them, no synchronisation).  Mirrors OCSORT_Disparity.predict between
// the detector backbone and the return value
// (mmtrack/models/mot/ocsort_disparity.py:79-109).
#include <cstdlib>
#include <new>
#include "st_common.cuh"

extern "C" int st_frame_step(st_handle *h, int S, const st_frame_io *io, void *stream) {
    if (!h || !io || S < 0 || S > h->cfg.max_sequences) return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
    const int cap = h->cfg.max_dets;
    int rc;
    if (io->raw) {
        rc = st_disp2depth(io->raw, S, io->H, io->W, io->Hr, io->Wr, io->Hp, io->Wp, io->depth_map,
                           io->disp3, io->mask, h->cfg.baseline, h->cfg.focal_length, stream);
        if (rc != ST_OK) return rc;
    }
    if (io->cls) {
        rc = st_decode_nms(h, io->cls, io->bbox, io->obj, io->num_levels, io->level_h, io->level_w,
                           io->strides, S, io->num_classes, io->scale_w, io->scale_h, io->ori_h, io->ori_w,
                           cap, io->det_boxes, io->det_scores, io->det_labels, io->det_prior, io->det_keep,
                           io->det_counts, stream);
        if (rc != ST_OK) return rc;
    }
    rc = st_box_depth(io->depth_map, S, io->Hp, io->Wp, io->det_boxes, io->det_counts, cap, io->det_depth,
                      io->det_scales, io->det_inflated, io->k3_workspace, stream);
    if (rc != ST_OK) return rc;
    rc = st_track_step(h, S, io->frame_ids, io->det_inflated, io->det_scores, io->det_labels, io->det_scales,
                       io->det_depth, io->det_counts, cap, io->trk_boxes, io->out_scores, io->out_labels,
                       io->out_scales, io->trk_depth, io->out_ids, io->out_counts, io->out_src, io->warp, stream);
    if (rc != ST_OK) return rc;
    rc = st_scale_boxes(io->trk_boxes, io->out_scales, io->out_counts, S, cap, 1, io->out_boxes, stream);
    if (rc != ST_OK) return rc;
    // rows whose truncated corners equal those of their detection take its statistic (st_box_depth_reuse)
    return st_box_depth_reuse(io->depth_map, S, io->Hp, io->Wp, io->out_boxes, io->out_counts, cap, io->out_depth,
                              io->out_scale2, nullptr, io->k3_workspace, io->out_src ? io->det_boxes : nullptr,
                              io->out_src, io->out_src ? io->det_depth : nullptr, stream);
}


// ---------------------------------------------------------------------------------
// Software pipeline over consecutive frames.  Only the tracker state couples frame
// f+1 to frame f; everything before the tracker (K1, K2, K3 on the detections) and
// after it (deflate, K3 on the tracks) depends on that frame's inputs alone, and K1
// and K2 do not depend on each other.  The latency-bound stages (one CTA per
// sequence or per box, far fewer CTAs than SMs) therefore overlap with each other
// and with the bandwidth-bound K1 of later frames instead of leaving the GPU idle.
// Five streams, each processing the frames in order:
//   stream 0: K1 (disparity -> depth)          stream 1: K2 (decode + NMS)
//   stream 2: K3 on the detections (needs 0 and 1)
//   stream 3: K4/K5, the tracker (needs 2; the only stage with cross-frame state)
//   stream 4: deflate + K3 on the tracks (needs 3 and the frame's depth map)
// Results are identical to st_frame_step: the same kernels run on the same data in
// the same per-frame order; the ring of `depth` caller-owned buffer sets keeps the
// intermediates of the frames in flight apart (a set is reused when the frame that
// last used it has left stream 4).
constexpr int kStages = 5;

struct st_pipeline {
    st_handle *h;
    int depth;
    long long submitted;
    cudaStream_t st[kStages];
    cudaEvent_t *ev;                   // [kStages][depth] stage completion of the frame in that slot
    cudaEvent_t fork_ev;
    // optional timeline (st_pipeline_trace): begin / end timing events of every stage of the
    // first trace_cap frames submitted after tracing was switched on
    cudaEvent_t *trace_ev;             // [trace_cap][kStages][2]
    cudaEvent_t trace_origin;
    int trace_cap, trace_n;
};

extern "C" int st_pipeline_create(st_handle *h, int depth, st_pipeline **out) {
    if (!h || !out || depth < 1 || depth > 64) return ST_ERR_INVALID_ARG;
    st_pipeline *p = new (std::nothrow) st_pipeline();
    if (!p) return ST_ERR_ALLOC;
    p->h = h; p->depth = depth; p->submitted = 0;
    p->trace_ev = nullptr; p->trace_cap = 0; p->trace_n = 0;
    p->ev = new (std::nothrow) cudaEvent_t[kStages * depth];
    if (!p->ev) { delete p; return ST_ERR_ALLOC; }
    bool ok = cudaEventCreateWithFlags(&p->fork_ev, cudaEventDisableTiming) == cudaSuccess;
    // Stream priorities: K1 (bandwidth-bound, fills every SM for its whole duration) on the lowest
    // priority, the latency-bound stages on the highest, so that their CTAs are placed first
    // whenever slots free up (+5 % frames/s at the bench workload; ST_PIPE_PRIO=0 switches it
    // off, 1 raises only the tracker: measured -12 %).
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const char *pe = getenv("ST_PIPE_PRIO");
    const int prio_mode = pe ? atoi(pe) : 2;
    for (int i = 0; ok && i < kStages; ++i) {
        int prio = 0;
        if (prio_mode == 1) prio = i == 3 ? prio_hi : prio_lo;
        if (prio_mode == 2) prio = i == 0 ? prio_lo : prio_hi;
        ok = cudaStreamCreateWithPriority(&p->st[i], cudaStreamNonBlocking, prio) == cudaSuccess;
    }
    for (int i = 0; ok && i < kStages * depth; ++i)
        ok = cudaEventCreateWithFlags(&p->ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { delete[] p->ev; delete p; return ST_ERR_CUDA; }
    *out = p;
    return ST_OK;
}

extern "C" int st_pipeline_destroy(st_pipeline *p) {
    if (!p) return ST_ERR_INVALID_ARG;
    for (int i = 0; i < kStages; ++i) cudaStreamSynchronize(p->st[i]);
    for (int i = 0; i < kStages * p->depth; ++i) cudaEventDestroy(p->ev[i]);
    cudaEventDestroy(p->fork_ev);
    if (p->trace_ev) {
        for (int i = 0; i < p->trace_cap * kStages * 2; ++i) cudaEventDestroy(p->trace_ev[i]);
        cudaEventDestroy(p->trace_origin);
        delete[] p->trace_ev;
    }
    for (int i = 0; i < kStages; ++i) cudaStreamDestroy(p->st[i]);
    delete[] p->ev;
    delete p;
    return ST_OK;
}

extern "C" int st_pipeline_fork(st_pipeline *p, void *stream) {
    if (!p) return ST_ERR_INVALID_ARG;
    ST_CUDA_OK(cudaEventRecord(p->fork_ev, (cudaStream_t)stream));
    for (int i = 0; i < kStages; ++i) ST_CUDA_OK(cudaStreamWaitEvent(p->st[i], p->fork_ev, 0));
    return ST_OK;
}

extern "C" int st_pipeline_join(st_pipeline *p, void *stream) {
    if (!p) return ST_ERR_INVALID_ARG;
    for (int i = 0; i < kStages; ++i) {
        ST_CUDA_OK(cudaEventRecord(p->fork_ev, p->st[i]));
        ST_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, p->fork_ev, 0));
    }
    return ST_OK;
}

extern "C" int st_pipeline_submit(st_pipeline *p, int S, const st_frame_io *io, void *inputs_ready,
                                  void *frame_done) {
    if (!p || !io || S < 0 || S > p->h->cfg.max_sequences) return ST_ERR_INVALID_ARG;
    if (S == 0) return ST_OK;
    st_handle *h = p->h;
    const int cap = h->cfg.max_dets;
    const int slot = (int)(p->submitted % p->depth);
    auto ev = [&](int stage) { return p->ev[stage * p->depth + slot]; };
    const bool reused = p->submitted >= p->depth;
    const bool traced = p->trace_ev && p->trace_n < p->trace_cap;
    auto mark = [&](int stage, int end) {            // timeline event on the stage's stream
        if (traced) cudaEventRecord(p->trace_ev[(p->trace_n * kStages + stage) * 2 + end], p->st[stage]);
    };
    int rc;
    // ---- stream 0: K1.  The buffer set of this slot is free once its previous frame left stream 4.
    if (reused) ST_CUDA_OK(cudaStreamWaitEvent(p->st[0], ev(4), 0));
    if (inputs_ready) ST_CUDA_OK(cudaStreamWaitEvent(p->st[0], (cudaEvent_t)inputs_ready, 0));
    // diagnostics only (ST_PIPE_SKIP, bit i = leave stage i's kernels out; results are then stale):
    // shows which stage slows which in the overlapped regime (tools/pipeline_timeline.py)
    static int skip = -1;
    if (skip < 0) { const char *e = getenv("ST_PIPE_SKIP"); skip = e ? atoi(e) : 0; }
    mark(0, 0);
    if (io->raw && !(skip & 1)) {
        rc = st_disp2depth(io->raw, S, io->H, io->W, io->Hr, io->Wr, io->Hp, io->Wp, io->depth_map,
                           io->disp3, io->mask, h->cfg.baseline, h->cfg.focal_length, p->st[0]);
        if (rc != ST_OK) return rc;
    }
    mark(0, 1);
    ST_CUDA_OK(cudaEventRecord(ev(0), p->st[0]));
    // ---- stream 1: K2 (its det_* outputs were last read by stream 3 and 4 of the slot's previous frame)
    if (reused) ST_CUDA_OK(cudaStreamWaitEvent(p->st[1], ev(4), 0));
    if (inputs_ready) ST_CUDA_OK(cudaStreamWaitEvent(p->st[1], (cudaEvent_t)inputs_ready, 0));
    mark(1, 0);
    if (io->cls && !(skip & 2)) {
        rc = st_decode_nms(h, io->cls, io->bbox, io->obj, io->num_levels, io->level_h, io->level_w,
                           io->strides, S, io->num_classes, io->scale_w, io->scale_h, io->ori_h, io->ori_w,
                           cap, io->det_boxes, io->det_scores, io->det_labels, io->det_prior, io->det_keep,
                           io->det_counts, p->st[1]);
        if (rc != ST_OK) return rc;
    }
    mark(1, 1);
    ST_CUDA_OK(cudaEventRecord(ev(1), p->st[1]));
    // ---- stream 2: K3 on the detections
    ST_CUDA_OK(cudaStreamWaitEvent(p->st[2], ev(0), 0));
    ST_CUDA_OK(cudaStreamWaitEvent(p->st[2], ev(1), 0));
    mark(2, 0);
    rc = (skip & 4) ? ST_OK
                    : st_box_depth(io->depth_map, S, io->Hp, io->Wp, io->det_boxes, io->det_counts, cap, io->det_depth,
                                   io->det_scales, io->det_inflated, io->k3_workspace, p->st[2]);
    if (rc != ST_OK) return rc;
    mark(2, 1);
    ST_CUDA_OK(cudaEventRecord(ev(2), p->st[2]));
    // ---- stream 3: the tracker, frames strictly in order
    ST_CUDA_OK(cudaStreamWaitEvent(p->st[3], ev(2), 0));
    mark(3, 0);
    rc = st_track_step(h, S, io->frame_ids, io->det_inflated, io->det_scores, io->det_labels, io->det_scales,
                       io->det_depth, io->det_counts, cap, io->trk_boxes, io->out_scores, io->out_labels,
                       io->out_scales, io->trk_depth, io->out_ids, io->out_counts, io->out_src, io->warp, p->st[3]);
    if (rc != ST_OK) return rc;
    mark(3, 1);
    ST_CUDA_OK(cudaEventRecord(ev(3), p->st[3]));
    // ---- stream 4: deflate and the output depth
    ST_CUDA_OK(cudaStreamWaitEvent(p->st[4], ev(3), 0));
    mark(4, 0);
    rc = st_scale_boxes(io->trk_boxes, io->out_scales, io->out_counts, S, cap, 1, io->out_boxes, p->st[4]);
    if (rc != ST_OK) return rc;
    rc = st_box_depth_reuse(io->depth_map, S, io->Hp, io->Wp, io->out_boxes, io->out_counts, cap, io->out_depth,
                            io->out_scale2, nullptr, io->k3_workspace, io->out_src ? io->det_boxes : nullptr,
                            io->out_src, io->out_src ? io->det_depth : nullptr, p->st[4]);
    if (rc != ST_OK) return rc;
    mark(4, 1);
    if (traced) p->trace_n += 1;
    ST_CUDA_OK(cudaEventRecord(ev(4), p->st[4]));
    if (frame_done) ST_CUDA_OK(cudaEventRecord((cudaEvent_t)frame_done, p->st[4]));
    p->submitted += 1;
    return ST_OK;
}

// Timeline of the pipeline (tracing aid): after st_pipeline_trace(p, n) the begin and end
// of every stage of the next n submitted frames are time-stamped with CUDA events on the
// stage's own stream; st_pipeline_read_trace synchronises and returns, per traced frame
// and stage, {begin, end} in milliseconds since tracing was switched on.
extern "C" int st_pipeline_trace(st_pipeline *p, int n_frames) {
    if (!p || n_frames < 0 || n_frames > 4096) return ST_ERR_INVALID_ARG;
    if (p->trace_ev) {
        for (int i = 0; i < p->trace_cap * kStages * 2; ++i) cudaEventDestroy(p->trace_ev[i]);
        cudaEventDestroy(p->trace_origin);
        delete[] p->trace_ev;
        p->trace_ev = nullptr;
    }
    p->trace_cap = n_frames; p->trace_n = 0;
    if (n_frames == 0) return ST_OK;
    p->trace_ev = new (std::nothrow) cudaEvent_t[(size_t)n_frames * kStages * 2];
    if (!p->trace_ev) { p->trace_cap = 0; return ST_ERR_ALLOC; }
    for (int i = 0; i < n_frames * kStages * 2; ++i) ST_CUDA_OK(cudaEventCreate(&p->trace_ev[i]));
    ST_CUDA_OK(cudaEventCreate(&p->trace_origin));
    ST_CUDA_OK(cudaEventRecord(p->trace_origin, p->st[0]));
    return ST_OK;
}

extern "C" int st_pipeline_read_trace(st_pipeline *p, float *ms_host, int max_frames) {
    if (!p || !ms_host || max_frames < 0) return ST_ERR_INVALID_ARG;
    if (!p->trace_ev) return 0;
    for (int i = 0; i < kStages; ++i) ST_CUDA_OK(cudaStreamSynchronize(p->st[i]));
    const int n = p->trace_n < max_frames ? p->trace_n : max_frames;
    for (int i = 0; i < n * kStages * 2; ++i)
        ST_CUDA_OK(cudaEventElapsedTime(&ms_host[i], p->trace_origin, p->trace_ev[i]));
    return n;
}
