// One lockstep frame of the whole path in a single C call: the launches of
// K1 -> K2 -> K3 -> K4/K5 -> deflate -> K3 in order on one stream (no host work
// between them, no synchronisation).  Mirrors OCSORT_Disparity.predict between
// the detector backbone and the return value
// (mmtrack/models/mot/ocsort_disparity.py:79-109).
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
                       io->out_scales, io->trk_depth, io->out_ids, io->out_counts, stream);
    if (rc != ST_OK) return rc;
    rc = st_scale_boxes(io->trk_boxes, io->out_scales, io->out_counts, S, cap, 1, io->out_boxes, stream);
    if (rc != ST_OK) return rc;
    return st_box_depth(io->depth_map, S, io->Hp, io->Wp, io->out_boxes, io->out_counts, cap, io->out_depth,
                        io->out_scale2, nullptr, io->k3_workspace, stream);
}


// ---------------------------------------------------------------------------------
// Software pipeline over consecutive frames.  Only the tracker state couples frame
// f+1 to frame f; everything before the tracker (K1, K2, K3 on the detections) and
// after it (deflate, K3 on the tracks) depends on that frame's inputs alone.  The
// latency-bound stages (one CTA per sequence or per box, far fewer CTAs than SMs)
// therefore overlap with the bandwidth-bound K1 of later frames instead of leaving
// the GPU idle:
//   stream A: K1 -> K2 -> K3(dets)   of frame f+2
//   stream B: K4/K5 (tracker)        of frame f+1
//   stream C: deflate -> K3(tracks)  of frame f
// Results are identical to st_frame_step: the same kernels run on the same data in
// the same per-frame order; the ring of `depth` caller-owned buffer sets keeps the
// intermediates of the frames in flight apart.
struct st_pipeline {
    st_handle *h;
    int depth;
    long long submitted;
    cudaStream_t sa, sb, sc;
    cudaEvent_t *ev_a, *ev_b, *ev_c;   // [depth] stage completion of the frame in that slot
    cudaEvent_t fork_ev;
};

extern "C" int st_pipeline_create(st_handle *h, int depth, st_pipeline **out) {
    if (!h || !out || depth < 1 || depth > 64) return ST_ERR_INVALID_ARG;
    st_pipeline *p = new (std::nothrow) st_pipeline();
    if (!p) return ST_ERR_ALLOC;
    p->h = h; p->depth = depth; p->submitted = 0;
    p->ev_a = new (std::nothrow) cudaEvent_t[3 * depth];
    if (!p->ev_a) { delete p; return ST_ERR_ALLOC; }
    p->ev_b = p->ev_a + depth; p->ev_c = p->ev_b + depth;
    bool ok = cudaStreamCreateWithFlags(&p->sa, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&p->sb, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&p->sc, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&p->fork_ev, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < 3 * depth; ++i)
        ok = cudaEventCreateWithFlags(&p->ev_a[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { delete[] p->ev_a; delete p; return ST_ERR_CUDA; }
    *out = p;
    return ST_OK;
}

extern "C" int st_pipeline_destroy(st_pipeline *p) {
    if (!p) return ST_ERR_INVALID_ARG;
    cudaStreamSynchronize(p->sa); cudaStreamSynchronize(p->sb); cudaStreamSynchronize(p->sc);
    for (int i = 0; i < 3 * p->depth; ++i) cudaEventDestroy(p->ev_a[i]);
    cudaEventDestroy(p->fork_ev);
    cudaStreamDestroy(p->sa); cudaStreamDestroy(p->sb); cudaStreamDestroy(p->sc);
    delete[] p->ev_a;
    delete p;
    return ST_OK;
}

extern "C" int st_pipeline_fork(st_pipeline *p, void *stream) {
    if (!p) return ST_ERR_INVALID_ARG;
    ST_CUDA_OK(cudaEventRecord(p->fork_ev, (cudaStream_t)stream));
    ST_CUDA_OK(cudaStreamWaitEvent(p->sa, p->fork_ev, 0));
    ST_CUDA_OK(cudaStreamWaitEvent(p->sb, p->fork_ev, 0));
    ST_CUDA_OK(cudaStreamWaitEvent(p->sc, p->fork_ev, 0));
    return ST_OK;
}

extern "C" int st_pipeline_join(st_pipeline *p, void *stream) {
    if (!p) return ST_ERR_INVALID_ARG;
    cudaStream_t all[3] = {p->sa, p->sb, p->sc};
    for (int i = 0; i < 3; ++i) {
        ST_CUDA_OK(cudaEventRecord(p->fork_ev, all[i]));
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
    int rc;
    // ---- stage A: the buffer set of this slot is free once frame (k - depth) left stage C
    if (p->submitted >= p->depth) ST_CUDA_OK(cudaStreamWaitEvent(p->sa, p->ev_c[slot], 0));
    if (inputs_ready) ST_CUDA_OK(cudaStreamWaitEvent(p->sa, (cudaEvent_t)inputs_ready, 0));
    if (io->raw) {
        rc = st_disp2depth(io->raw, S, io->H, io->W, io->Hr, io->Wr, io->Hp, io->Wp, io->depth_map,
                           io->disp3, io->mask, h->cfg.baseline, h->cfg.focal_length, p->sa);
        if (rc != ST_OK) return rc;
    }
    if (io->cls) {
        rc = st_decode_nms(h, io->cls, io->bbox, io->obj, io->num_levels, io->level_h, io->level_w,
                           io->strides, S, io->num_classes, io->scale_w, io->scale_h, io->ori_h, io->ori_w,
                           cap, io->det_boxes, io->det_scores, io->det_labels, io->det_prior, io->det_keep,
                           io->det_counts, p->sa);
        if (rc != ST_OK) return rc;
    }
    rc = st_box_depth(io->depth_map, S, io->Hp, io->Wp, io->det_boxes, io->det_counts, cap, io->det_depth,
                      io->det_scales, io->det_inflated, io->k3_workspace, p->sa);
    if (rc != ST_OK) return rc;
    ST_CUDA_OK(cudaEventRecord(p->ev_a[slot], p->sa));
    // ---- stage B: the tracker, frames strictly in order on one stream
    ST_CUDA_OK(cudaStreamWaitEvent(p->sb, p->ev_a[slot], 0));
    rc = st_track_step(h, S, io->frame_ids, io->det_inflated, io->det_scores, io->det_labels, io->det_scales,
                       io->det_depth, io->det_counts, cap, io->trk_boxes, io->out_scores, io->out_labels,
                       io->out_scales, io->trk_depth, io->out_ids, io->out_counts, p->sb);
    if (rc != ST_OK) return rc;
    ST_CUDA_OK(cudaEventRecord(p->ev_b[slot], p->sb));
    // ---- stage C: deflate and the output depth
    ST_CUDA_OK(cudaStreamWaitEvent(p->sc, p->ev_b[slot], 0));
    rc = st_scale_boxes(io->trk_boxes, io->out_scales, io->out_counts, S, cap, 1, io->out_boxes, p->sc);
    if (rc != ST_OK) return rc;
    rc = st_box_depth(io->depth_map, S, io->Hp, io->Wp, io->out_boxes, io->out_counts, cap, io->out_depth,
                      io->out_scale2, nullptr, io->k3_workspace, p->sc);
    if (rc != ST_OK) return rc;
    ST_CUDA_OK(cudaEventRecord(p->ev_c[slot], p->sc));
    if (frame_done) ST_CUDA_OK(cudaEventRecord((cudaEvent_t)frame_done, p->sc));
    p->submitted += 1;
    return ST_OK;
}
