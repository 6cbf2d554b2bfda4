// One lockstep frame of the whole path in a single C call: the launches of
// K1 -> K2 -> K3 -> K4/K5 -> deflate -> K3 in order on one stream (no host work
// between them, no synchronisation).  Mirrors OCSORT_Disparity.predict between
// the detector backbone and the return value
// (mmtrack/models/mot/ocsort_disparity.py:79-109).
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
