"""Model shell of the hot path: one frame of OCSORT_Disparity.predict.

Follows mmtrack/models/mot/ocsort_disparity.py:51-111 between the detector
call (:79) and the return (:111): depth from disparity, per-box depth and
scale, inflate, tracker.track, deflate, depth recomputed on deflated boxes.
"""
import numpy as np

from . import boxdepth, disparity, head
from .tracker import OCSortDisparityOracle


class StereoTrackOracle:
    """One sequence.  ``tracker_cfg`` = config keys of
    configs/stereo_tracking/ocsort/yolox_s_mmyolo_mot_airdrone_disp.py:48-58."""

    DEFAULT_TRACKER = dict(obj_score_thr=0.3, init_track_thr=0.7,
                           weight_iou_with_det_scores=False, match_iou_thr=0.1,
                           num_tentatives=3, vel_consist_weight=0.2,
                           vel_delta_t=3, num_frames_retain=30)

    def __init__(self, tracker_cfg=None, baseline=0.25, focal_length=640,
                 img_scale=None, size_divisor=32, score_thr=0.01, nms_iou=0.5,
                 strides=(8, 16, 32), depth_reducer=None):
        # depth_reducer: None = extract_depth (:136-175); 'truncated_mean' | 'mean' | 'median' | 'center'
        # = extract_depth decorated as in mmtrack/models/mot/depth_extraction_comparison.py
        self.depth_reducer = depth_reducer
        self.tracker = OCSortDisparityOracle(**(self.DEFAULT_TRACKER if tracker_cfg is None else tracker_cfg))
        self.baseline, self.focal_length = baseline, focal_length
        self.img_scale, self.size_divisor = img_scale, size_divisor
        self.score_thr, self.nms_iou, self.strides = score_thr, nms_iou, strides

    def depth_map(self, raw_u16):
        return disparity.preprocess(raw_u16, self.img_scale, self.size_divisor,
                                    self.baseline, self.focal_length)

    def detect(self, cls_maps, bbox_maps, obj_maps, ori_hw):
        return head.postprocess(cls_maps, bbox_maps, obj_maps, ori_hw,
                                strides=self.strides, score_thr=self.score_thr,
                                iou_thr=self.nms_iou)

    def step_from_dets(self, frame_id, depth, det_bboxes, det_scores, det_labels, warp=None):
        """ocsort_disparity.py:82-109 given the detector output (``warp``: the frame's
        camera-motion affine when cmc is configured, else None)."""
        extract = boxdepth.extract_depth if self.depth_reducer is None else (
            lambda dm, bb: boxdepth.extract_depth_ablation(dm, bb, self.depth_reducer))
        d, scales = extract(depth, det_bboxes)
        infl = boxdepth.scale_bbox(det_bboxes, scales)                      # bbox_postp_depth (:113-130)
        trk = self.tracker.track(frame_id, infl, det_scores, det_labels, scales, d, warp=warp)
        with np.errstate(divide='ignore', invalid='ignore'):
            inv = (np.float32(1) / trk['scales']).astype(np.float32)
        out_boxes = boxdepth.scale_bbox(trk['bboxes'], inv)
        d_out, _ = extract(depth, out_boxes)
        res = dict(trk)
        res.update(bboxes=out_boxes, depth=d_out, gt_depth=d_out,
                   det_inflated=infl, det_scales=scales, det_depth=d)
        return res

    def step(self, frame_id, raw_u16, cls_maps, bbox_maps, obj_maps):
        pre = self.depth_map(raw_u16)
        det = self.detect(cls_maps, bbox_maps, obj_maps, raw_u16.shape)
        res = self.step_from_dets(frame_id, pre['depth'], det['bboxes'],
                                  det['scores'], det['labels'])
        res.update(det=det, depth_map=pre['depth'])
        return res
