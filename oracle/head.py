"""Row A2: YOLOX head post-process (decode, score, threshold, sort, NMS, clamp).

The code being restated is third-party and NOT in the reference tree; the call
site is mmtrack/models/detectors/yolo_detector_disparity_v1.py:121-122 with
the config configs/_base_/yolox_s_8x8_mmyolo.py:75-81 merged with
configs/stereo_tracking/ocsort/yolox_s_mmyolo_mot_airdrone_disp.py:42:
  mmyolo 0.2.0  YOLOv5Head.predict_by_feat, YOLOXBBoxCoder.decode,
                MlvlPointGenerator(offset=0)
  mmdet 3.0.0rc4 filter_scores_and_topk, BaseDenseHead._bbox_post_process
  mmcv 2.0.0rc3  batched_nms / nms (greedy, IoU > thr suppresses, offset 0)
Versions from README.md:29 / reproducibility.md:373-377.  Parity for this row
is unpinned by the reference; cross-checked with torchvision.ops.nms.

Tie rule (mmcv's descending sort is unstable): equal scores are ordered by
ascending prior index.
"""
import numpy as np


def exp_f32(x):
    """exp evaluated in float64, rounded once to float32."""
    return np.exp(np.asarray(x, np.float32).astype(np.float64)).astype(np.float32)


def sigmoid_f32(x):
    """1/(1+exp(-x)) evaluated in float64, rounded once to float32."""
    x64 = np.asarray(x, np.float32).astype(np.float64)
    return (1.0 / (1.0 + np.exp(-x64))).astype(np.float32)


def flatten_levels(cls_maps, bbox_maps, obj_maps, strides=(8, 16, 32)):
    """predict_by_feat: per level map.permute(0,2,3,1).reshape(-1, C), levels
    concatenated; priors (x*s, y*s) y-major (MlvlPointGenerator offset=0).

    Inputs are per-level arrays [C,h,w] (C=num_classes / 4 / 1) of ONE image.
    Returns cls [P,nc], reg [P,4], obj [P], prior_xy [P,2], stride [P]."""
    cls_l, reg_l, obj_l, pri_l, str_l = [], [], [], [], []
    for c, b, o, s in zip(cls_maps, bbox_maps, obj_maps, strides):
        nc, h, w = c.shape
        cls_l.append(np.asarray(c, np.float32).transpose(1, 2, 0).reshape(-1, nc))
        reg_l.append(np.asarray(b, np.float32).transpose(1, 2, 0).reshape(-1, 4))
        obj_l.append(np.asarray(o, np.float32).reshape(-1))
        ys, xs = np.meshgrid(np.arange(h, dtype=np.float32),
                             np.arange(w, dtype=np.float32), indexing='ij')
        pri_l.append(np.stack([xs.reshape(-1) * np.float32(s),
                               ys.reshape(-1) * np.float32(s)], -1))
        str_l.append(np.full(h * w, s, np.float32))
    return (np.concatenate(cls_l), np.concatenate(reg_l), np.concatenate(obj_l),
            np.concatenate(pri_l), np.concatenate(str_l))


def decode(reg, prior_xy, stride):
    """YOLOXBBoxCoder.decode: xy = p*s + prior; wh = exp(p)*s; xyxy."""
    s = stride[:, None]
    xy = reg[:, :2] * s + prior_xy
    wh = exp_f32(reg[:, 2:]) * s
    half = wh / np.float32(2)
    return np.concatenate([xy - half, xy + half], -1).astype(np.float32)


def iou_gt(a, b, thr):
    """mmcv nms kernel test: inter / (Sa + Sb - inter) > thr, offset 0."""
    w = np.maximum(np.minimum(a[2], b[2]) - np.maximum(a[0], b[0]), np.float32(0))
    h = np.maximum(np.minimum(a[3], b[3]) - np.maximum(a[1], b[1]), np.float32(0))
    inter = np.float32(w * h)
    sa = np.float32((a[2] - a[0]) * (a[3] - a[1]))
    sb = np.float32((b[2] - b[0]) * (b[3] - b[1]))
    with np.errstate(invalid='ignore', divide='ignore'):
        return np.float32(inter / np.float32(np.float32(sa + sb) - inter)) > np.float32(thr)


def greedy_nms(boxes, iou_thr):
    """Greedy NMS over boxes already in descending-score order; returns the
    kept positions in that order (mmcv nms semantics)."""
    n = len(boxes)
    if n == 0:
        return np.zeros(0, np.int64)
    b = np.asarray(boxes, np.float32)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    removed = np.zeros(n, bool)
    keep = []
    thr = np.float32(iou_thr)
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        if i + 1 == n:
            break
        r = b[i + 1:]
        w = np.maximum(np.minimum(b[i, 2], r[:, 2]) - np.maximum(b[i, 0], r[:, 0]), np.float32(0))
        h = np.maximum(np.minimum(b[i, 3], r[:, 3]) - np.maximum(b[i, 1], r[:, 1]), np.float32(0))
        inter = w * h
        with np.errstate(invalid='ignore', divide='ignore'):
            iou = inter / ((area[i] + area[i + 1:]) - inter)
        removed[i + 1:] |= iou > thr
    return np.asarray(keep, np.int64)


def postprocess(cls_maps, bbox_maps, obj_maps, ori_hw, scale_factor=(1.0, 1.0),
                strides=(8, 16, 32), score_thr=0.01, iou_thr=0.5,
                return_debug=False):
    """One image.  Returns dict(bboxes [D,4] f32, scores [D] f32 (descending),
    labels [D] i64, prior_idx [D] i64, keep [D] i64 = positions of the kept
    boxes in the score-sorted candidate list)."""
    cls, reg, obj, pri, strd = flatten_levels(cls_maps, bbox_maps, obj_maps, strides)
    scores_all = sigmoid_f32(cls) * sigmoid_f32(obj)[:, None]
    # multi_label & (num_classes > 1) is False for the 1-class config ->
    # max over classes then filter_scores_and_topk(score_thr, nms_pre=100000)
    labels_all = scores_all.argmax(1)
    smax = scores_all.max(1)
    cand = np.nonzero(smax > np.float32(score_thr))[0]
    order = np.lexsort((cand, -smax[cand].astype(np.float64)))  # score desc, idx asc
    cand = cand[order][:100000]
    boxes = decode(reg[cand], pri[cand], strd[cand])
    sf = np.asarray([scale_factor[0], scale_factor[1]] * 2, np.float32)
    boxes = boxes / sf
    scores = smax[cand]
    labels = labels_all[cand].astype(np.int64)
    if len(cand):
        # batched_nms: offsets = label * (max_coord + 1); single nms call
        off = labels.astype(np.float32) * (boxes.max() + np.float32(1))
        keep = greedy_nms(boxes + off[:, None], iou_thr)
    else:
        keep = np.zeros(0, np.int64)
    out_boxes = boxes[keep].copy()
    out_boxes[:, 0::2] = np.clip(out_boxes[:, 0::2], 0, np.float32(ori_hw[1]))
    out_boxes[:, 1::2] = np.clip(out_boxes[:, 1::2], 0, np.float32(ori_hw[0]))
    res = dict(bboxes=out_boxes.astype(np.float32), scores=scores[keep],
               labels=labels[keep], prior_idx=cand[keep].astype(np.int64),
               keep=keep)
    if return_debug:
        res.update(cand_prior_idx=cand, cand_boxes=boxes, cand_scores=scores)
    return res
