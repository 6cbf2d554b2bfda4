"""Rows A6-A9: the OC-SORT tracker with depth-scaled boxes, one sequence.

Follows
  mmtrack/models/trackers/ocsort_tracker_disparity.py  (track :345-618,
      ocm_assign_ids :187-265, ocr_assign_ids :273-317, online_smooth
      :319-343, init_track :105-126, update_track :128-146, vel_direction
      :148-156, vel_direction_batch :158-171, k_step_observation :173-185,
      last_obs :267-271)
  mmtrack/models/trackers/kalman_tracker_base.py:50-88
  mmtrack/models/trackers/base_tracker.py:34-121
  mmtrack/structures/bbox/transforms.py:72-101 (cxcyah <-> xyxy, float32)
  mmdet 3.0.0rc4 bbox_overlaps (third-party; eps=1e-6 clamp on the union)

float32 torch arithmetic of the reference is reproduced with NumPy float32 in
the same evaluation order; acos is float64-evaluated then rounded (see
oracle/__init__.py).  Track containers are plain Python objects kept in a
dict in insertion (= ascending id) order like the reference's ``self.tracks``.
"""
import numpy as np

from . import kalman
from .lapjv import lapjv

F32 = np.float32
PI_F32 = F32(np.pi)
HALF_PI_F32 = F32(np.pi / 2.0)


def xyxy_to_cxcyah(b):
    """structures/bbox/transforms.py:72-86 for one box, float32."""
    b = np.asarray(b, F32)
    w = b[2] - b[0]
    h = b[3] - b[1]
    with np.errstate(invalid='ignore', divide='ignore'):
        return np.asarray([(b[2] + b[0]) / F32(2), (b[3] + b[1]) / F32(2),
                           w / h, h], F32)


def cxcyah_to_xyxy(m):
    """structures/bbox/transforms.py:89-101 for [T,4] float32."""
    m = np.asarray(m, F32).reshape(-1, 4)
    w = m[:, 2] * m[:, 3]
    hw = w / F32(2)
    hh = m[:, 3] / F32(2)
    return np.stack([m[:, 0] - hw, m[:, 1] - hh, m[:, 0] + hw, m[:, 1] + hh], -1)


def bbox_overlaps(b1, b2):
    """mmdet bbox_overlaps(mode='iou', eps=1e-6), float32, [R,4]x[C,4]."""
    b1 = np.asarray(b1, F32).reshape(-1, 4)
    b2 = np.asarray(b2, F32).reshape(-1, 4)
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = np.maximum(b1[:, None, :2], b2[None, :, :2])
    rb = np.minimum(b1[:, None, 2:], b2[None, :, 2:])
    wh = np.maximum(rb - lt, F32(0))
    inter = wh[..., 0] * wh[..., 1]
    union = np.maximum(a1[:, None] + a2[None, :] - inter, F32(1e-6))
    with np.errstate(invalid='ignore'):
        return (inter / union).astype(F32)


def _sum4(b):
    """torch ``tensor.sum()`` of 4 float32 values (sign test only)."""
    return F32(F32(F32(b[0] + b[1]) + b[2]) + b[3])


def acos_f32(x):
    return np.arccos(np.asarray(x, F32).astype(np.float64)).astype(F32)


class Track:
    __slots__ = ('id', 'tentative', 'tracked', 'hits', 'last_frame', 'mean',
                 'cov', 'saved', 'obs', 'velocity', 'bbox', 'score', 'label',
                 'scale', 'depth')


class OCSortDisparityOracle:
    """Constructor keys follow ocsort_tracker_disparity.py:42-52 /
    kalman_tracker_base.py:33-47 / base_tracker.py:22-32 (same defaults)."""

    def __init__(self, obj_score_thr=0.3, init_track_thr=0.7,
                 weight_iou_with_det_scores=True, match_iou_thr=0.3,
                 num_tentatives=3, vel_consist_weight=0.2, vel_delta_t=3,
                 num_frames_retain=10, cmc=None, **_ignored):
        self.cmc = cmc        # the warp itself is passed to track(); estimation is host-side OpenCV
        self.obj_score_thr = obj_score_thr
        self.init_track_thr = init_track_thr
        self.weight_iou_with_det_scores = weight_iou_with_det_scores
        self.match_iou_thr = match_iou_thr
        self.num_tentatives = num_tentatives
        self.vel_consist_weight = vel_consist_weight
        self.vel_delta_t = vel_delta_t
        self.num_frames_retain = num_frames_retain
        self.reset()

    # base_tracker.py:34-37
    def reset(self):
        self.num_tracks = 0
        self.tracks = {}

    @property
    def confirmed_ids(self):
        return [i for i, t in self.tracks.items() if not t.tentative]

    @property
    def unconfirmed_ids(self):
        return [i for i, t in self.tracks.items() if t.tentative]

    # ---- observation helpers (:173-185, :267-271) -------------------------
    @staticmethod
    def last_obs(t):
        for b in reversed(t.obs):
            if b is not None:
                return b
        return None

    def k_step_observation(self, t):
        n = len(t.obs)
        if n == 0:
            return np.full(4, -1, F32)
        if n > self.vel_delta_t and t.obs[n - 1 - self.vel_delta_t] is not None:
            return t.obs[n - 1 - self.vel_delta_t]
        return self.last_obs(t)

    @staticmethod
    def vel_direction(b1, b2):
        """:148-156 -> (dy, dx)/(norm+1e-6), or the (-1,-1) sentinel."""
        if _sum4(b1) < 0 or _sum4(b2) < 0:
            return np.asarray([-1, -1], F32)
        cx1, cy1 = (b1[0] + b1[2]) / F32(2), (b1[1] + b1[3]) / F32(2)
        cx2, cy2 = (b2[0] + b2[2]) / F32(2), (b2[1] + b2[3]) / F32(2)
        sy, sx = F32(cy2 - cy1), F32(cx2 - cx1)
        norm = F32(np.sqrt(F32(F32(sy * sy) + F32(sx * sx)))) + F32(1e-6)
        return np.asarray([sy / norm, sx / norm], F32)

    # ---- association costs (:187-265, :273-317) ----------------------------
    def ocm_cost(self, ids, det_bboxes, det_scores):
        T, D = len(ids), len(det_bboxes)
        means = np.zeros((0, 4))
        for i in ids:
            means = np.concatenate([means, self.tracks[i].mean[:4][None]], 0)
        tb = cxcyah_to_xyxy(means.astype(F32))
        ious = bbox_overlaps(tb, det_bboxes)
        if self.weight_iou_with_det_scores:
            ious = ious * np.asarray(det_scores, F32)[None]
        dists = (F32(1) - ious).astype(F32)
        if T > 0 and D > 0:
            vel = np.stack([self.tracks[i].velocity for i in ids]).astype(F32)
            kobs = np.stack([self.k_step_observation(self.tracks[i])
                             for i in ids]).astype(F32)
            valid = ((vel[:, 0] + vel[:, 1]) != F32(-2)) & \
                (np.asarray([_sum4(k) for k in kobs]) != F32(-4))
            two = F32(2)
            cx1 = (kobs[:, 0] + kobs[:, 2]) / two
            cy1 = (kobs[:, 1] + kobs[:, 3]) / two
            cx2 = (det_bboxes[:, 0] + det_bboxes[:, 2]) / two
            cy2 = (det_bboxes[:, 1] + det_bboxes[:, 3]) / two
            dy = cy2[None, :] - cy1[:, None]
            dx = cx2[None, :] - cx1[:, None]
            norm = np.sqrt(dy * dy + dx * dx) + F32(1e-6)
            dy = dy / norm
            dx = dx / norm
            cosang = dy * vel[:, 0:1] + dx * vel[:, 1:2]
            cosang = np.clip(cosang, F32(-1), F32(1))
            ang = acos_f32(cosang)
            norm_ang = (ang - HALF_PI_F32) / PI_F32
            term = norm_ang * valid[:, None].astype(F32)
            dists = dists + term * F32(self.vel_consist_weight)
        return dists.astype(F32)

    def ocr_cost(self, track_obs, det_bboxes, det_scores):
        ious = bbox_overlaps(track_obs, det_bboxes)
        if self.weight_iou_with_det_scores:
            ious = ious * np.asarray(det_scores, F32)[None]
        return (F32(1) - ious).astype(F32)

    def _assign(self, dists):
        R, C = dists.shape
        if dists.size > 0:
            return lapjv(dists, 1 - self.match_iou_thr)
        return np.full(R, -1, np.int32), np.full(C, -1, np.int32)

    # ---- track bookkeeping (base_tracker.py:54-121 and subclasses) ---------
    def _init_track(self, tid, box, score, label, scale, depth, frame_id):
        t = Track()
        t.id = tid
        t.hits = 1
        t.last_frame = frame_id
        t.tentative = frame_id != 0                       # :108-111
        t.mean, t.cov = kalman.initiate(xyxy_to_cxcyah(box))
        t.obs = [box]
        t.tracked = True
        t.saved = None
        t.velocity = np.asarray([-1, -1], F32)            # :125-126
        t.bbox, t.score, t.label, t.scale, t.depth = box, score, label, scale, depth
        self.tracks[tid] = t

    def _update_track(self, tid, box, score, label, scale, depth, frame_id):
        t = self.tracks[tid]
        t.hits += 1
        t.last_frame = frame_id
        t.bbox, t.score, t.label, t.scale, t.depth = box, score, label, scale, depth
        if t.tentative and t.hits >= self.num_tentatives:  # kalman_tracker_base.py:68-70
            t.tentative = False
        t.mean, t.cov = kalman.update(t.mean, t.cov, xyxy_to_cxcyah(box))
        t.tracked = True
        t.obs.append(box)
        t.velocity = self.vel_direction(self.k_step_observation(t), box)

    def _pop_invalid(self, frame_id):
        """kalman_tracker_base.py:77-88."""
        dead = [i for i, t in self.tracks.items()
                if frame_id - t.last_frame >= self.num_frames_retain
                or (t.tentative and t.last_frame != frame_id)]
        for i in dead:
            self.tracks.pop(i)

    def _online_smooth(self, t, box):
        """:319-343."""
        last = self.last_obs(t)
        miss = 0
        for b in reversed(t.obs):
            if b is None:
                miss += 1
            else:
                break
        shift = (box - last) / F32(miss + 1)
        t.mean, t.cov = t.saved
        for i in range(miss):
            virt = (last + F32(i + 1) * shift).astype(F32)
            t.mean, t.cov = kalman.update(t.mean, t.cov, xyxy_to_cxcyah(virt))

    # ---- one frame (:345-618) ------------------------------------------------
    def apply_gmc(self, ids, warp):
        """apply_gmc_to_tracks_cxcyah (mmtrack/models/trackers/gmc.py:20-45): the 2x3
        background affine moves the persistent Kalman state of the given tracks:
        centre and its velocity through R (+ t), h and vh by s = sqrt(max(det R,
        1e-12)), covariance through the same linear map A P A^T."""
        warp = np.asarray(warp)
        R = warp[:2, :2].astype(np.float64)
        t = warp[:2, 2].astype(np.float64)
        s = float(np.sqrt(max(np.linalg.det(R), 1e-12)))
        A = np.eye(8, dtype=float)
        A[0:2, 0:2] = R
        A[4:6, 4:6] = R
        A[3, 3] = s
        A[7, 7] = s
        for i in ids:
            tr = self.tracks[i]
            mean = tr.mean.copy()
            mean[0:2] = R.dot(mean[0:2]) + t
            mean[3] *= s
            mean[4:6] = R.dot(mean[4:6])
            mean[7] *= s
            tr.mean = mean
            tr.cov = A.dot(tr.cov).dot(A.T)

    def track(self, frame_id, bboxes, scores, labels, scales, depth, debug=None, warp=None):
        """Inputs are the pred_det_instances fields (boxes already inflated); ``warp`` is
        the 2x3 camera-motion affine of this frame or None (ocsort_tracker_disparity.py
        :428-445, cmc=dict(method='glme_affine')).
        Returns dict(bboxes, labels, scores, scales, depth, instances_id)."""
        bboxes = np.asarray(bboxes, F32).reshape(-1, 4)
        scores = np.asarray(scores, F32).reshape(-1)
        labels = np.asarray(labels, np.int64).reshape(-1)
        scales = np.asarray(scales, F32).reshape(-1)
        depth = np.asarray(depth, F32).reshape(-1)
        if frame_id == 0:
            self.reset()
        if not self.tracks or len(bboxes) == 0:
            sel = scores > F32(self.init_track_thr)
            bboxes, scores, labels = bboxes[sel], scores[sel], labels[sel]
            scales, depth = scales[sel], depth[sel]
            n = len(bboxes)
            ids = np.arange(self.num_tracks, self.num_tracks + n, dtype=np.int64)
            self.num_tracks += n
        else:
            with np.errstate(invalid='ignore'):
                area = (bboxes[:, 2] - bboxes[:, 0]) * (bboxes[:, 3] - bboxes[:, 1])
                sel = (scores > F32(self.obj_score_thr)) & (area > F32(100))
            det = np.nonzero(sel)[0]
            det_ids = np.full(len(det), -1, np.int64)

            confirmed = self.confirmed_ids
            for i in confirmed:
                t = self.tracks[i]
                if t.last_frame != frame_id - 1:
                    t.mean = t.mean.copy()
                    t.mean[7] = 0
                if t.tracked:
                    t.saved = (t.mean, t.cov)
                t.mean, t.cov = kalman.predict(t.mean, t.cov)
            if warp is not None:                      # :443-445
                self.apply_gmc(confirmed, warp)

            # stage 1: confirmed tracks (:448-474)
            c1 = self.ocm_cost(confirmed, bboxes[det], scores[det])
            _, col = self._assign(c1)
            m1 = col > -1
            det_ids[m1] = np.asarray(confirmed, np.int64)[col[m1]] if m1.any() else det_ids[m1]
            matched = [det[m1]]
            rest = det[~m1]
            rest_ids = det_ids[~m1]

            # stage 2: tentative tracks vs leftovers (:476-503)
            unconfirmed = self.unconfirmed_ids
            c2 = self.ocm_cost(unconfirmed, bboxes[rest], scores[rest])
            _, col2 = self._assign(c2)
            m2 = col2 > -1
            if m2.any():
                rest_ids[m2] = np.asarray(unconfirmed, np.int64)[col2[m2]]
            matched.append(rest[m2])
            matched_ids = np.concatenate([det_ids[m1], rest_ids[m2]])
            rest = rest[~m2]

            # stage 3: observation-centric recovery (:505-569)
            all_ids = list(self.tracks.keys())
            pool = [i for i in all_ids if i not in set(matched_ids.tolist())]
            c3 = None
            if len(pool) > 0:
                last = np.stack([self.last_obs(self.tracks[i]) for i in pool]).astype(F32)
                c3 = self.ocr_cost(last, bboxes[rest], scores[rest])
                _, col3 = self._assign(c3)
                m3 = col3 > -1
                ids3 = np.asarray(pool, np.int64)[col3[m3]] if m3.any() else np.zeros(0, np.int64)
                matched.append(rest[m3])
                matched_ids = np.concatenate([matched_ids, ids3])
                rest = rest[~m3]
            matched = np.concatenate(matched)

            # 5. summarise (:571-589)
            for di, tid in zip(matched, matched_ids):
                t = self.tracks[int(tid)]
                if not t.tracked:
                    self._online_smooth(t, bboxes[di])
            mset = set(matched_ids.tolist())
            for i in all_ids:
                if i not in mset:
                    self.tracks[i].tracked = False
                    self.tracks[i].obs.append(None)

            order = np.concatenate([matched, rest])
            n_new = len(rest)
            ids = np.concatenate([matched_ids, np.arange(
                self.num_tracks, self.num_tracks + n_new, dtype=np.int64)])
            self.num_tracks += n_new
            if debug is not None:
                debug.update(cost1=c1, cost2=c2, cost3=c3, det=det)
            bboxes, scores, labels = bboxes[order], scores[order], labels[order]
            scales, depth = scales[order], depth[order]

        # base_tracker.py:54-91 update()
        for k in range(len(ids)):
            tid = int(ids[k])
            args = (tid, bboxes[k].copy(), scores[k], labels[k], scales[k], depth[k], frame_id)
            if tid in self.tracks:
                self._update_track(*args)
            else:
                self._init_track(*args)
        self._pop_invalid(frame_id)
        return dict(bboxes=bboxes, labels=labels, scores=scores, scales=scales,
                    depth=depth, instances_id=ids)

    # ---- state export for parity checks ------------------------------------
    def export_state(self):
        ids = list(self.tracks.keys())
        ts = [self.tracks[i] for i in ids]
        return dict(
            ids=np.asarray(ids, np.int64),
            mean=np.stack([t.mean for t in ts]) if ts else np.zeros((0, 8)),
            cov=np.stack([t.cov for t in ts]) if ts else np.zeros((0, 8, 8)),
            tentative=np.asarray([t.tentative for t in ts], bool),
            tracked=np.asarray([t.tracked for t in ts], bool),
            hits=np.asarray([t.hits for t in ts], np.int64),
            last_frame=np.asarray([t.last_frame for t in ts], np.int64),
            velocity=np.stack([t.velocity for t in ts]).astype(F32) if ts else np.zeros((0, 2), F32),
            num_tracks=self.num_tracks)
