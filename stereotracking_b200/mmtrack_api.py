"""Drop-in classes for the reference's registry API (the boundary of the path).

Same type names, constructor keywords, call signatures and error behaviour as

  OCSORTTracker_Disparity  mmtrack/models/trackers/ocsort_tracker_disparity.py:20
      ctor :42-52 (+ KalmanTrackerBase kalman_tracker_base.py:33-47, BaseTracker
      base_tracker.py:22-32), track() :345-352, reset() base_tracker.py:34-37
  OCSORT_Disparity         mmtrack/models/mot/ocsort_disparity.py:16
      ctor :30-40, predict() :51-111, bbox_postp_depth :113-130
  KalmanFilter             mmtrack/models/motion/kalman_filter.py:12

so a config such as configs/stereo_tracking/ocsort/
yolox_s_mmyolo_mot_airdrone_disp.py builds them through ``MODELS.build`` /
``TASK_UTILS.build`` unchanged.  All arithmetic runs in the CUDA library through
the C ABI (include/stereotrack_b200.h); there is no CPU or PyTorch fallback -
constructing a tracker without the library or without a GPU raises.
One tracker instance follows one sequence, as in the reference; the batched
lockstep form is ``stereotracking_b200.engine``.
"""
import torch

from . import _lib, compat
from .compat import MODELS, TASK_UTILS, InstanceData
from .engine import EngineConfig, StereoTrackEngine

# BaseMultiObjectTracker(mmengine BaseModel) when mmtrack is installed, a stand-in otherwise
_ShellBase = compat.model_shell_base()


def _mesh_affine(curr_img, prev_img, metainfo, step=16, winsize=31, ransac_thr=5.0, min_inlier_ratio=0.3):
    """Host-side camera-motion estimator with the reference's parameters
    (trackers/utils.py:6-55 GLME_affine): both frames as 255x255 equalised grey
    thumbnails, Farneback flow (pyr 0.5, 5 levels, 3 iterations, poly 5 / 1.2), the
    median flow of every step x step cell scaled back to image pixels, and a RANSAC
    4-DoF affine through the cell centres.  Returns the float32 2x3 matrix or None."""
    import cv2
    import numpy as np
    h0, w0 = metainfo['img_shape'][:2]

    def thumb(t):
        a = t.detach().cpu().numpy()[0].transpose(1, 2, 0).astype(np.uint8)[:h0, :w0, :3][:, :, ::-1]
        a = cv2.resize(np.ascontiguousarray(a), (255, 255))
        return cv2.equalizeHist(cv2.cvtColor(a, cv2.COLOR_RGB2GRAY))

    cur, prev = thumb(curr_img), thumb(prev_img)
    fx, fy = w0 / 255, h0 / 255
    flow = cv2.calcOpticalFlowFarneback(prev, cur, None, pyr_scale=0.5, levels=5, winsize=winsize, iterations=3,
                                        poly_n=5, poly_sigma=1.2, flags=0) * np.array([fx, fy])
    gh, gw = cur.shape[0] // step, cur.shape[1] // step
    mesh = np.median(flow[:gh * step, :gw * step].reshape(gh, step, gw, step, 2), axis=(1, 3)).reshape(-1, 2)
    ys, xs = np.mgrid[0:gh, 0:gw]
    src = np.stack([(xs + 0.5) * step * fx, (ys + 0.5) * step * fy], -1).reshape(-1, 2).astype(np.float32)
    dst = (src + mesh).astype(np.float32)
    mat, inl = cv2.estimateAffinePartial2D(src, dst, method=cv2.RANSAC, ransacReprojThreshold=ransac_thr)
    if mat is None or inl is None or float(inl.sum()) / len(src) < min_inlier_ratio:
        return None
    return mat.astype(np.float32)


@TASK_UTILS.register_module(force=True)      # replaces the reference class of the same name
class KalmanFilter:
    """Constant-velocity Kalman filter on (cx, cy, a, h) (kalman_filter.py:12-58).

    The tracker keeps its Kalman state on the device and never calls this object;
    it exists so that ``motion=dict(type='KalmanFilter')`` keeps building and so
    that user code calling initiate / predict / update gets the same numbers
    (float64, computed by the CUDA kernels, one warp per track)."""
    chi2inv95 = {1: 3.8415, 2: 5.9915, 3: 7.8147, 4: 9.4877, 5: 11.070, 6: 12.592, 7: 14.067, 8: 15.507, 9: 16.919}

    def __init__(self, center_only: bool = False, use_nsa: bool = False):
        self.center_only = center_only
        self.gating_threshold = self.chi2inv95[2 if center_only else 4]
        self.use_nsa = use_nsa
        self._lib = _lib.lib()

    def _noise_scale(self, bbox_score):
        """kalman_filter.py:144-145: with use_nsa every measurement std is scaled by 1 - bbox_score."""
        if not self.use_nsa:
            return None
        return torch.tensor([1 - bbox_score], dtype=torch.float64, device='cuda')

    def _run(self, op, mean, cov, z, bbox_score=0.):
        import ctypes as C
        import numpy as np
        if not torch.cuda.is_available():
            raise _lib.StereoTrackError('KalmanFilter needs a CUDA device (no CPU fallback)')
        m = torch.as_tensor(np.asarray(mean, np.float64).reshape(1, 8)).cuda().contiguous()
        c = torch.as_tensor(np.asarray(cov, np.float64).reshape(1, 64)).cuda().contiguous()
        zt = None if z is None else torch.as_tensor(np.asarray(z, np.float32).reshape(1, 4)).cuda().contiguous()
        ns = self._noise_scale(bbox_score) if op == 2 else None
        _lib.check(self._lib.st_kalman_batch_scored(op, 1, C.c_void_p(m.data_ptr()), C.c_void_p(c.data_ptr()),
                                                    C.c_void_p(0 if zt is None else zt.data_ptr()),
                                                    C.c_void_p(0 if ns is None else ns.data_ptr()),
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return m.cpu().numpy().reshape(8), c.cpu().numpy().reshape(8, 8)

    def initiate(self, measurement):
        import numpy as np
        return self._run(0, np.zeros(8), np.zeros((8, 8)), measurement)

    def predict(self, mean, covariance):
        return self._run(1, mean, covariance, None)

    def update(self, mean, covariance, measurement, bbox_score: float = 0.):
        return self._run(2, mean, covariance, measurement, bbox_score)

    def _states(self, mean, covariance):
        import numpy as np
        if not torch.cuda.is_available():
            raise _lib.StereoTrackError('KalmanFilter needs a CUDA device (no CPU fallback)')
        m = torch.as_tensor(np.asarray(mean, np.float64).reshape(-1, 8)).cuda().contiguous()
        c = torch.as_tensor(np.asarray(covariance, np.float64).reshape(-1, 64)).cuda().contiguous()
        return m, c

    def project(self, mean, covariance, bbox_score: float = 0.):
        """kalman_filter.py:121-152 (``bbox_score`` only matters with use_nsa)."""
        m, c = self._states(mean, covariance)
        pm = torch.empty(m.shape[0], 4, dtype=torch.float64, device='cuda')
        pc = torch.empty(m.shape[0], 16, dtype=torch.float64, device='cuda')
        ns = self._noise_scale(bbox_score)
        _lib.check(self._lib.st_kalman_project_scored(m.shape[0], m.data_ptr(), c.data_ptr(),
                                                      None if ns is None else ns.data_ptr(), pm.data_ptr(),
                                                      pc.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   'st_kalman_project_scored')
        return pm.cpu().numpy().reshape(4), pc.cpu().numpy().reshape(4, 4)

    def gating_distance(self, mean, covariance, measurements, only_position: bool = False):
        """kalman_filter.py:190-234: squared Mahalanobis distance of each measurement (N,4)."""
        import numpy as np
        m, c = self._states(mean, covariance)
        z = torch.as_tensor(np.asarray(measurements, np.float64).reshape(-1, 4)).cuda().contiguous()
        out = torch.empty(z.shape[0], dtype=torch.float64, device='cuda')
        _lib.check(self._lib.st_kalman_gating(1, m.data_ptr(), c.data_ptr(), z.shape[0], z.data_ptr(),
                                              int(bool(only_position)), out.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), 'st_kalman_gating')
        return out.cpu().numpy()

    def track(self, tracks, bboxes):
        """kalman_filter.py:236-260: predict every track, gate the boxes (N,4 cxcyah) against it;
        distances above ``gating_threshold`` become NaN.  One predict and one gating launch for
        all tracks."""
        import numpy as np
        ids = list(tracks.keys())
        m, c = self._states(np.stack([tracks[i].mean for i in ids]), np.stack([tracks[i].covariance for i in ids]))
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self._lib.st_kalman_batch(1, len(ids), m.data_ptr(), c.data_ptr(), None, st), 'st_kalman_batch')
        z = torch.as_tensor(bboxes.detach().cpu().numpy().astype(np.float64).reshape(-1, 4)).cuda().contiguous()
        out = torch.empty(len(ids), z.shape[0], dtype=torch.float64, device='cuda')
        _lib.check(self._lib.st_kalman_gating(len(ids), m.data_ptr(), c.data_ptr(), z.shape[0], z.data_ptr(),
                                              int(bool(self.center_only)), out.data_ptr(), st), 'st_kalman_gating')
        mh, ch = m.cpu().numpy(), c.cpu().numpy()
        for k, i in enumerate(ids):
            tracks[i].mean, tracks[i].covariance = mh[k].copy(), ch[k].reshape(8, 8).copy()
        costs = out.cpu().numpy()
        costs[costs > self.gating_threshold] = np.nan
        return tracks, costs


class compat_dict(dict):
    """addict.Dict stand-in for ``memo``: items also readable as attributes."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


@MODELS.register_module(force=True)
class OCSORTTracker_Disparity:
    """OC-SORT tracker carrying per-detection ``scales`` / ``depth``."""

    def __init__(self, obj_score_thr=0.3, init_track_thr=0.7, weight_iou_with_det_scores=True,
                 match_iou_thr=0.3, num_tentatives=3, vel_consist_weight=0.2, vel_delta_t=3, cmc=None,
                 reid=None, momentums=None, num_frames_retain=10, max_tracks=256, max_dets=256,
                 device='cuda:0', keep_memo=True, **kwargs):
        # keep_memo: keep every frame's output rows for ``memo`` / ``get`` (base_tracker.py:120-177;
        # grows without bound like the reference's per-track lists - switch off for endless streams)
        self.keep_memo, self._frames = bool(keep_memo), []
        self._n_live = 0          # live tracks after the last step (host copy; `empty` of base_tracker.py:39-41)
        if kwargs:
            raise TypeError(f'unexpected keyword arguments {sorted(kwargs)}')
        if momentums is not None:
            assert isinstance(momentums, dict), 'momentums must be a dict'   # base_tracker.py:26-27
        method = cmc.get('method') if cmc is not None else None                # :63-76
        if method not in (None, 'glme_affine'):
            raise ValueError(f"Unknown cmc method '{method}', expected 'glme_affine' or None.")
        self.cmc_mode = None if method is None else 'mesh_affine'
        self.prev_cmc_img = None
        self.obj_score_thr, self.init_track_thr = obj_score_thr, init_track_thr
        self.weight_iou_with_det_scores, self.match_iou_thr = weight_iou_with_det_scores, match_iou_thr
        self.num_tentatives, self.vel_consist_weight, self.vel_delta_t = num_tentatives, vel_consist_weight, vel_delta_t
        self.num_frames_retain, self.momentums, self.reid, self.cmc_cfg = num_frames_retain, momentums, reid, cmc
        self.fp16_enabled = False
        self._engine = StereoTrackEngine(EngineConfig(
            n_seq=1, max_tracks=max_tracks, max_dets=max_dets,
            tracker=dict(obj_score_thr=obj_score_thr, init_track_thr=init_track_thr,
                         weight_iou_with_det_scores=weight_iou_with_det_scores, match_iou_thr=match_iou_thr,
                         num_tentatives=num_tentatives, vel_consist_weight=vel_consist_weight,
                         vel_delta_t=vel_delta_t, num_frames_retain=num_frames_retain)), device=device)

    # ---- BaseTracker surface (base_tracker.py:34-52) ---------------------------
    @property
    def with_cmc(self):
        return self.cmc_mode is not None

    def reset_cmc(self):
        self.prev_cmc_img = None

    def estimate_camera_motion(self, img, metainfo):
        """The frame's 2x3 background affine or None (:87-97).  The estimator itself
        (dense Farneback flow -> per-cell median -> RANSAC partial affine,
        trackers/utils.py:6-55) is host-side OpenCV on a 255x255 thumbnail and stays on
        the host; only its application to the Kalman states runs in the tracker kernel."""
        warp = None
        if self.cmc_mode == 'mesh_affine' and self.prev_cmc_img is not None:
            warp = _mesh_affine(img, self.prev_cmc_img, metainfo, **(self.cmc_cfg.get('glme', None) or {}))
        self.prev_cmc_img = img
        return warp

    def reset(self):
        self._engine.reset()
        self._frames = []
        self._n_live = 0

    # memo items in the order of the reference's update(...) call (:598-605)
    MEMO_ITEMS = ('ids', 'bboxes', 'scores', 'labels', 'scales', 'depth', 'frame_ids')

    def _buffers(self, ids):
        """Per-track lists of the memo items as BaseTracker.update builds them
        (base_tracker.py:54-91,:102-118: one ``v[None]`` entry per frame the track was output in),
        for the live tracks ``ids``, rebuilt from the kept per-frame outputs."""
        if not self.keep_memo:
            raise RuntimeError('memo / get need keep_memo=True')
        want = {int(i) for i in ids}
        buf = {i: {k: [] for k in self.MEMO_ITEMS} for i in want}
        for frame_id, rows in self._frames:
            for r, tid in enumerate(rows['ids'].tolist()):
                if tid in want:
                    for k in self.MEMO_ITEMS:
                        buf[tid][k].append(torch.tensor([frame_id]) if k == 'frame_ids' else rows[k][r][None])
        return buf

    @property
    def memo(self):
        """base_tracker.py:120-139: the latest entry of every memo item for all live tracks."""
        live = self.ids
        buf = self._buffers(live)
        outs = compat_dict()
        for k in self.MEMO_ITEMS:
            outs[k] = torch.cat([buf[i][k][-1] for i in live], dim=0)
        return outs

    def get(self, item, ids=None, num_samples=None, behavior=None):
        """base_tracker.py:141-177."""
        if ids is None:
            ids = self.ids
        buf = self._buffers(ids)
        outs = []
        for i in ids:
            out = buf[int(i)][item]
            if num_samples is not None:
                out = torch.cat(out[-num_samples:], dim=0)
                if behavior == 'mean':
                    out = out.mean(dim=0, keepdim=True)
                elif behavior is None:
                    out = out[None]
                else:
                    raise NotImplementedError()
            else:
                out = out[-1]
            outs.append(out)
        return torch.cat(outs, dim=0)

    def _state(self):
        return self._engine.export_state(0)

    @property
    def num_tracks(self):
        return self._state()['num_tracks']

    @property
    def empty(self):
        return self._n_live == 0

    @property
    def ids(self):
        return self._state()['ids'].tolist()

    @property
    def confirmed_ids(self):
        st = self._state()
        return st['ids'][~st['tentative']].tolist()

    @property
    def unconfirmed_ids(self):
        st = self._state()
        return st['ids'][st['tentative']].tolist()

    # ---- track() (:345-618) -----------------------------------------------------
    def track(self, model, img, feats, data_sample, data_preprocessor=None, rescale=False, **kwargs):
        det = data_sample.pred_det_instances
        return self._track(img, data_sample.metainfo, det.bboxes, det.scores, det.labels, det.scales, det.depth)

    def _track(self, img, metainfo, bboxes, scores, labels, scales, depth):
        """One frame given the pred_det_instances fields as tensors (any device)."""
        eng = self._engine
        frame_id = metainfo.get('frame_id', -1)
        n = int(bboxes.shape[0])
        if n > eng.cfg.max_dets:
            raise ValueError(f'{n} detections exceed max_dets={eng.cfg.max_dets}')
        dev = eng.device
        eng.det_inflated[0, :n] = bboxes.to(dev, torch.float32)
        eng.det_scores[0, :n] = scores.to(dev, torch.float32)
        eng.det_labels[0, :n] = labels.to(dev, torch.int64)
        eng.det_scales[0, :n] = scales.to(dev, torch.float32)
        eng.det_depth[0, :n] = depth.to(dev, torch.float32)
        eng.det_counts.fill_(n)
        eng.frame_ids.fill_(int(frame_id))
        warp = None
        if frame_id == 0:
            # :385-387: reset() and reset_cmc() come first, so frame 0 of every video takes the
            # init-only branch (:391) and never reaches the estimator (the kernel resets the
            # device state itself when it sees frame_id == 0)
            self.reset_cmc()
            self._frames = []
            self._n_live = 0
        if self.with_cmc:
            # the reference estimates only in the association branch (:391, :428-429)
            w = self.estimate_camera_motion(img, metainfo) if (n > 0 and not self.empty) else None
            warp = torch.full((1, 6), float('nan'), dtype=torch.float32, device=dev)
            if w is not None:
                warp[0] = torch.as_tensor(w, dtype=torch.float32).reshape(6).to(dev)
        out = eng.track_step(eng.frame_ids, eng.det_inflated, eng.det_scores, eng.det_labels, eng.det_scales,
                             eng.det_depth, eng.det_counts, warp=warp)
        m = int(out['counts'][0])
        self._n_live = int(eng.live_counts()[0])
        status = int(eng.status()[0])
        if status:
            raise _lib.StereoTrackError(f'tracker capacity exceeded (status bits {status}); raise max_tracks / max_dets')
        tgt = bboxes.device
        res = InstanceData()
        res.bboxes = out['bboxes'][0, :m].clone().to(tgt)
        res.labels = out['labels'][0, :m].clone().to(tgt)
        res.scores = out['scores'][0, :m].clone().to(tgt)
        res.scales = out['scales'][0, :m].clone().to(tgt)
        res.depth = out['depth'][0, :m].clone().to(tgt)
        res.instances_id = out['instances_id'][0, :m].clone().to(tgt)
        if self.keep_memo:
            self._frames.append((int(frame_id), dict(ids=res.instances_id, bboxes=res.bboxes, scores=res.scores,
                                                     labels=res.labels, scales=res.scales, depth=res.depth)))
        return res


@MODELS.register_module(force=True)
class OCSORTTracker(OCSORTTracker_Disparity):
    """RGB-only OC-SORT tracker (mmtrack/models/trackers/ocsort_tracker.py:21): the same
    tracker without the per-detection ``scales`` / ``depth`` fields - its diff against
    the disparity tracker is only those two fields - so it runs on the same kernels
    with scales = 1 and depth = 0 and returns bboxes / labels / scores / instances_id."""

    MEMO_ITEMS = ('ids', 'bboxes', 'scores', 'labels', 'frame_ids')          # ocsort_tracker.py update(...) call

    def track(self, model, img, feats, data_sample, data_preprocessor=None, rescale=False, **kwargs):
        det = data_sample.pred_det_instances                 # left untouched: no scales / depth fields here
        n = int(det.bboxes.shape[0])
        ones = torch.ones(n, dtype=torch.float32, device=det.bboxes.device)
        full = self._track(img, data_sample.metainfo, det.bboxes, det.scores, det.labels, ones, torch.zeros_like(ones))
        res = InstanceData()
        res.bboxes, res.labels, res.scores, res.instances_id = full.bboxes, full.labels, full.scores, full.instances_id
        return res


class OCSORT(_ShellBase):
    """RGB-only model shell (mmtrack/models/mot/ocsort.py:13, predict :71-114): detector ->
    tracker, no depth stage.  A ``BaseMultiObjectTracker`` like the reference class
    (mot/ocsort.py:27-42): ``data_preprocessor`` and ``init_cfg`` go to the base model.
    NOT registered on import: see ``register_model_shells``."""

    def __init__(self, detector=None, tracker=None, motion=None, data_preprocessor=None, init_cfg=None):
        super().__init__(data_preprocessor, init_cfg)
        if detector is not None:
            self.detector = MODELS.build(detector) if isinstance(detector, dict) else detector
        if motion is not None:
            self.motion = TASK_UTILS.build(motion) if isinstance(motion, dict) else motion
        if tracker is not None:
            self.tracker = MODELS.build(tracker) if isinstance(tracker, dict) else tracker

    def loss(self, inputs, data_samples, **kwargs):                      # mot/ocsort.py:44-67
        img = inputs['img']
        assert img.size(1) == 1
        return self.detector.loss(torch.squeeze(img, dim=1), data_samples, **kwargs)

    def predict(self, inputs, data_samples, **kwargs):
        img = inputs['img']
        assert img.dim() == 5, 'The img must be 5D Tensor (N, T, C, H, W).'
        assert img.size(0) == 1, 'OCSort inference only support 1 batch size per gpu for now.'
        img = img[0]
        assert len(data_samples) == 1, 'OCSort inference only support 1 batch size per gpu for now.'
        sample = data_samples[0]
        det_results = self.detector.predict(img, data_samples)
        assert len(det_results) == 1, 'Batch inference is not supported.'
        sample.pred_det_instances = det_results[0].pred_instances.clone()
        sample.pred_track_instances = self.tracker.track(model=self, img=img, feats=None, data_sample=sample,
                                                         **kwargs)
        return [sample]


class OCSORT_Disparity(_ShellBase):
    """Model shell of the hot path: detector -> depth -> per-box depth/scale ->
    tracker -> deflate -> output depth (ocsort_disparity.py:51-111).  ``detector``
    may be any object with ``predict(data, data_samples)`` returning a list whose
    first element has ``pred_instances`` (the reference PyTorch detector stays
    the detector; it is not part of this path).  A ``BaseMultiObjectTracker`` like the
    reference class (ocsort_disparity.py:15-40 via mot/ocsort.py:27-42).  NOT registered on
    import: see ``register_model_shells``."""

    def __init__(self, detector=None, tracker=None, motion=None, data_preprocessor=None, init_cfg=None,
                 baseline: float = 0.25, focal_length: int = 640, depth_reducer: str = 'window'):
        super().__init__(data_preprocessor, init_cfg)
        # depth_reducer: 'window' = extract_depth (:136-175); 'truncated_mean' | 'mean' | 'median' |
        # 'center' = what decorating extract_depth with depth_extraction_comparison.py's
        # *_decorator does in the reference's ablation runs
        if depth_reducer not in StereoTrackEngine.DEPTH_MODES:
            raise ValueError(f'unknown depth_reducer {depth_reducer!r}')
        self.depth_reducer = depth_reducer
        self.detector = MODELS.build(detector) if isinstance(detector, dict) else detector
        self.motion = TASK_UTILS.build(motion) if isinstance(motion, dict) else motion
        self.tracker = MODELS.build(tracker) if isinstance(tracker, dict) else tracker
        self.baseline, self.focal_length = baseline, focal_length
        if self.tracker is not None and hasattr(self.tracker, '_engine'):
            self.tracker._engine.cfg.baseline = float(baseline)
            self.tracker._engine.cfg.focal_length = float(focal_length)

    def loss(self, inputs, data_samples, **kwargs):                      # ocsort_disparity.py:42-48
        raise NotImplementedError('training the detector is outside this path; use the reference OCSORT_Disparity')

    def disp2depth(self, disp):
        return self.tracker._engine.disp_to_depth(disp.contiguous().float())

    @staticmethod
    def _depth_list(t):
        """extract_depth returns a Python list of numpy float32 means and the int -1 for boxes
        without a valid depth (ocsort_disparity.py:148,:168); results.csv prints them as such."""
        import numpy as np
        return [(-1 if v == -1 else np.float32(v)) for v in t.cpu().numpy()]

    def _check_center(self, eng):
        # depth_extraction_comparison.py:141 ``depth[cy, cx]``: NumPy raises when the box centre is off the map
        if self.depth_reducer == 'center' and int(eng.depth_oob.sum()) > 0:
            raise IndexError('box centre outside the depth map')

    def bbox_postp_depth(self, pred_instances, disp, gt_depth=None):
        """:113-130 - returns (pred_instances with inflated bboxes / scales / depth, depth dict)."""
        eng = self.tracker._engine
        depth = self.disp2depth(disp[:, 0:1, :, :]).reshape(1, disp.shape[-2], disp.shape[-1])
        boxes = pred_instances['bboxes']
        n = int(boxes.shape[0])
        cap = max(n, 1)
        b = torch.zeros(1, cap, 4, dtype=torch.float32, device=eng.device)
        b[0, :n] = boxes.to(eng.device, torch.float32)
        cnt = torch.tensor([n], dtype=torch.int32, device=eng.device)
        infl = torch.zeros_like(b)
        d, sc, _ = eng.box_depth(depth, b, cnt, inflated_out=infl, mode=self.depth_reducer)
        self._check_center(eng)
        values = dict(d_values=self._depth_list(d[0, :n]))
        if gt_depth is not None:
            gd, _, _ = eng.box_depth(gt_depth.reshape(1, gt_depth.shape[-2], gt_depth.shape[-1]).contiguous().float(), b, cnt,
                                     mode=self.depth_reducer)
            self._check_center(eng)
            values['gt_d_values'] = self._depth_list(gd[0, :n])
        pred_instances['bboxes'] = infl[0, :n].to(boxes)
        pred_instances['scales'] = sc[0, :n].to(boxes)
        pred_instances['depth'] = d[0, :n].to(boxes)
        return pred_instances, values

    def predict(self, inputs, data_samples, **kwargs):
        img, disp_postp, disp_mask = inputs['img'], inputs['disp_postp'], inputs['disp_mask']
        depth_postp = inputs.get('depth_postp', None)
        assert img.dim() == 5, 'The img must be 5D Tensor (N, T, C, H, W).'
        assert img.size(0) == 1, 'OCSort inference only support 1 batch size per gpu for now.'
        img, disp_postp, disp_mask = img[0], disp_postp[0], disp_mask[0]
        if depth_postp is not None:
            depth_postp = depth_postp[0]
        assert len(data_samples) == 1, 'OCSort inference only support 1 batch size per gpu for now.'
        sample = data_samples[0]
        det_results = self.detector.predict(dict(img=img, disp_postp=disp_postp, disp_mask=disp_mask), data_samples)
        assert len(det_results) == 1, 'Batch inference is not supported.'
        scaled, _ = self.bbox_postp_depth(det_results[0].pred_instances.clone(), disp_postp, depth_postp)
        sample.pred_det_instances = scaled
        tracks = self.tracker.track(model=self, img=img, feats=None, data_sample=sample, **kwargs)
        eng = self.tracker._engine
        m = int(tracks.bboxes.shape[0])
        cnt = torch.tensor([m], dtype=torch.int32, device=eng.device)
        bb = torch.zeros(1, max(m, 1), 4, dtype=torch.float32, device=eng.device)
        sc = torch.ones(1, max(m, 1), dtype=torch.float32, device=eng.device)
        bb[0, :m] = tracks.bboxes.to(eng.device)
        sc[0, :m] = tracks.scales.to(eng.device)
        tracks.bboxes = eng.scale_boxes(bb, sc, cnt, True)[0, :m].to(tracks.scores)      # :96-97
        _, depth = self.bbox_postp_depth(tracks.clone(), disp_postp, depth_postp)       # :100-101
        tracks['depth'] = depth['d_values']
        tracks['gt_depth'] = depth.get('gt_d_values', depth['d_values'])
        sample.pred_det_instances = det_results[0].pred_instances.clone()
        sample.pred_track_instances = tracks
        return [sample]


def register_model_shells(force=True):
    """Explicit opt-in: register the model shells ``OCSORT`` / ``OCSORT_Disparity`` under the
    reference's type names (importing ``stereotracking_b200.mmtrack_shell`` calls this).  Importing
    this module alone registers only the trackers and ``KalmanFilter`` - the tracker-only drop-in
    under the reference's own model shell."""
    MODELS.register_module(name='OCSORT', force=force, module=OCSORT)
    MODELS.register_module(name='OCSORT_Disparity', force=force, module=OCSORT_Disparity)
    return OCSORT, OCSORT_Disparity
