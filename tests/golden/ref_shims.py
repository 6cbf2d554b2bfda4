"""Load the UNMODIFIED reference modules of the hot path in the build container.

The reference (/root/reference, read-only) cannot be imported normally here:
mmengine / mmcv / mmdet / mmyolo / lap / addict are not installed and
``mmtrack/__init__.py`` imports them.  This module registers light stand-ins
for exactly those third-party names in ``sys.modules`` and exposes the
reference's own files

    mmtrack/models/trackers/ocsort_tracker_disparity.py
    mmtrack/models/trackers/kalman_tracker_base.py, base_tracker.py, utils.py
    mmtrack/models/motion/kalman_filter.py
    mmtrack/models/mot/ocsort_disparity.py
    mmtrack/structures/bbox/transforms.py

executed as they are from /root/reference (nothing is copied).  It is used only
by ``make_golden.py`` to generate the committed fixtures; /root/reference does
not exist on the GPU box, so nothing else imports this file.

Stand-ins (third-party code that is not in the reference tree):
  * ``lap.lapjv``                -> scipy.optimize.linear_sum_assignment on the
    extended matrix ``lap`` builds (exact optimum; identical to lap on tie-free
    costs - deliberately independent of oracle/lapjv.c)
  * ``mmdet...bbox_overlaps``    -> torch float32 restatement (eps=1e-6)
  * ``mmengine.structures.InstanceData`` / ``addict.Dict`` -> minimal containers
  * registries / typing aliases  -> no-op decorators
Environment shims (the reference pins NumPy 1.x, this container has 2.x):
  * ``np.int`` alias (used at ocsort_disparity.py:141)
  * KalmanFilter.initiate receives the float32 measurement promoted to float64,
    which is what NumPy 1.x value-based casting does inside initiate
    (python-float * np.float32 -> float64); NumPy 2 would keep float32.
"""
import importlib
import os
import sys
import types

import numpy as np
import scipy.optimize
import torch

REF = os.environ.get('STEREOTRACKING_REF', '/root/reference')


class InstanceData:
    """Stand-in for mmengine.structures.InstanceData (fields by attribute or
    by string key, boolean/index selection over all fields, clone)."""

    def __init__(self, **kw):
        object.__setattr__(self, '_f', {})
        for k, v in kw.items():
            self._f[k] = v

    def __setattr__(self, k, v):
        self._f[k] = v

    __setitem__ = __setattr__

    def __getattr__(self, k):
        try:
            return object.__getattribute__(self, '_f')[k]
        except KeyError:
            raise AttributeError(k)

    def __getitem__(self, item):
        if isinstance(item, str):
            return self._f[item]
        return InstanceData(**{k: v[item] for k, v in self._f.items()})

    def get(self, k, default=None):
        return self._f.get(k, default)

    def keys(self):
        return self._f.keys()

    def clone(self):
        return InstanceData(**{k: (v.clone() if torch.is_tensor(v) else v)
                               for k, v in self._f.items()})

    def __len__(self):
        for v in self._f.values():
            return len(v)
        return 0


class TrackDataSample:
    def __init__(self, frame_id=0, **meta):
        self.metainfo = dict(frame_id=frame_id, **meta)
        self.frame_id = frame_id


class AttrDict(dict):
    """Stand-in for addict.Dict."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def lapjv_standin(cost, extend_cost=False, cost_limit=np.inf, return_cost=True):
    c = np.ascontiguousarray(cost, dtype=np.float64)
    R, C = c.shape
    n = R + C
    ext = np.full((n, n), cost_limit / 2.0)
    ext[R:, C:] = 0
    ext[:R, :C] = c
    rows, cols = scipy.optimize.linear_sum_assignment(ext)
    x = np.empty(n, np.int32)
    y = np.empty(n, np.int32)
    x[rows] = cols
    y[cols] = rows
    x[x >= C] = -1
    y[y >= R] = -1
    return float(ext[rows, cols].sum()), x[:R], y[:C]


def bbox_overlaps_standin(b1, b2, mode='iou', is_aligned=False, eps=1e-6):
    a1 = (b1[..., 2] - b1[..., 0]) * (b1[..., 3] - b1[..., 1])
    a2 = (b2[..., 2] - b2[..., 0]) * (b2[..., 3] - b2[..., 1])
    lt = torch.max(b1[..., :, None, :2], b2[..., None, :, :2])
    rb = torch.min(b1[..., :, None, 2:], b2[..., None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    union = a1[..., None] + a2[..., None, :] - inter
    union = torch.max(union, union.new_tensor([eps]))
    return inter / union


class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name, path):
    m = _mod(name)
    m.__path__ = [path]
    return m


_LOADED = None


def load():
    """Returns a namespace with the reference classes/functions."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not os.path.isdir(REF):
        raise RuntimeError(f'reference tree not found at {REF}')
    if not hasattr(np, 'int'):
        np.int = int
    _mod('lap', lapjv=lapjv_standin)
    _mod('addict', Dict=AttrDict)
    _mod('mmengine')
    _mod('mmengine.structures', InstanceData=InstanceData, BaseDataElement=object)
    _mod('mmdet')
    _mod('mmdet.structures')
    _mod('mmdet.structures.bbox', bbox_overlaps=bbox_overlaps_standin)
    _mod('mmdet.structures.bbox.transforms', bbox_xyxy_to_cxcywh=None)
    mm = os.path.join(REF, 'mmtrack')
    _pkg('mmtrack', mm)
    _mod('mmtrack.registry', MODELS=_Registry(), TASK_UTILS=_Registry())
    _mod('mmtrack.utils', OptConfigType=None, OptMultiConfig=None, SampleList=None,
         imrenormalize=None)
    _mod('mmtrack.utils.collect_results',
         save_prediction_results=lambda file_path: (lambda f: f))
    _mod('mmtrack.structures', TrackDataSample=TrackDataSample)
    _pkg('mmtrack.structures.bbox', os.path.join(mm, 'structures', 'bbox'))
    tr = importlib.import_module('mmtrack.structures.bbox.transforms')
    sys.modules['mmtrack.structures.bbox'].bbox_cxcyah_to_xyxy = tr.bbox_cxcyah_to_xyxy
    sys.modules['mmtrack.structures.bbox'].bbox_xyxy_to_cxcyah = tr.bbox_xyxy_to_cxcyah
    _pkg('mmtrack.models', os.path.join(mm, 'models'))
    _pkg('mmtrack.models.trackers', os.path.join(mm, 'models', 'trackers'))
    _pkg('mmtrack.models.motion', os.path.join(mm, 'models', 'motion'))
    _pkg('mmtrack.models.mot', os.path.join(mm, 'models', 'mot'))
    _mod('mmtrack.models.mot.ocsort', OCSORT=type('OCSORT', (), {}))
    trk = importlib.import_module('mmtrack.models.trackers.ocsort_tracker_disparity')
    kfm = importlib.import_module('mmtrack.models.motion.kalman_filter')
    mot = importlib.import_module('mmtrack.models.mot.ocsort_disparity')
    utl = importlib.import_module('mmtrack.models.trackers.utils')

    class KalmanFilterNp1(kfm.KalmanFilter):
        """NumPy-1.x promotion inside initiate (see module docstring)."""

        def initiate(self, measurement):
            return super().initiate(np.asarray(measurement, np.float64))

    ns = types.SimpleNamespace(
        OCSORTTracker_Disparity=trk.OCSORTTracker_Disparity,
        KalmanFilter=KalmanFilterNp1, KalmanFilterRaw=kfm.KalmanFilter,
        OCSORT_Disparity=mot.OCSORT_Disparity, scale_bbox=utl.scale_bbox,
        bbox_xyxy_to_cxcyah=tr.bbox_xyxy_to_cxcyah,
        bbox_cxcyah_to_xyxy=tr.bbox_cxcyah_to_xyxy,
        InstanceData=InstanceData, TrackDataSample=TrackDataSample)
    _LOADED = ns
    return ns


class FakeDetector:
    """Returns canned detections from ``self.next`` (InstanceData)."""

    def predict(self, data, data_samples):
        return [types.SimpleNamespace(pred_instances=self.next)]


def build_model(ns, tracker_cfg, baseline=0.25, focal_length=640):
    """An OCSORT_Disparity instance without running mmengine's BaseModel
    constructor: the attributes predict() reads are set directly."""
    model = object.__new__(ns.OCSORT_Disparity)
    model.detector = FakeDetector()
    model.tracker = ns.OCSORTTracker_Disparity(**tracker_cfg)
    model.motion = ns.KalmanFilter()
    model.baseline = baseline
    model.focal_length = focal_length
    return model
