"""Host-side stand-ins used when mmengine / mmtrack are not importable (compat.py) and the
registration of the drop-in classes under the reference's type names (CPU only)."""
import numpy as np
import pytest
import torch

from stereotracking_b200 import compat


def test_registry_register_build_and_force():
    if compat.HAVE_MMTRACK:
        pytest.skip('real mmengine registries in use')
    reg = compat.Registry('unit')

    @reg.register_module()
    class A:
        def __init__(self, x=1, y=2):
            self.x, self.y = x, y

    obj = reg.build(dict(type='A', x=5), y=7)
    assert isinstance(obj, A) and (obj.x, obj.y) == (5, 7)
    assert reg.build(dict(type='mmtrack.A')).x == 1            # scoped names resolve to the last part
    with pytest.raises(KeyError):
        reg.register_module()(A)                               # duplicate without force
    reg.register_module(force=True)(A)
    with pytest.raises(KeyError):
        reg.build(dict(type='Missing'))


def test_drop_in_classes_are_registered_under_the_reference_names():
    from stereotracking_b200 import mmtrack_api  # noqa: F401  (registers on import, no GPU needed)
    for name in ('OCSORTTracker_Disparity', 'OCSORT_Disparity', 'OCSORTTracker', 'OCSORT'):
        assert compat.MODELS.get(name) is not None, name
    assert compat.TASK_UTILS.get('KalmanFilter') is not None
    # constructor keywords and defaults of ocsort_tracker_disparity.py:42-52
    import inspect
    sig = inspect.signature(compat.MODELS.get('OCSORTTracker_Disparity').__init__).parameters
    want = dict(obj_score_thr=0.3, init_track_thr=0.7, weight_iou_with_det_scores=True, match_iou_thr=0.3,
                num_tentatives=3, vel_consist_weight=0.2, vel_delta_t=3, cmc=None, num_frames_retain=10)
    for k, v in want.items():
        assert sig[k].default == v, k
    sig = inspect.signature(compat.MODELS.get('OCSORT_Disparity').__init__).parameters
    assert sig['baseline'].default == 0.25 and sig['focal_length'].default == 640      # ocsort_disparity.py:36-37


def test_instance_data_and_track_data_sample_stand_ins():
    if compat.HAVE_MMTRACK:
        pytest.skip('real mmengine structures in use')
    d = compat.InstanceData(bboxes=torch.arange(12.).reshape(3, 4), scores=torch.tensor([.9, .5, .1]))
    d.labels = torch.zeros(3, dtype=torch.int64)
    d['depth'] = [1.0, 2.0, 3.0]
    assert len(d) == 3 and 'labels' in d and sorted(d.keys()) == ['bboxes', 'depth', 'labels', 'scores']
    sel = compat.InstanceData(bboxes=d.bboxes, scores=d.scores)[d.scores > 0.3]
    assert len(sel) == 2 and torch.equal(sel.bboxes, d.bboxes[:2])
    c = d.clone()
    c.bboxes[0, 0] = -1
    assert d.bboxes[0, 0] == 0 and d.get('missing', 7) == 7
    with pytest.raises(AttributeError):
        d.nothing
    s = compat.TrackDataSample(metainfo=dict(frame_id=4))
    assert s.frame_id == 4 and s.metainfo['frame_id'] == 4
    s.set_metainfo(dict(img_shape=(720, 1280)))
    assert s.metainfo['img_shape'] == (720, 1280)
    s.pred_track_instances = d
    assert s.pred_track_instances is d


def test_tracker_needs_a_gpu_and_says_so():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from stereotracking_b200 import _lib, mmtrack_api
    with pytest.raises(_lib.StereoTrackError):
        mmtrack_api.OCSORTTracker_Disparity()
