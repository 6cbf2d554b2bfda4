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
    # importing mmtrack_api swaps the trackers and the Kalman filter only ...
    for name in ('OCSORTTracker_Disparity', 'OCSORTTracker'):
        assert compat.MODELS.get(name) is not None, name
    assert compat.TASK_UTILS.get('KalmanFilter') is not None
    # ... the model shells are an explicit opt-in (stereotracking_b200.mmtrack_shell)
    from stereotracking_b200 import mmtrack_shell
    assert compat.MODELS.get('OCSORT_Disparity') is mmtrack_shell.OCSORT_Disparity
    assert compat.MODELS.get('OCSORT') is mmtrack_shell.OCSORT
    base = compat.model_shell_base()
    assert issubclass(mmtrack_shell.OCSORT_Disparity, mmtrack_api._ShellBase) and base.__name__ == 'BaseMultiObjectTracker'
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


def test_model_shell_is_an_mmengine_base_model_under_a_real_registry(monkeypatch):
    """With mmengine / mmtrack importable the shells derive from the reference's own
    BaseMultiObjectTracker(BaseModel) (mot/base.py:12) and hand ``data_preprocessor`` / ``init_cfg``
    to it, so a Runner can ``.to()`` the model and call ``test_step``.  mmengine is not in this
    image: minimal fakes of the four modules compat imports are put in sys.modules, the shipped
    config block (configs/stereo_tracking/ocsort/yolox_s_mmyolo_mot_airdrone_disp.py:29-58) is
    built through the fake registry, and the modules are restored afterwards."""
    import importlib
    import sys
    import types

    if compat.HAVE_MMTRACK:
        pytest.skip('real mmengine in use')

    class FakeBaseModel(torch.nn.Module):                      # mmengine.model.BaseModel surface
        def __init__(self, data_preprocessor=None, init_cfg=None):
            super().__init__()
            if isinstance(data_preprocessor, dict):
                data_preprocessor = FAKE_MODELS.build(data_preprocessor)
            self.data_preprocessor, self.init_cfg = data_preprocessor, init_cfg

        def test_step(self, data):
            data = self.data_preprocessor(data, False)
            return self(**data, mode='predict')

    class FakeBaseMOT(FakeBaseModel):                          # mmtrack.models.mot.base.BaseMultiObjectTracker
        def forward(self, inputs, data_samples=None, mode='predict', **kw):
            assert mode == 'predict'
            return self.predict(inputs, data_samples, **kw)

    FAKE_MODELS, FAKE_TASKS = compat.Registry('model'), compat.Registry('task util')
    fakes = {
        'mmengine': types.ModuleType('mmengine'),
        'mmengine.structures': types.ModuleType('mmengine.structures'),
        'mmtrack': types.ModuleType('mmtrack'),
        'mmtrack.registry': types.ModuleType('mmtrack.registry'),
        'mmtrack.structures': types.ModuleType('mmtrack.structures'),
        'mmtrack.models': types.ModuleType('mmtrack.models'),
        'mmtrack.models.mot': types.ModuleType('mmtrack.models.mot'),
        'mmtrack.models.mot.base': types.ModuleType('mmtrack.models.mot.base'),
    }
    fakes['mmengine.structures'].InstanceData = compat.InstanceData
    fakes['mmtrack.registry'].MODELS, fakes['mmtrack.registry'].TASK_UTILS = FAKE_MODELS, FAKE_TASKS
    fakes['mmtrack.structures'].TrackDataSample = compat.TrackDataSample
    fakes['mmtrack.models.mot.base'].BaseMultiObjectTracker = FakeBaseMOT
    for k, v in fakes.items():
        monkeypatch.setitem(sys.modules, k, v)
    from stereotracking_b200 import engine, mmtrack_api, mmtrack_shell
    try:
        importlib.reload(compat)
        assert compat.HAVE_MMTRACK and compat.MODELS is FAKE_MODELS
        importlib.reload(mmtrack_api)

        class FakeEngine:                                      # no GPU in this test: the tracker's engine is not under test
            DEPTH_MODES = engine.StereoTrackEngine.DEPTH_MODES

            def __init__(self, cfg, device='cuda:0'):
                self.cfg = cfg

        monkeypatch.setattr(mmtrack_api, 'StereoTrackEngine', FakeEngine)
        assert FAKE_MODELS.get('OCSORT_Disparity') is None     # not registered by the tracker-only import
        importlib.reload(mmtrack_shell)
        assert issubclass(mmtrack_shell.OCSORT_Disparity, FakeBaseMOT) and issubclass(mmtrack_shell.OCSORT, FakeBaseMOT)

        @FAKE_MODELS.register_module()
        class TrackDataPreprocessor_Disparity_V1(torch.nn.Module):
            def __init__(self, pad_size_divisor=32):
                super().__init__()
                self.pad_size_divisor = pad_size_divisor

            def forward(self, data, training=False):
                return dict(inputs=data['inputs'], data_samples=data['data_samples'])

        @FAKE_MODELS.register_module()
        class YOLODetector_Disparity_V1(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.w = torch.nn.Parameter(torch.zeros(1))

        model = FAKE_MODELS.build(dict(
            type='OCSORT_Disparity',
            data_preprocessor=dict(type='TrackDataPreprocessor_Disparity_V1', pad_size_divisor=32),
            detector=dict(type='YOLODetector_Disparity_V1'),
            motion=dict(type='KalmanFilter'),
            tracker=dict(type='OCSORTTracker_Disparity', obj_score_thr=0.3, init_track_thr=0.7,
                         weight_iou_with_det_scores=False, match_iou_thr=0.1, num_tentatives=3,
                         vel_consist_weight=0.2, vel_delta_t=3, num_frames_retain=30)))
        assert isinstance(model, FakeBaseModel) and isinstance(model, torch.nn.Module)
        assert isinstance(model.data_preprocessor, TrackDataPreprocessor_Disparity_V1)   # not dropped
        assert model.to('cpu') is model and model.eval() is model
        assert [n for n, _ in model.named_parameters()] == ['detector.w']               # the detector is a sub-module
        assert model.tracker.match_iou_thr == 0.1 and model.tracker._engine.cfg.focal_length == 640.0
        seen = {}
        monkeypatch.setattr(type(model), 'predict', lambda self, inputs, data_samples, **kw: seen.update(
            inputs=inputs, n=len(data_samples)) or ['ok'])
        assert model.test_step(dict(inputs=dict(img=1), data_samples=[object()])) == ['ok'] and seen['n'] == 1
    finally:
        monkeypatch.undo()
        importlib.reload(compat)
        importlib.reload(mmtrack_api)
        importlib.reload(mmtrack_shell)
        importlib.reload(engine)
    assert not compat.HAVE_MMTRACK
