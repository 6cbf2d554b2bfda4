"""The drop-in boundary on the GPU: the classes registered under the reference's names
(OCSORT_Disparity / OCSORTTracker_Disparity / KalmanFilter), built through the registry
from the keys of configs/stereo_tracking/ocsort/yolox_s_mmyolo_mot_airdrone_disp.py:29-58
and driven exactly as the reference drives its own (model.predict on TrackDataSample),
replay the fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

# model block of the shipped config (:29-58), detector replaced by a stand-in
TRACKER_CFG = dict(type='OCSORTTracker_Disparity', obj_score_thr=0.3, init_track_thr=0.7,
                   weight_iou_with_det_scores=False, match_iou_thr=0.1, num_tentatives=3,
                   vel_consist_weight=0.2, vel_delta_t=3, num_frames_retain=30)


class _Det:
    """Stands where the reference PyTorch detector stays: returns the fixture's boxes."""

    def __init__(self, api):
        self.api, self.next = api, None

    def predict(self, data, data_samples, **kw):
        class R:
            pass
        r = R()
        r.pred_instances = self.next
        return [r]


@pytest.fixture(scope='module')
def api():
    import torch
    from stereotracking_b200 import compat, mmtrack_api
    from stereotracking_b200 import mmtrack_shell  # noqa: F401  (opt-in: registers OCSORT / OCSORT_Disparity)
    from oracle import disparity as odisp
    return dict(torch=torch, compat=compat, api=mmtrack_api, odisp=odisp)


@pytest.mark.parametrize('name', ['basic', 'occlusion', 'cmc', 'defaults', 'median'])
def test_model_predict_replays_reference_fixture(api, name):
    torch, compat = api['torch'], api['compat']
    g = np.load(os.path.join(GOLDEN, f'ref_track_{name}.npz'))
    det = _Det(api)
    extra = dict(cmc=dict(method='glme_affine')) if name == 'cmc' else {}
    base = dict(type='OCSORTTracker_Disparity') if name == 'defaults' else TRACKER_CFG     # ctor defaults
    # 'median': the reference run had extract_depth decorated with median_decorator
    # (mmtrack/models/mot/depth_extraction_comparison.py:83-115)
    shell = dict(depth_reducer='median') if name == 'median' else {}
    model = compat.MODELS.build(dict(type='OCSORT_Disparity', detector=det, motion=dict(type='KalmanFilter'),
                                     tracker=dict(base, max_dets=64, max_tracks=128, **extra), **shell))
    assert type(model.tracker).__name__ == 'OCSORTTracker_Disparity' and model.tracker.empty
    if name == 'cmc':
        # as in the fixture: the host-side estimator is replaced by the scripted warps, the
        # tracker applies them (with_cmc path of track())
        assert model.tracker.with_cmc
        model.tracker.estimate_camera_motion = lambda img, meta: (
            None if np.isnan(g['warp'][meta['frame_id'], 0]) else g['warp'][meta['frame_id']].reshape(2, 3))
    for f in range(len(g['raw'])):
        n = int(g['det_count'][f])
        pre = api['odisp'].preprocess(g['raw'][f])
        disp = torch.from_numpy(pre['disp'])[None, None, None].repeat(1, 1, 3, 1, 1).cuda()
        mask = torch.from_numpy(pre['mask'])[None, None, None].cuda()
        det.next = compat.InstanceData(bboxes=torch.from_numpy(g['det_boxes'][f, :n].copy()).cuda(),
                                       scores=torch.from_numpy(g['det_scores'][f, :n].copy()).cuda(),
                                       labels=torch.zeros(n, dtype=torch.int64).cuda())
        sample = compat.TrackDataSample(metainfo=dict(frame_id=f))
        out = model.predict(dict(img=torch.zeros_like(disp), disp_postp=disp, disp_mask=mask), [sample])[0]
        ti = out.pred_track_instances
        m = int(g['out_count'][f])
        assert ti.instances_id.dtype == torch.int64 and ti.bboxes.is_cuda
        np.testing.assert_array_equal(ti.instances_id.cpu().numpy(), g['out_ids'][f, :m], err_msg=f'frame {f}')
        np.testing.assert_array_equal(ti.labels.cpu().numpy(), g['out_labels'][f, :m])
        np.testing.assert_array_equal(ti.scores.cpu().numpy(), g['out_scores'][f, :m])
        np.testing.assert_allclose(ti.bboxes.cpu().numpy(), g['out_bboxes'][f, :m], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(ti.scales.cpu().numpy(), g['out_scales'][f, :m], rtol=2e-6)
        np.testing.assert_allclose(np.asarray(ti['depth'], np.float32), g['out_depth'][f, :m], rtol=1e-5,
                                   equal_nan=True)
        assert ti['gt_depth'] == ti['depth']                       # ocsort_disparity.py:103-104 fallback
        # raw detections are handed back untouched (:107)
        np.testing.assert_array_equal(out.pred_det_instances.bboxes.cpu().numpy(), g['det_boxes'][f, :n])
        # BaseTracker surface
        k = int(g['st_count'][f])
        assert model.tracker.ids == g['out_st_ids'][f, :k].tolist()
        assert model.tracker.confirmed_ids == g['out_st_ids'][f, :k][~g['out_st_tent'][f, :k]].tolist()
        assert model.tracker.num_tracks == int(g['num_tracks'][f])
    model.tracker.reset()
    assert model.tracker.empty and model.tracker.num_tracks == 0


def test_tracker_memo_and_get_match_reference_fixture(api):
    """BaseTracker.memo / get (base_tracker.py:120-177) while replaying the occlusion fixture."""
    torch, compat = api['torch'], api['compat']
    g = np.load(os.path.join(GOLDEN, 'ref_track_occlusion.npz'))
    mm = np.load(os.path.join(GOLDEN, 'ref_memo_occlusion.npz'))
    det = _Det(api)
    model = compat.MODELS.build(dict(type='OCSORT_Disparity', detector=det, motion=dict(type='KalmanFilter'),
                                     tracker=dict(TRACKER_CFG, max_dets=64, max_tracks=128)))
    probes = set(int(v) for v in mm['probes'])
    for f in range(len(g['raw'])):
        n = int(g['det_count'][f])
        pre = api['odisp'].preprocess(g['raw'][f])
        disp = torch.from_numpy(pre['disp'])[None, None, None].repeat(1, 1, 3, 1, 1).cuda()
        mask = torch.from_numpy(pre['mask'])[None, None, None].cuda()
        det.next = compat.InstanceData(bboxes=torch.from_numpy(g['det_boxes'][f, :n].copy()).cuda(),
                                       scores=torch.from_numpy(g['det_scores'][f, :n].copy()).cuda(),
                                       labels=torch.zeros(n, dtype=torch.int64).cuda())
        model.predict(dict(img=torch.zeros_like(disp), disp_postp=disp, disp_mask=mask),
                      [compat.TrackDataSample(metainfo=dict(frame_id=f))])
        if f not in probes:
            continue
        trk = model.tracker
        memo = trk.memo
        for k in ('ids', 'labels', 'frame_ids'):
            np.testing.assert_array_equal(memo[k].cpu().numpy(), mm[f'memo_{k}_{f}'], err_msg=f'{k} frame {f}')
        np.testing.assert_array_equal(memo.scores.cpu().numpy(), mm[f'memo_scores_{f}'])
        np.testing.assert_allclose(memo['bboxes'].cpu().numpy(), mm[f'memo_bboxes_{f}'], rtol=1e-5, atol=1e-4)
        np.testing.assert_allclose(memo['scales'].cpu().numpy(), mm[f'memo_scales_{f}'], rtol=2e-6)
        np.testing.assert_allclose(memo['depth'].cpu().numpy(), mm[f'memo_depth_{f}'], rtol=1e-5, equal_nan=True)
        np.testing.assert_allclose(trk.get('bboxes', num_samples=3, behavior='mean').cpu().numpy(),
                                   mm[f'get_bboxes_mean3_{f}'], rtol=1e-5, atol=1e-4)
        np.testing.assert_array_equal(trk.get('scores').cpu().numpy(), mm[f'get_scores_last_{f}'])
        full = mm[f'get_full_{f}'].tolist()
        if full:
            got = trk.get('bboxes', ids=full, num_samples=2)
            assert tuple(got.shape) == mm[f'get_bboxes_none2_{f}'].shape
            np.testing.assert_allclose(got.cpu().numpy(), mm[f'get_bboxes_none2_{f}'], rtol=1e-5, atol=1e-4)
        with pytest.raises(NotImplementedError):
            trk.get('bboxes', num_samples=2, behavior='median')
    model.tracker.reset()
    assert model.tracker._frames == []


def test_kalman_filter_object_matches_reference_fixture(api):
    g = np.load(os.path.join(GOLDEN, 'ref_kf.npz'))
    kf = api['compat'].TASK_UTILS.build(dict(type='KalmanFilter'))
    keys = set(g.files)
    assert {'mean', 'cov', 'z', 'op'} <= keys, keys
    mean, cov = None, None
    for i in range(len(g['op'])):
        op = int(g['op'][i])
        if op == 0:
            mean, cov = kf.initiate(g['z'][i].astype(np.float32))
        elif op == 1:
            mean, cov = kf.predict(mean, cov)
        else:
            mean, cov = kf.update(mean, cov, g['z'][i].astype(np.float32))
        np.testing.assert_allclose(mean, g['mean'][i], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(cov, g['cov'][i], rtol=1e-9, atol=1e-12)


def test_kalman_filter_project_gating_track_match_reference_fixture(api):
    """KalmanFilter.project / gating_distance / track (kalman_filter.py:121-152,:190-260)."""
    import types
    g = np.load(os.path.join(GOLDEN, 'ref_kf_gating.npz'))
    k = np.load(os.path.join(GOLDEN, 'ref_kf.npz'))
    kf = api['compat'].TASK_UTILS.build(dict(type='KalmanFilter'))
    for n, i in enumerate(g['idx']):
        pm, pc = kf.project(k['mean'][i], k['cov'][i])
        np.testing.assert_array_equal(pm, g['pmean'][n])
        np.testing.assert_allclose(pc, g['pcov'][n], rtol=1e-15, atol=0)
        np.testing.assert_allclose(kf.gating_distance(k['mean'][i], k['cov'][i], g['meas'][n]), g['dist4'][n],
                                   rtol=1e-10)
        np.testing.assert_allclose(kf.gating_distance(k['mean'][i], k['cov'][i], g['meas'][n], only_position=True),
                                   g['dist2'][n], rtol=1e-10)
    a, b = (int(v) for v in g['track_in'])
    tracks = {7: types.SimpleNamespace(mean=k['mean'][a].copy(), covariance=k['cov'][a].copy()),
              9: types.SimpleNamespace(mean=k['mean'][b].copy(), covariance=k['cov'][b].copy())}
    out, costs = kf.track(tracks, api['torch'].from_numpy(g['track_boxes']).cuda())
    np.testing.assert_array_equal(np.isnan(costs), np.isnan(g['track_costs']))
    np.testing.assert_allclose(costs, g['track_costs'], rtol=1e-10, equal_nan=True)
    np.testing.assert_allclose(np.stack([out[7].mean, out[9].mean]), g['track_mean'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.stack([out[7].covariance, out[9].covariance]), g['track_cov'], rtol=1e-11, atol=1e-13)
    assert not np.isnan(g['track_costs']).all() and np.isnan(g['track_costs']).any()


def test_kalman_filter_use_nsa_matches_reference_fixture(api):
    """KalmanFilter(use_nsa=True) (kalman_filter.py:38,:144-145): project / update with bbox scores
    against the unmodified reference; with score 0 (what the OC-SORT trackers pass) it is the
    plain filter."""
    g = np.load(os.path.join(GOLDEN, 'ref_kf_nsa.npz'))
    k = np.load(os.path.join(GOLDEN, 'ref_kf.npz'))
    kf = api['compat'].TASK_UTILS.build(dict(type='KalmanFilter', use_nsa=True))
    plain = api['compat'].TASK_UTILS.build(dict(type='KalmanFilter'))
    assert kf.use_nsa and not plain.use_nsa
    for n, i in enumerate(g['idx']):
        sc = float(g['score'][n])
        pm, pc = kf.project(k['mean'][i], k['cov'][i], sc)
        np.testing.assert_array_equal(pm, g['pmean'][n])
        np.testing.assert_allclose(pc, g['pcov'][n], rtol=1e-15, atol=0)
        m, c = kf.update(k['mean'][i], k['cov'][i], g['z'][n], sc)
        np.testing.assert_allclose(m, g['umean'][n], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(c, g['ucov'][n], rtol=1e-9, atol=1e-12)
    m0, c0 = kf.update(k['mean'][3], k['cov'][3], g['z'][0])                 # bbox_score defaults to 0
    m1, c1 = plain.update(k['mean'][3], k['cov'][3], g['z'][0])
    np.testing.assert_array_equal(m0, m1)
    np.testing.assert_array_equal(c0, c1)


def test_cmc_state_resets_with_every_new_video(api):
    """frame_id == 0 resets tracks AND camera-motion state before anything else
    (ocsort_tracker_disparity.py:385-391): the second video on a tracker instance must run exactly
    like the same video on a fresh instance - in particular no warp is estimated on its frames 0
    and 1 (the estimator first needs a previous image from an association-branch frame)."""
    torch, compat = api['torch'], api['compat']
    g = np.load(os.path.join(GOLDEN, 'ref_track_cmc.npz'))
    F = 12

    def run(tracker, calls):
        # a deterministic stand-in for the OpenCV estimator body; the bookkeeping of
        # estimate_camera_motion (prev_cmc_img, :87-97) stays the tracker's own
        def fake(img, prev, meta, **kw):
            calls.append(meta['frame_id'])
            return np.asarray([[1.01, 0.002, 1.5], [-0.002, 1.01, -0.8]], np.float32)
        api['api']._mesh_affine, keep = fake, api['api']._mesh_affine
        try:
            outs = []
            for f in range(F):
                n = int(g['det_count'][f])
                r = tracker._track(torch.full((1, 3, 8, 8), float(f)), dict(frame_id=f, img_shape=(8, 8)),
                                   torch.from_numpy(g['det_boxes'][f, :n].copy()).cuda(),
                                   torch.from_numpy(g['det_scores'][f, :n].copy()).cuda(),
                                   torch.zeros(n, dtype=torch.int64).cuda(), torch.ones(n).cuda(), torch.ones(n).cuda())
                outs.append((r.instances_id.cpu().numpy(), r.bboxes.cpu().numpy()))
            return outs, tracker._engine.export_state(0)
        finally:
            api['api']._mesh_affine = keep

    cfg = dict(TRACKER_CFG, cmc=dict(method='glme_affine'), max_dets=64, max_tracks=128)
    reused, fresh = compat.MODELS.build(dict(cfg)), compat.MODELS.build(dict(cfg))
    run(reused, [])                                                   # video A
    calls_b, calls_f = [], []
    out_b, st_b = run(reused, calls_b)                                # video B on the same instance
    out_f, st_f = run(fresh, calls_f)                                 # video B on a fresh instance
    assert calls_b == calls_f and 0 not in calls_b and 1 not in calls_b and 2 in calls_b
    for (ia, ba), (ib, bb) in zip(out_b, out_f):
        np.testing.assert_array_equal(ia, ib)
        np.testing.assert_array_equal(ba, bb)
    np.testing.assert_array_equal(st_b['mean'], st_f['mean'])
    np.testing.assert_array_equal(st_b['cov'], st_f['cov'])


def test_constructor_errors_follow_the_reference(api):
    T = api['api'].OCSORTTracker_Disparity
    with pytest.raises(ValueError):
        T(cmc=dict(method='ecc'))                                  # :74-75
    assert T(cmc=dict(method='glme_affine')).with_cmc and not T().with_cmc
    with pytest.raises(AssertionError):
        T(momentums=[0.1])                                         # base_tracker.py:26-27
    assert api['api'].KalmanFilter(use_nsa=True).use_nsa            # kalman_filter.py:38
    with pytest.raises(ValueError):
        api['api'].OCSORT_Disparity(depth_reducer='mode')           # not one of depth_extraction_comparison.py's


def test_rgb_only_ocsort_matches_oracle_tracker(api):
    """OCSORT / OCSORTTracker (the RGB-only pair, ocsort.py / ocsort_tracker.py): the
    same association with scales = 1, against the oracle tracker on jittered boxes."""
    torch, compat = api['torch'], api['compat']
    from oracle.tracker import OCSortDisparityOracle
    g = np.load(os.path.join(GOLDEN, 'ref_track_occlusion.npz'))
    cfg = {k: v for k, v in TRACKER_CFG.items() if k != 'type'}
    det = _Det(api)
    model = compat.MODELS.build(dict(type='OCSORT', detector=det, motion=dict(type='KalmanFilter'),
                                     tracker=dict(type='OCSORTTracker', max_dets=64, max_tracks=128, **cfg)))
    ref = OCSortDisparityOracle(**cfg)
    rows = 0
    for f in range(len(g['raw'])):
        n = int(g['det_count'][f])
        boxes, scores = g['det_boxes'][f, :n].copy(), g['det_scores'][f, :n].copy()
        det.next = compat.InstanceData(bboxes=torch.from_numpy(boxes).cuda(), scores=torch.from_numpy(scores).cuda(),
                                       labels=torch.zeros(n, dtype=torch.int64).cuda())
        sample = compat.TrackDataSample(metainfo=dict(frame_id=f))
        out = model.predict(dict(img=torch.zeros(1, 1, 3, 32, 32).cuda()), [sample])[0].pred_track_instances
        want = ref.track(f, boxes, scores, np.zeros(n, np.int64), np.ones(n, np.float32), np.zeros(n, np.float32))
        assert sorted(out.keys()) == ['bboxes', 'instances_id', 'labels', 'scores']
        np.testing.assert_array_equal(out.instances_id.cpu().numpy(), want['instances_id'], err_msg=f'frame {f}')
        np.testing.assert_array_equal(out.bboxes.cpu().numpy(), want['bboxes'])
        np.testing.assert_array_equal(out.scores.cpu().numpy(), want['scores'])
        rows += len(want['instances_id'])
    assert rows > 100



def test_predict_with_gt_depth_map_fills_gt_depth(api):
    """inputs['depth_postp'] (a ground-truth depth map, ocsort_disparity.py:56-59,:100-104):
    gt_depth of every track is the same box statistic taken from that map."""
    torch, compat = api['torch'], api['compat']
    from oracle import boxdepth
    g = np.load(os.path.join(GOLDEN, 'ref_track_basic.npz'))
    det = _Det(api)
    model = compat.MODELS.build(dict(type='OCSORT_Disparity', detector=det, motion=dict(type='KalmanFilter'),
                                     tracker=dict(TRACKER_CFG, max_dets=64, max_tracks=128)))
    rng = np.random.default_rng(9)
    checked = 0
    for f in range(6):
        n = int(g['det_count'][f])
        pre = api['odisp'].preprocess(g['raw'][f])
        gt_map = rng.uniform(1.0, 120.0, pre['depth'].shape).astype(np.float32)
        disp = torch.from_numpy(pre['disp'])[None, None, None].repeat(1, 1, 3, 1, 1).cuda()
        det.next = compat.InstanceData(bboxes=torch.from_numpy(g['det_boxes'][f, :n].copy()).cuda(),
                                       scores=torch.from_numpy(g['det_scores'][f, :n].copy()).cuda(),
                                       labels=torch.zeros(n, dtype=torch.int64).cuda())
        sample = compat.TrackDataSample(metainfo=dict(frame_id=f))
        out = model.predict(dict(img=torch.zeros_like(disp), disp_postp=disp,
                                 disp_mask=torch.ones(1, 1, 1, *pre['depth'].shape).cuda(),
                                 depth_postp=torch.from_numpy(gt_map)[None, None, None].cuda()), [sample])[0]
        ti = out.pred_track_instances
        want, _ = boxdepth.extract_depth(gt_map, ti.bboxes.cpu().numpy())
        np.testing.assert_allclose(np.asarray(ti['gt_depth'], np.float32), np.asarray(want, np.float32), rtol=1e-5,
                                   equal_nan=True)
        m = int(g['out_count'][f])
        np.testing.assert_allclose(np.asarray(ti['depth'], np.float32), g['out_depth'][f, :m], rtol=1e-5, equal_nan=True)
        checked += m
    assert checked > 20
