"""The oracle against the fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import boxdepth, disparity, kalman
from oracle.model import StereoTrackOracle
from oracle.tracker import cxcyah_to_xyxy, xyxy_to_cxcyah


def test_kalman_chain_matches_reference():
    g = np.load(os.path.join(GOLDEN, 'ref_kf.npz'))
    m = c = None
    for i, op in enumerate(g['op']):
        if op == 0:
            m, c = kalman.initiate(g['z'][i].astype(np.float32))
        elif op == 1:
            m, c = kalman.predict(m, c)
        else:
            m, c = kalman.update(m, c, g['z'][i].astype(np.float32))
        np.testing.assert_allclose(m, g['mean'][i], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(c, g['cov'][i], rtol=1e-11, atol=1e-13)


def test_kalman_project_and_gating_match_reference():
    g = np.load(os.path.join(GOLDEN, 'ref_kf_gating.npz'))
    k = np.load(os.path.join(GOLDEN, 'ref_kf.npz'))
    for n, i in enumerate(g['idx']):
        pm, pc = kalman.project(k['mean'][i], k['cov'][i])
        np.testing.assert_array_equal(pm, g['pmean'][n])
        np.testing.assert_allclose(pc, g['pcov'][n], rtol=1e-15, atol=0)
        np.testing.assert_allclose(kalman.gating_distance(k['mean'][i], k['cov'][i], g['meas'][n]), g['dist4'][n],
                                   rtol=1e-12)
        np.testing.assert_allclose(kalman.gating_distance(k['mean'][i], k['cov'][i], g['meas'][n], True),
                                   g['dist2'][n], rtol=1e-12)


def test_boxdepth_matches_reference():
    g = np.load(os.path.join(GOLDEN, 'ref_boxdepth.npz'))
    pre = disparity.preprocess(g['raw'])
    np.testing.assert_array_equal(pre['depth'], g['depth_map'])
    d, s = boxdepth.extract_depth(pre['depth'], g['boxes'])
    np.testing.assert_array_equal(d, g['d'])          # NaN == NaN positions too
    np.testing.assert_array_equal(s, g['scales'])
    assert np.isnan(g['d']).any() and (g['d'] == -1).any()
    infl = boxdepth.scale_bbox(g['boxes'], s)
    np.testing.assert_array_equal(infl, g['inflated'])
    with np.errstate(invalid='ignore', divide='ignore'):
        defl = boxdepth.scale_bbox(infl, np.float32(1) / s)
    np.testing.assert_array_equal(defl, g['deflated'])
    with np.errstate(invalid='ignore', divide='ignore'):
        cx = np.stack([xyxy_to_cxcyah(b) for b in g['boxes']])
    np.testing.assert_array_equal(cx, g['cxcyah'])
    np.testing.assert_array_equal(cxcyah_to_xyxy(cx), g['xyxy_back'])


def test_pack_image_matches_reference_lines():
    """RGB half of the detector input (pad 114, HWC->CHW, swap, float, normalise, stack_batch)."""
    g = np.load(os.path.join(GOLDEN, 'ref_frontend.npz'))
    for tag in 'ab':
        np.testing.assert_array_equal(disparity.pack_image(g[f'img_{tag}']), g[f'out_plain_{tag}'])
        np.testing.assert_array_equal(
            disparity.pack_image(g[f'img_{tag}'], mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375),
                                 swap_rb=True), g[f'out_norm_{tag}'])
    assert (g['out_plain_a'][:, 40:, :] == 114).all() and (g['out_plain_a'][:, :, 52:] == 114).all()


def test_bilinear_resize_matches_opencv_vectors():
    """Resize_Disparity's image branch (mmcv.imrescale 'bilinear' = cv2.resize INTER_LINEAR), bit-exact, and
    the packed detector input behind it."""
    g = np.load(os.path.join(GOLDEN, 'ref_resize_bilinear.npz'))
    for tag in ('down', 'up', 'half', 'odd', 'tall'):
        img, scale, ref = g[f'img_{tag}'], tuple(int(v) for v in g[f'scale_{tag}']), g[f'out_{tag}']
        nh, nw = disparity.rescale_size(img.shape[0], img.shape[1], scale)
        assert (nh, nw) == ref.shape[:2], tag
        np.testing.assert_array_equal(disparity.bilinear_resize_u8(img, nh, nw), ref)
        np.testing.assert_array_equal(disparity.pack_image(img, img_scale=scale), disparity.pack_image(ref))
    one = g['img_down'][:, :, 0]
    np.testing.assert_array_equal(disparity.bilinear_resize_u8(one, 36, 64), g['out_down'][:, :, 0])


@pytest.mark.parametrize('mode', boxdepth.ABLATION_MODES)
def test_boxdepth_ablation_matches_reference(mode):
    """The four reducers of depth_extraction_comparison.py, box by box (center raises
    IndexError exactly where the reference did)."""
    g = np.load(os.path.join(GOLDEN, 'ref_boxdepth_ablation.npz'))
    seen = dict(minus1=0, nan=0, oob=0, clipped=0)
    for tag in 'ab':
        depth = disparity.preprocess(g[f'raw_{tag}'])['depth']
        for i, box in enumerate(g[f'boxes_{tag}']):
            if g[f'oob_{mode}_{tag}'][i]:
                with pytest.raises(IndexError):
                    boxdepth.box_depth_ablation(depth, box, mode)
                seen['oob'] += 1
                continue
            with np.errstate(all='ignore'):
                d, s = boxdepth.box_depth_ablation(depth, box, mode)
            np.testing.assert_array_equal(np.float32(d), g[f'd_{mode}_{tag}'][i], err_msg=f'{tag}{i}')
            np.testing.assert_array_equal(np.float32(s), g[f'scale_{mode}_{tag}'][i], err_msg=f'{tag}{i}')
            seen['minus1'] += int(d == -1)
            seen['nan'] += int(np.isnan(np.float32(d)))
            seen['clipped'] += int(np.float32(s) == 3)
    assert seen['minus1'] > 0 and (seen['clipped'] > 0 or mode == 'median')
    assert (seen['oob'] > 0) == (mode == 'center')
    assert (seen['nan'] > 0) == (mode == 'truncated_mean')     # n == 1: d_sorted[:-1] is empty


@pytest.mark.parametrize('name', ['basic', 'occlusion', 'stress', 'cmc', 'defaults', 'median'])
def test_tracker_sequence_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, f'ref_track_{name}.npz'))
    # 'defaults': the constructor defaults of OCSORTTracker_Disparity, not the AirDrone config
    # 'median': an ablation run, extract_depth decorated with median_decorator
    o = StereoTrackOracle(tracker_cfg={}) if name == 'defaults' else StereoTrackOracle(
        depth_reducer='median' if name == 'median' else None)
    lost_recovered = 0
    for f in range(len(g['raw'])):
        n = int(g['det_count'][f])
        pre = o.depth_map(g['raw'][f])
        # 'cmc': the scripted camera-motion affine of the frame (NaN row = the estimator returned None)
        warp = None
        if 'warp' in g.files and not np.isnan(g['warp'][f, 0]):
            warp = g['warp'][f].reshape(2, 3)
        r = o.step_from_dets(f, pre['depth'], g['det_boxes'][f, :n], g['det_scores'][f, :n],
                             np.zeros(n, np.int64), warp=warp)
        m = int(g['out_count'][f])
        np.testing.assert_array_equal(r['instances_id'], g['out_ids'][f, :m], err_msg=f'frame {f}')
        np.testing.assert_array_equal(r['bboxes'], g['out_bboxes'][f, :m])
        np.testing.assert_array_equal(r['scores'], g['out_scores'][f, :m])
        np.testing.assert_array_equal(r['scales'], g['out_scales'][f, :m])
        np.testing.assert_array_equal(r['depth'], g['out_depth'][f, :m])
        st = o.tracker.export_state()
        k = int(g['st_count'][f])
        np.testing.assert_array_equal(st['ids'], g['out_st_ids'][f, :k])
        np.testing.assert_array_equal(st['tentative'], g['out_st_tent'][f, :k])
        np.testing.assert_array_equal(st['tracked'], g['out_st_tracked'][f, :k])
        np.testing.assert_allclose(st['mean'], g['out_st_mean'][f, :k], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(st['cov'], g['out_st_cov'][f, :k], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(st['velocity'], g['out_st_vel'][f, :k], rtol=1e-6, atol=1e-7)
        assert st['num_tracks'] == g['num_tracks'][f]
        lost_recovered += int((~g['out_st_tracked'][f, :k]).sum())
    if name != 'basic':
        assert lost_recovered > 0      # the scenario exercised lost tracks


def test_writers_match_the_reference_decorator_and_metric_lines(tmp_path):
    """shard.write_results_csv / write_mot_txt on the rows of ref_track_basic.npz against the text
    the reference's own save_prediction_results decorator (collect_results.py:1-41) and the
    MOTDroneMetrics statements (mot_drone_metrics.py:195-207,:231-236) produced for them."""
    from stereotracking_b200 import shard
    w = np.load(os.path.join(GOLDEN, 'ref_writers.npz'))
    g = np.load(os.path.join(GOLDEN, 'ref_track_basic.npz'))
    rows = []
    for f in w['frames']:
        m = int(g['out_count'][f])
        res = [dict(instances_id=g['out_ids'][f, :m], bboxes=g['out_bboxes'][f, :m], scores=g['out_scores'][f, :m],
                    depth=g['out_depth'][f, :m], labels=np.zeros(m, np.int64))]
        rows.append(shard.tracks_to_rows(res, int(f), first_sequence=3))
    rows = np.concatenate(rows)
    shard.write_results_csv(rows, 3, str(tmp_path / 'r.csv'))
    assert open(tmp_path / 'r.csv', 'rb').read() == bytes(w['csv_text'])
    shard.write_mot_txt(rows, 3, str(tmp_path / 'r.txt'), depth_thr=int(w['depth_thr']))
    assert open(tmp_path / 'r.txt', 'rb').read() == bytes(w['mot_text'])


def test_kalman_nsa_matches_reference():
    """KalmanFilter(use_nsa=True).project / update with bbox scores (kalman_filter.py:144-145)."""
    from oracle import kalman
    g = np.load(os.path.join(GOLDEN, 'ref_kf_nsa.npz'))
    k = np.load(os.path.join(GOLDEN, 'ref_kf.npz'))
    for n, i in enumerate(g['idx']):
        sc = float(g['score'][n])
        pm, pc = kalman.project(k['mean'][i], k['cov'][i], sc, use_nsa=True)
        np.testing.assert_array_equal(pm, g['pmean'][n])
        np.testing.assert_allclose(pc, g['pcov'][n], rtol=1e-15, atol=0)
        m, c = kalman.update(k['mean'][i], k['cov'][i], g['z'][n], sc, use_nsa=True)
        np.testing.assert_allclose(m, g['umean'][n], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(c, g['ucov'][n], rtol=1e-10, atol=1e-13)
