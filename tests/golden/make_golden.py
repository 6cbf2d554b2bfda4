"""Generate the committed golden fixtures by running the UNMODIFIED reference
modules (loaded from /root/reference through ``ref_shims.py``) on seeded
synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (small .npz files next to this script):
  ref_kf.npz          KalmanFilter.initiate/predict/update chains
  ref_kf_gating.npz   KalmanFilter.project / gating_distance / track on states of ref_kf.npz
  ref_boxdepth.npz    OCSORT_Disparity.extract_depth / bbox_postp_depth /
                      scale_bbox on random and edge-case boxes
  ref_boxdepth_ablation.npz  the four ablation reducers of depth_extraction_comparison.py
  ref_frontend.npz    pad + HWC->CHW + TrackDataPreprocessor_Disparity_V1 lines + stack_batch on u8 images
  ref_memo_occlusion.npz  BaseTracker.memo / get while replaying ref_track_occlusion.npz
  ref_kf_nsa.npz      KalmanFilter(use_nsa=True).project / update with bbox scores
  ref_writers.npz     results.csv text written by the reference's own save_prediction_results
                      decorator, and the MOT txt lines of MOTDroneMetrics, on rows of ref_track_basic.npz
  ref_track_<name>.npz  OCSORT_Disparity.predict() over whole sequences:
                      inputs (raw disparity, detections) and the reference's
                      pred_track_instances + tracker state per frame
The oracle (oracle/) is then checked against these files by
tests/test_oracle_golden.py, and the CUDA path by tests/test_gpu_*.py.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.environ.get('GOLDEN_OUT', HERE)       # where the fixtures are written (and re-read by later steps)
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shims  # noqa: E402
from oracle import disparity as odisp  # noqa: E402  (A1 preprocessing of inputs only)
from stereotracking_b200.synthetic import SceneConfig, SyntheticStereo  # noqa: E402

TRACKER_CFG = dict(obj_score_thr=0.3, init_track_thr=0.7,
                   weight_iou_with_det_scores=False, match_iou_thr=0.1,
                   num_tentatives=3, vel_consist_weight=0.2, vel_delta_t=3,
                   num_frames_retain=30)


def ragged(list_of_arrays, width, dtype):
    n = max([len(a) for a in list_of_arrays] + [1])
    shape = (len(list_of_arrays), n) + ((width,) if width else ())
    out = np.zeros(shape, dtype)
    cnt = np.zeros(len(list_of_arrays), np.int64)
    for i, a in enumerate(list_of_arrays):
        a = np.asarray(a, dtype).reshape((-1,) + ((width,) if width else ()))
        out[i, :len(a)] = a
        cnt[i] = len(a)
    return out, cnt


def make_kf(ns):
    rng = np.random.default_rng(11)
    kf = ns.KalmanFilter()
    z0 = np.asarray([[300.5, 200.25, 0.8, 40.0], [15.0, 700.0, 1.3, 11.5]], np.float32)
    means, covs, zs, ops = [], [], [], []
    for z in z0:
        m, c = kf.initiate(z)
        means.append(m.copy()); covs.append(c.copy()); zs.append(z.astype(np.float64)); ops.append(0)
        for step in range(12):
            if rng.random() < 0.7:
                m, c = kf.predict(m, c)
                means.append(m.copy()); covs.append(c.copy()); zs.append(np.zeros(4)); ops.append(1)
            zz = (m[:4] + rng.normal(0, [2, 2, 0.01, 1])).astype(np.float32)
            m, c = kf.update(m, c, zz)
            means.append(m.copy()); covs.append(c.copy()); zs.append(zz.astype(np.float64)); ops.append(2)
    np.savez_compressed(os.path.join(OUT, 'ref_kf.npz'), mean=np.stack(means),
                        cov=np.stack(covs), z=np.stack(zs), op=np.asarray(ops),
                        chain_start=np.asarray([0, ops.index(0, 1)]))


def make_kf_gating(ns):
    """KalmanFilter.project / gating_distance / track of the unmodified reference on the states
    of ref_kf.npz."""
    import types
    g = np.load(os.path.join(OUT, 'ref_kf.npz'))
    rng = np.random.default_rng(13)
    kf = ns.KalmanFilterRaw()
    idx = np.arange(0, len(g['op']), 3)
    pm, pc, meas, d4, d2 = [], [], [], [], []
    for i in idx:
        m, c = g['mean'][i], g['cov'][i]
        a, b = kf.project(m, c)
        pm.append(a); pc.append(b)
        z = (m[:4] + rng.normal(0, [6, 6, 0.05, 4], (5, 4))).astype(np.float32)
        meas.append(z)
        d4.append(kf.gating_distance(m, c, z))
        d2.append(kf.gating_distance(m, c, z, only_position=True))
    tracks = {7: types.SimpleNamespace(mean=g['mean'][5].copy(), covariance=g['cov'][5].copy()),
              9: types.SimpleNamespace(mean=g['mean'][30].copy(), covariance=g['cov'][30].copy())}
    boxes = torch.from_numpy(np.concatenate([meas[1][:3], meas[10][:3]]))
    _, costs = kf.track(tracks, boxes)
    np.savez_compressed(os.path.join(OUT, 'ref_kf_gating.npz'), idx=idx, pmean=np.stack(pm), pcov=np.stack(pc),
                        meas=np.stack(meas), dist4=np.stack(d4), dist2=np.stack(d2),
                        track_in=np.asarray([5, 30]), track_boxes=boxes.numpy(), track_costs=costs,
                        track_mean=np.stack([tracks[7].mean, tracks[9].mean]),
                        track_cov=np.stack([tracks[7].covariance, tracks[9].covariance]))
    print('kf gating:', len(idx), 'states; track costs NaN', int(np.isnan(costs).sum()), 'of', costs.size)


def edge_depth_map(rng, H=96, W=128):
    raw = rng.integers(10, 900, (H, W)).astype(np.uint16)
    raw[rng.random((H, W)) < 0.2] = 65535
    raw[40:60, 50:80] = 320            # a constant-depth object
    raw[0:8, 0:8] = 65535              # an all-invalid corner
    raw[70:72, 90:92] = 400            # tiny valid island inside invalid area
    raw[64:80, 84:100][raw[64:80, 84:100] != 400] = 65535
    return raw


def make_boxdepth(ns):
    rng = np.random.default_rng(5)
    model = ref_shims.build_model(ns, TRACKER_CFG)
    raw = edge_depth_map(rng)
    pre = odisp.preprocess(raw)
    disp = torch.from_numpy(pre['disp'])[None, None].repeat(1, 3, 1, 1)
    H, W = pre['depth'].shape
    boxes = []
    for _ in range(60):
        x1, y1 = rng.uniform(-5, W - 4), rng.uniform(-5, H - 4)
        boxes.append([x1, y1, x1 + rng.uniform(1, 60), y1 + rng.uniform(1, 60)])
    boxes += [[50, 40, 80, 60], [52.3, 41.7, 77.9, 58.2], [0, 0, 8, 8], [0, 0, 3, 3],
              [70, 70, 72, 72], [70.2, 70.9, 71.9, 71.1], [84, 64, 100, 80], [90, 70, 91, 71],
              [-3.5, -2.5, 6, 7], [W - 3, H - 3, W + 5, H + 5], [W, H, W + 4, H + 4],
              [10, 10, 10, 30], [10, 10, 30, 10], [1, 1, 2.5, 2.5], [0.5, 0.5, 1.9, 1.9],
              [20, 20, 19, 40], [5, 90, 120, 96], [0, 0, W, H], [126.5, 3, 128, 9]]
    boxes = torch.tensor(boxes, dtype=torch.float32)
    depth_t = model.disp2depth(disp[:, 0:1])
    d, s = model.extract_depth(depth_t, boxes)
    inst = ns.InstanceData(bboxes=boxes.clone())
    inst2, dv = model.bbox_postp_depth(inst, disp, None)
    np.savez_compressed(
        os.path.join(OUT, 'ref_boxdepth.npz'), raw=raw, boxes=boxes.numpy(),
        depth_map=depth_t.numpy().squeeze(), d=np.asarray(d, np.float32),
        scales=np.asarray(s, np.float32), inflated=inst2.bboxes.numpy(),
        deflated=ns.scale_bbox(inst2.bboxes, 1 / inst2.scales).numpy(),
        cxcyah=ns.bbox_xyxy_to_cxcyah(boxes).numpy(),
        xyxy_back=ns.bbox_cxcyah_to_xyxy(ns.bbox_xyxy_to_cxcyah(boxes)).numpy())


def make_boxdepth_ablation(ns):
    """The four ablation reducers of mmtrack/models/mot/depth_extraction_comparison.py on the
    inputs of ref_boxdepth.npz plus a larger map with bigger boxes.  The decorators replace
    ``extract_depth`` (they ignore the wrapped function) and are called box by box so that an
    IndexError of ``center`` (centre pixel outside the map) can be recorded per box."""
    import sys as _sys
    mod = _sys.modules['mmtrack.models.mot.depth_extraction_comparison']
    base = np.load(os.path.join(OUT, 'ref_boxdepth.npz'))
    rng = np.random.default_rng(17)
    raw2 = rng.integers(8, 700, (200, 320)).astype(np.uint16)
    raw2 //= 4                                   # many duplicates, few distinct values
    raw2[raw2 < 2] = 2
    raw2[rng.random(raw2.shape) < 0.25] = 65535
    raw2[60:140, 100:220] = 96
    boxes2 = []
    for _ in range(40):
        x1, y1 = rng.uniform(-20, 300), rng.uniform(-20, 190)
        boxes2.append([x1, y1, x1 + rng.uniform(2, 200), y1 + rng.uniform(2, 150)])
    boxes2 += [[0, 0, 320, 200], [100, 60, 220, 140], [310, 190, 400, 260], [-50, -40, 30, 20],
               [150, 100, 151, 101], [150, 100, 152, 101], [150, 100, 152, 102], [10, 10, 13, 11]]
    out = {}
    for tag, raw, boxes in (('a', base['raw'], base['boxes']), ('b', raw2, np.asarray(boxes2, np.float32))):
        depth = odisp.preprocess(raw)['depth']
        depth_t = torch.from_numpy(depth)[None, None]
        out[f'raw_{tag}'] = raw
        out[f'boxes_{tag}'] = np.asarray(boxes, np.float32)
        for mode in ('truncated_mean', 'mean', 'median', 'center'):
            fn = getattr(mod, mode + '_decorator')(lambda *a, **k: None)
            d = np.full(len(boxes), np.nan, np.float32)
            sc = np.ones(len(boxes), np.float32)
            err = np.zeros(len(boxes), np.int32)
            for i, b in enumerate(boxes):
                try:
                    with np.errstate(all='ignore'):
                        dv, sv = fn(None, depth_t, torch.tensor(np.asarray(b, np.float32)[None]))
                    d[i], sc[i] = dv[0], sv[0]
                except IndexError:
                    err[i] = 1
            out[f'd_{mode}_{tag}'], out[f'scale_{mode}_{tag}'], out[f'oob_{mode}_{tag}'] = d, sc, err
    np.savez_compressed(os.path.join(OUT, 'ref_boxdepth_ablation.npz'), **out)
    print('ablation:', {k: (int(np.isnan(v).sum()) if v.dtype.kind == 'f' else int(v.sum()))
                        for k, v in out.items() if k.startswith(('oob_',))})


def make_frontend(ns):
    """RGB half of the detector input: the reference's own lines of
    TrackDataPreprocessor_Disparity_V1.forward (data_preprocessor_disparity_v1.py:38-52, executed
    here statement by statement on torch tensors - the class itself needs mmengine) and the
    reference's ``stack_batch`` (mmtrack/utils/misc.py:13-64, loaded from /root/reference), after
    the u8 pad of mmdet's Pad (third-party, restated as a NumPy bottom/right pad with 114) and
    PackTrackInputs_Disparity's HWC -> CHW transpose (formatting_disparity.py:166-176)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('mmtrack.utils.misc', os.path.join(ref_shims.REF, 'mmtrack', 'utils', 'misc.py'))
    misc = importlib.util.module_from_spec(spec)
    misc.__package__ = 'mmtrack.utils'
    spec.loader.exec_module(misc)
    rng = np.random.default_rng(23)
    out = {}
    for tag, (H, W) in (('a', (40, 52)), ('b', (30, 45))):
        img = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        Hp, Wp = -(-H // 32) * 32, -(-W // 32) * 32
        padded = np.full((Hp, Wp, 3), 114, np.uint8)           # mmcv.impad(img, shape=..., pad_val=(114.0,)*3)
        padded[:H, :W] = img
        packed = torch.from_numpy(np.ascontiguousarray(padded[None].transpose(0, 3, 1, 2)))   # (T=1,C,H,W)
        out[f'img_{tag}'] = img
        for name, (mean, std, swap) in (('plain', (None, None, False)),
                                       ('norm', ((123.675, 116.28, 103.53), (58.395, 57.12, 57.375), True))):
            imgs = [packed]
            if swap:
                imgs = [_img[:, [2, 1, 0], ...] for _img in imgs]                              # :44-45
            imgs = [_img.float() for _img in imgs]                                             # :47
            if mean is not None:
                m = torch.tensor(mean).view(1, -1, 1, 1)                                       # data_preprocessor.py:77-80
                sd = torch.tensor(std).view(1, -1, 1, 1)
                imgs = [(_img - m) / sd for _img in imgs]                                      # :49
            res = misc.stack_batch(imgs, 32, 0)                                                # :51-52
            assert res.shape == (1, 1, 3, Hp, Wp)
            out[f'out_{name}_{tag}'] = res[0, 0].numpy()
    np.savez_compressed(os.path.join(OUT, 'ref_frontend.npz'), **out)
    print('frontend:', {k: v.shape for k, v in out.items()})


def synth_dets(gen, F, rng, jitter=0.6, miss=0.05, clutter=3, blackout=()):
    """Noisy-GT detections per frame (what the detector would return)."""
    boxes, vis, _ = gen.gt_boxes(0, F)
    H, W = gen.cfg.height, gen.cfg.width
    out = []
    for f in range(F):
        b, s = [], []
        if f not in blackout:
            for k in range(boxes.shape[2]):
                if not vis[f, 0, k] or rng.random() < miss:
                    continue
                bb = boxes[f, 0, k].numpy() + rng.normal(0, jitter, 4)
                bb[0::2] = np.clip(bb[0::2], 0, W); bb[1::2] = np.clip(bb[1::2], 0, H)
                b.append(bb); s.append(rng.uniform(0.25, 0.98))
            for _ in range(rng.integers(0, clutter + 1)):
                x, y = rng.uniform(0, W - 30), rng.uniform(0, H - 30)
                b.append([x, y, x + rng.uniform(4, 30), y + rng.uniform(4, 30)])
                s.append(rng.uniform(0.05, 0.9))
        b = np.asarray(b, np.float32).reshape(-1, 4)
        s = np.asarray(s, np.float32)
        o = np.argsort(-s, kind='stable')
        out.append((b[o], s[o]))
    return out


def ego_warps(gen, F):
    """Frame-to-frame background affine of a heavy_ego scene: the generator maps a
    static scene point u0 to u_t = s_t R_t (u0 - c) + c + T_t, so u_t = A u_{t-1} + b with
    A = s_t R_t (s_{t-1} R_{t-1})^-1 and b = c + T_t - A (c + T_{t-1}).  Every fifth
    frame has no warp (the reference's estimator returns None when RANSAC fails)."""
    import math
    cfg = gen.cfg
    c = np.array([cfg.width / 2.0, cfg.height / 2.0])
    fr = np.arange(F, dtype=np.float64)
    sc = 1 + 0.02 * np.sin(fr / 7)
    ang = math.radians(0.5) * np.sin(fr / 11)
    M = [sc[t] * np.array([[math.cos(ang[t]), -math.sin(ang[t])], [math.sin(ang[t]), math.cos(ang[t])]]) for t in range(F)]
    T = gen.ego_t[0].numpy()
    warps = [None]
    for t in range(1, F):
        if t % 5 == 4:
            warps.append(None)
            continue
        A = M[t] @ np.linalg.inv(M[t - 1])
        b = c + T[t] - A @ (c + T[t - 1])
        warps.append(np.concatenate([A, b[:, None]], axis=1).astype(np.float32))
    return warps


def run_reference_sequence(ns, raws, dets, warps=None, tracker_cfg=None, reducer=None, probe=None):
    cfg = dict(TRACKER_CFG if tracker_cfg is None else tracker_cfg)
    if warps is not None:
        cfg['cmc'] = dict(method='glme_affine')
    model = ref_shims.build_model(ns, cfg)
    if reducer is not None:
        # an ablation run: extract_depth decorated with depth_extraction_comparison.py's <reducer>_decorator
        import types
        dec = getattr(sys.modules['mmtrack.models.mot.depth_extraction_comparison'], reducer + '_decorator')
        model.extract_depth = types.MethodType(dec(type(model).extract_depth), model)
    if warps is not None:
        # the estimator (Farneback flow + RANSAC on the images, trackers/utils.py:6-55) is
        # host-side OpenCV and out of scope; the track() code path and
        # apply_gmc_to_tracks_cxcyah (gmc.py:20-45) run unmodified on scripted warps
        model.tracker.estimate_camera_motion = lambda img, metainfo: warps[metainfo['frame_id']]
    frames = []
    for f, (raw, (b, s)) in enumerate(zip(raws, dets)):
        pre = odisp.preprocess(raw)
        disp = torch.from_numpy(pre['disp'])[None, None, None].repeat(1, 1, 3, 1, 1)
        mask = torch.from_numpy(pre['mask'])[None, None, None]
        img = torch.zeros_like(disp)
        model.detector.next = ns.InstanceData(
            bboxes=torch.from_numpy(b.copy()), scores=torch.from_numpy(s.copy()),
            labels=torch.zeros(len(b), dtype=torch.int64))
        sample = ns.TrackDataSample(frame_id=f)
        res = model.predict(dict(img=img, disp_postp=disp, disp_mask=mask), [sample])[0]
        ti = res.pred_track_instances
        trk = model.tracker
        if probe is not None:
            probe(f, trk)
        ids = list(trk.tracks.keys())
        frames.append(dict(
            bboxes=ti.bboxes.numpy(), ids=ti.instances_id.numpy(), scores=ti.scores.numpy(),
            labels=ti.labels.numpy(), scales=ti.scales.numpy(),
            depth=np.asarray(ti.depth, np.float32),
            st_ids=np.asarray(ids, np.int64),
            st_mean=np.stack([np.asarray(trk.tracks[i].mean, np.float64) for i in ids]) if ids else np.zeros((0, 8)),
            st_cov=np.stack([trk.tracks[i].covariance for i in ids]) if ids else np.zeros((0, 8, 8)),
            st_tent=np.asarray([bool(trk.tracks[i].tentative) for i in ids]),
            st_tracked=np.asarray([bool(trk.tracks[i].tracked) for i in ids]),
            st_vel=np.stack([trk.tracks[i].velocity.numpy().astype(np.float32) for i in ids]) if ids else np.zeros((0, 2), np.float32),
            num_tracks=int(trk.num_tracks)))
    return frames


def save_sequence(name, raws, dets, frames, warps=None):
    pack = {}
    if warps is not None:
        w = np.full((len(warps), 6), np.nan, np.float32)
        for i, m in enumerate(warps):
            if m is not None:
                w[i] = np.asarray(m, np.float32).reshape(6)
        pack['warp'] = w
    pack['raw'] = np.stack(raws)
    pack['det_boxes'], pack['det_count'] = ragged([d[0] for d in dets], 4, np.float32)
    pack['det_scores'], _ = ragged([d[1] for d in dets], 0, np.float32)
    for key, width, dt in [('bboxes', 4, np.float32), ('ids', 0, np.int64), ('scores', 0, np.float32),
                           ('labels', 0, np.int64), ('scales', 0, np.float32), ('depth', 0, np.float32),
                           ('st_ids', 0, np.int64), ('st_mean', 8, np.float64), ('st_tent', 0, np.bool_),
                           ('st_tracked', 0, np.bool_), ('st_vel', 2, np.float32)]:
        pack['out_' + key], cnt = ragged([fr[key] for fr in frames], width, dt)
        if key == 'ids':
            pack['out_count'] = cnt
        if key == 'st_ids':
            pack['st_count'] = cnt
    n = pack['out_st_ids'].shape[1]
    cov = np.zeros((len(frames), n, 8, 8))
    for i, fr in enumerate(frames):
        cov[i, :len(fr['st_cov'])] = fr['st_cov']
    pack['out_st_cov'] = cov
    pack['num_tracks'] = np.asarray([fr['num_tracks'] for fr in frames], np.int64)
    np.savez_compressed(os.path.join(OUT, f'ref_track_{name}.npz'), **pack)
    total = sum(len(fr['ids']) for fr in frames)
    print(f'{name}: {len(frames)} frames, {total} track rows, max id {max(fr["num_tracks"] for fr in frames)}')


def make_memo(ns):
    """BaseTracker.memo / get (base_tracker.py:120-177) of the unmodified reference tracker while
    it replays the inputs of ref_track_occlusion.npz."""
    g = np.load(os.path.join(OUT, 'ref_track_occlusion.npz'))
    F = len(g['raw'])
    raws = [g['raw'][f] for f in range(F)]
    dets = [(g['det_boxes'][f, :g['det_count'][f]], g['det_scores'][f, :g['det_count'][f]]) for f in range(F)]
    pack = {}
    probes = (6, 25, F - 1)

    def probe(f, trk):
        if f not in probes:
            return
        memo = trk.memo
        for k in ('ids', 'bboxes', 'scores', 'labels', 'scales', 'depth', 'frame_ids'):
            pack[f'memo_{k}_{f}'] = memo[k].numpy()
        pack[f'get_bboxes_mean3_{f}'] = trk.get('bboxes', num_samples=3, behavior='mean').numpy()
        pack[f'get_scores_last_{f}'] = trk.get('scores').numpy()
        two = trk.ids[:2]
        pack[f'get_two_{f}'] = np.asarray(two, np.int64)
        # tracks with >= 2 entries each, so that the (1, n, 4) pieces can be concatenated
        full = [i for i in trk.ids if len(trk.tracks[i]['bboxes']) >= 2][:2]
        pack[f'get_full_{f}'] = np.asarray(full, np.int64)
        if full:
            pack[f'get_bboxes_none2_{f}'] = trk.get('bboxes', ids=full, num_samples=2).numpy()
    frames = run_reference_sequence(ns, raws, dets, probe=probe)
    for f in range(F):       # the replay is the stored run
        assert np.array_equal(frames[f]['ids'], g['out_ids'][f, :g['out_count'][f]])
    pack['probes'] = np.asarray(probes)
    np.savez_compressed(os.path.join(OUT, 'ref_memo_occlusion.npz'), **pack)
    print('memo:', {k: v.shape for k, v in pack.items() if k.startswith('memo_ids')})


def make_kf_nsa(ns):
    """KalmanFilter(use_nsa=True) of the unmodified reference (kalman_filter.py:38,:144-145):
    project and update with bbox scores on states of ref_kf.npz."""
    g = np.load(os.path.join(OUT, 'ref_kf.npz'))
    rng = np.random.default_rng(17)
    kf = ns.KalmanFilterRaw(use_nsa=True)
    idx = np.arange(1, len(g['op']), 4)
    score = rng.uniform(0.05, 0.99, len(idx))
    score[0] = 0.0
    pm, pc, zs, um, uc = [], [], [], [], []
    for i, sc in zip(idx, score):
        m, c = g['mean'][i], g['cov'][i]
        a, b = kf.project(m, c, float(sc))
        pm.append(a); pc.append(b)
        z = (m[:4] + rng.normal(0, [2, 2, 0.01, 1])).astype(np.float32)
        m2, c2 = kf.update(m, c, z, float(sc))
        zs.append(z); um.append(m2); uc.append(c2)
    np.savez_compressed(os.path.join(OUT, 'ref_kf_nsa.npz'), idx=idx, score=score, pmean=np.stack(pm), pcov=np.stack(pc),
                        z=np.stack(zs), umean=np.stack(um), ucov=np.stack(uc))
    print('kf nsa:', len(idx), 'states')


def make_writers(ns):
    """The two output writers on rows of ref_track_basic.npz.
    results.csv: the reference's own decorator (mmtrack/utils/collect_results.py:1-41, loaded from
    its file; it needs only csv + os) around a predict function that returns the canned frames.
    MOT txt: the statements of MOTDroneMetrics.process / _save_one_video_gts_preds
    (mmtrack/evaluation/metrics/mot_drone_metrics.py:195-207,:231-236; the module itself needs
    mmengine + trackeval, so its lines are executed here on the same torch tensors)."""
    import importlib.util
    import tempfile
    spec = importlib.util.spec_from_file_location('ref_collect_results',
                                                  os.path.join(ref_shims.REF, 'mmtrack/utils/collect_results.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(OUT, 'ref_track_basic.npz'))
    frames = [2, 3, 10, 20, 21]
    canned = {}
    for f in frames:
        m = int(g['out_count'][f])
        depth = [(-1 if v == -1 else np.float32(v)) for v in g['out_depth'][f, :m]]     # extract_depth's list (:148,:168)
        inst = ns.InstanceData(bboxes=torch.from_numpy(g['out_bboxes'][f, :m].copy()),
                               instances_id=torch.from_numpy(g['out_ids'][f, :m].copy()),
                               labels=torch.zeros(m, dtype=torch.int64),
                               scores=torch.from_numpy(g['out_scores'][f, :m].copy()), depth=depth, gt_depth=depth)
        canned[f] = inst
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, 'results.csv')

        class Sample:
            pass

        def predict(f):
            s = Sample()
            s.pred_track_instances, s.frame_id = canned[f], f
            return [s]

        wrapped = mod.save_prediction_results(path)(predict)
        for f in frames:
            wrapped(f)
        csv_text = open(path, 'rb').read()
    # MOT txt, depth_thr = 50 (ignore_depth False): mot_drone_metrics.py:195-207 then :231-236
    depth_thr, pred_tracks = 50, []
    for f in frames:
        pred_instances = canned[f]
        pred_tracks.extend([
            np.array([
                f + 1, pred_instances['instances_id'][i].cpu(),
                pred_instances['bboxes'][i][0].cpu(),
                pred_instances['bboxes'][i][1].cpu(),
                (pred_instances['bboxes'][i][2] -
                 pred_instances['bboxes'][i][0]).cpu(),
                (pred_instances['bboxes'][i][3] -
                 pred_instances['bboxes'][i][1]).cpu(),
                pred_instances['scores'][i].cpu()
            ]) for i in range(len(pred_instances['instances_id'])) if
            (depth_thr >= pred_instances['depth'][i] > 0)
        ])
    pred_tracks = np.array(pred_tracks)
    mot_text = ''.join('%d,%d,%.3f,%.3f,%.3f,%.3f,%.3f,-1,-1,-1\n' % (
        t[0], t[1], t[2], t[3], t[4], t[5], t[6]) for t in pred_tracks).encode()
    np.savez_compressed(os.path.join(OUT, 'ref_writers.npz'), frames=np.asarray(frames),
                        csv_text=np.frombuffer(csv_text, np.uint8), mot_text=np.frombuffer(mot_text, np.uint8),
                        depth_thr=np.asarray(depth_thr))
    print('writers:', len(csv_text), 'bytes of csv,', len(mot_text), 'bytes of MOT txt')


def make_resize(ns):
    """Image branch of Resize_Disparity at a resize factor != 1 (transforms_disparity.py:128 runs mmdet's
    Resize._resize_img: ``mmcv.imrescale(img, scale, interpolation='bilinear', return_scale=True,
    backend='cv2')``).  mmdet / mmcv are third-party and absent; their two steps are executed here as
    they are written there: new size = int(side * factor + 0.5) with factor = min(long / max(h, w),
    short / min(h, w)) (mmcv rescale_size), then ``cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)``
    with the OpenCV of this image (the library the reference itself calls)."""
    import cv2
    rng = np.random.default_rng(31)
    out = {}
    cases = dict(down=((54, 96), (36, 64)), up=((45, 80), (72, 128)), half=((64, 96), (32, 48)),
                 odd=((37, 53), (50, 90)), tall=((80, 30), (64, 64)))
    for tag, ((H, W), scale) in cases.items():
        img = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        if tag == 'odd':                                  # smooth content too, not only noise
            yy, xx = np.mgrid[0:H, 0:W]
            img = np.stack([(xx * 255 // (W - 1)), (yy * 255 // (H - 1)), ((xx + yy) % 256)], -1).astype(np.uint8)
        factor = min(max(scale) / max(H, W), min(scale) / min(H, W))
        nw, nh = int(W * float(factor) + 0.5), int(H * float(factor) + 0.5)
        res = cv2.resize(img, (nw, nh), dst=None, interpolation=cv2.INTER_LINEAR)
        out[f'img_{tag}'] = img
        out[f'scale_{tag}'] = np.asarray(scale, np.int64)
        out[f'out_{tag}'] = res
    np.savez_compressed(os.path.join(OUT, 'ref_resize_bilinear.npz'), **out)
    print('resize:', {k: v.shape for k, v in out.items() if k.startswith('out_')})


def make_sequences(ns, only=None):
    scen = {
        'basic': dict(cfg=SceneConfig(160, 224, 5, 8, n_clutter=0), F=36, seed=3, miss=0.05, clutter=3, blackout=()),
        'occlusion': dict(cfg=SceneConfig(160, 224, 6, 9, n_clutter=0), F=48, seed=4, miss=0.30, clutter=2,
                          blackout=(20, 21, 33)),
        'stress': dict(cfg=SceneConfig(192, 256, 8, 14, invalid_prob=0.30, n_clutter=0, heavy_ego=True),
                       F=40, seed=5, miss=0.15, clutter=8, blackout=(0, 9)),
        # constructor defaults of OCSORTTracker_Disparity (:42-52) instead of the AirDrone config:
        # weight_iou_with_det_scores=True, match_iou_thr=0.3, num_frames_retain=10
        'defaults': dict(cfg=SceneConfig(160, 224, 6, 9, n_clutter=0), F=40, seed=7, miss=0.25, clutter=3,
                         blackout=(12, 13), tracker_cfg={}),
        # camera-motion compensation on (cmc=dict(method='glme_affine')), scripted warps
        'cmc': dict(cfg=SceneConfig(192, 256, 6, 10, n_clutter=0, heavy_ego=True),
                    F=40, seed=6, miss=0.20, clutter=3, blackout=(17,), cmc=True),
        # ablation run: extract_depth replaced by the median reducer, scale = clip(d*d/400, 1, 3)
        'median': dict(cfg=SceneConfig(160, 224, 6, 9, n_clutter=0), F=36, seed=8, miss=0.15, clutter=3,
                       blackout=(14,), reducer='median'),
    }
    for name, sc in scen.items():
        if only is not None and name not in only:
            continue
        gen = SyntheticStereo(sc['cfg'], 1, sc['F'], seed=sc['seed'])
        fr = gen.frames(0, sc['F'])
        raws = [fr['raw'][f, 0].numpy() for f in range(sc['F'])]
        rng = np.random.default_rng(sc['seed'] + 100)
        dets = synth_dets(gen, sc['F'], rng, miss=sc['miss'], clutter=sc['clutter'],
                          blackout=sc['blackout'])
        warps = ego_warps(gen, sc['F']) if sc.get('cmc') else None
        frames = run_reference_sequence(ns, raws, dets, warps, sc.get('tracker_cfg'), sc.get('reducer'))
        save_sequence(name, raws, dets, frames, warps)


if __name__ == '__main__':
    ns = ref_shims.load()
    if sys.argv[1:] == ['ablation']:             # only the ablation fixtures (reads ref_boxdepth.npz)
        make_boxdepth_ablation(ns)
        make_sequences(ns, only=('median',))
        sys.exit(0)
    if sys.argv[1:] == ['frontend']:
        make_frontend(ns)
        sys.exit(0)
    if sys.argv[1:] == ['memo']:
        make_memo(ns)
        sys.exit(0)
    if sys.argv[1:] == ['kf_gating']:
        make_kf_gating(ns)
        sys.exit(0)
    if sys.argv[1:] == ['round2']:               # fixtures added in round 2 (read ref_kf.npz / ref_track_basic.npz)
        make_kf_nsa(ns)
        make_writers(ns)
        make_resize(ns)
        sys.exit(0)
    make_kf(ns)
    make_kf_gating(ns)
    make_boxdepth(ns)
    make_boxdepth_ablation(ns)
    make_frontend(ns)
    make_sequences(ns)
    make_memo(ns)
    make_kf_nsa(ns)
    make_writers(ns)
    make_resize(ns)
    print('golden fixtures written to', OUT)
