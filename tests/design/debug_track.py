"""Drive st_track_step alone with synthetic detections and compare with the oracle tracker."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, ctypes as C
from stereotracking_b200.engine import EngineConfig, StereoTrackEngine
from oracle.tracker import OCSortDisparityOracle
from oracle.model import StereoTrackOracle
n_obj, n_clut, F, D = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
rng = np.random.default_rng(n_obj * 1000 + n_clut)
eng = StereoTrackEngine(EngineConfig(n_seq=1, max_dets=D))
ora = OCSortDisparityOracle(**StereoTrackOracle.DEFAULT_TRACKER)
pos = rng.uniform(50, 1200, (n_obj, 2)); vel = rng.normal(0, 4, (n_obj, 2)); size = rng.uniform(15, 60, (n_obj, 2))
for f in range(F):
    b, s = [], []
    for k in range(n_obj):
        if rng.random() < 0.1: continue
        c = pos[k] + vel[k] * f + rng.normal(0, 0.5, 2)
        b.append([c[0] - size[k, 0] / 2, c[1] - size[k, 1] / 2, c[0] + size[k, 0] / 2, c[1] + size[k, 1] / 2]); s.append(rng.uniform(0.35, 0.95))
    for _ in range(n_clut):
        c = rng.uniform(0, 1250, 2); w = rng.uniform(12, 50, 2)
        b.append([c[0], c[1], c[0] + w[0], c[1] + w[1]]); s.append(rng.uniform(0.05, 0.95))
    b = np.asarray(b, np.float32).reshape(-1, 4); s = np.asarray(s, np.float32)
    o = np.argsort(-s, kind='stable'); b, s = b[o], s[o]
    n = len(b)
    sc = np.full(n, 3.0, np.float32); dp = rng.uniform(2, 80, n).astype(np.float32)
    eng.det_inflated.zero_(); eng.det_scores.zero_(); eng.det_labels.zero_()
    eng.det_inflated[0, :n] = torch.as_tensor(b).cuda(); eng.det_scores[0, :n] = torch.as_tensor(s).cuda()
    eng.det_scales[0, :n] = torch.as_tensor(sc).cuda(); eng.det_depth[0, :n] = torch.as_tensor(dp).cuda()
    eng.det_counts[0] = n; eng.frame_ids[0] = f
    t = time.time()
    eng.track_step(eng.frame_ids, eng.det_inflated, eng.det_scores, eng.det_labels, eng.det_scales, eng.det_depth, eng.det_counts)
    ev = torch.cuda.Event(); ev.record()
    while not ev.query():
        if time.time() - t > 8:
            a = np.zeros((1, 4096), np.int32); eng.lib.st_read_debug(eng._h, a.ctypes.data_as(C.c_void_p), 1)
            print(f'frame {f}: STUCK dets {n} phase {a[0,4090]} trace {a[0,1:1+a[0,0]].tolist()[-60:]} n={a[0,0]}', flush=True); os._exit(3)
        time.sleep(0.002)
    ms = (time.time() - t) * 1e3
    m = int(eng.out_counts[0]); ids = eng.out_ids[0, :m].cpu().numpy()
    ref = ora.track(f, b, s, np.zeros(n, np.int64), sc, dp)
    same = len(ids) == len(ref['instances_id']) and (ids == ref['instances_id']).all()
    print(f'frame {f}: dets {n} out {m} live {len(ora.tracks)} {ms:.2f} ms status {eng.status().tolist()} ids_equal {same}', flush=True)
    if not same: os._exit(4)
print('done', flush=True)
