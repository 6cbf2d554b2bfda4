"""Diagnostics: run the C4 micro workload with diagnostics bit 9 set and dump the cost matrices of the stages the
certified assignment paths refused (gpurun_out/tie_dump.npz; analyse with tests/design/sparse_assign_model.py).

    python tools/k4_tie_dump.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from stereotracking_b200.engine import EngineConfig, StereoTrackEngine
from stereotracking_b200.synthetic import CONFIGS, SyntheticStereo
F, WL, S = 32, 'C4', 64
cfg = CONFIGS[WL]
gen = SyntheticStereo(cfg, S, F, seed=20240604 + 1000, device='cuda')
parts = [gen.frames(f, min(f + 2, F)) for f in range(0, F, 2)]
raw = torch.cat([p['raw'] for p in parts])
heads = [[torch.cat([p[k][l] for p in parts]) for l in range(3)] for k in ('cls', 'bbox', 'obj')]
eng = StereoTrackEngine(EngineConfig(n_seq=S, max_dets=1024, max_tracks=512, max_candidates=4096))
Hp, Wp = cfg.padded_hw
eng.lib.st_debug_set_skip(512)
last = np.zeros(S, np.int64)
out = {}
for f in range(F):
    eng.decode_nms([c[f] for c in heads[0]], [c[f] for c in heads[1]], [c[f] for c in heads[2]], (cfg.height, cfg.width))
    eng.box_depth_raw(raw[f], (Hp, Wp), eng.det_boxes, eng.det_counts, eng.det_depth, eng.det_scales, eng.det_inflated)
    fid = torch.full((S,), f, dtype=torch.int32, device='cuda')
    eng.track_step(fid, eng.det_inflated, eng.det_scores, eng.det_labels, eng.det_scales, eng.det_depth, eng.det_counts)
    torch.cuda.synchronize()
    d = eng.debug_words().astype(np.int64)
    for q in np.nonzero(d[:, 6] > last)[0]:
        R, C = int(d[q, 26]), int(d[q, 27])
        buf = np.zeros(R * C, np.float32)
        eng.lib.st_debug_read_cost(eng._h, int(q), buf.ctypes.data, R * C)
        out[f'f{f}_s{q}_stage{int(d[q, 23])}'] = buf.reshape(R, C)
        print('refused: frame', f, 'seq', q, 'stage', int(d[q, 23]), 'R', R, 'C', C, 'why', d[q, 24:32].tolist(), flush=True)
    last = d[:, 6].copy()
np.savez_compressed('gpurun_out/tie_dump.npz', **out)
print('dumped', len(out))
