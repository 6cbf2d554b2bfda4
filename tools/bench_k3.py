"""K3 micro-benchmark at the bench workload (C2, 64 sequences): the per-box depth statistic from
the float32 depth map (st_box_depth) and from the raw u16 frame (st_box_depth_raw) on the
detections K2 produces, alone on one stream (CUDA events, launches back to back).

    python tools/bench_k3.py [frames=6] [reps=20]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from stereotracking_b200.engine import EngineConfig, StereoTrackEngine  # noqa: E402
from stereotracking_b200.synthetic import CONFIGS, SyntheticStereo  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 6
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg, S = CONFIGS['C2'], 64
gen = SyntheticStereo(cfg, S, F, seed=20240604 + 1000, device='cuda')
parts = [gen.frames(f, min(f + 2, F)) for f in range(0, F, 2)]
raw = torch.cat([p['raw'] for p in parts])
heads = [[torch.cat([p[k][l] for p in parts]) for l in range(3)] for k in ('cls', 'bbox', 'obj')]
eng = StereoTrackEngine(EngineConfig(n_seq=S, max_candidates=2048))
Hp, Wp = cfg.padded_hw
boxes, counts, depth = [], [], []
for f in range(F):
    eng.decode_nms([c[f] for c in heads[0]], [c[f] for c in heads[1]], [c[f] for c in heads[2]], (cfg.height, cfg.width))
    boxes.append(eng.det_boxes.clone()); counts.append(eng.det_counts.clone())
    depth.append(eng.disp2depth(raw[f])[0])
torch.cuda.synchronize()
b = torch.cat([boxes[f][s, :int(counts[f][s])] for f in range(F) for s in range(S)]).cpu().numpy()
area = np.clip(np.trunc(b[:, 2]) - np.trunc(b[:, 0]), 0, None) * np.clip(np.trunc(b[:, 3]) - np.trunc(b[:, 1]), 0, None)
out = dict(boxes_per_step=len(b) / F, small=float((area <= 512).mean()), medium=float(((area > 512) & (area <= 8192)).mean()),
           huge=float((area > 8192).mean()), max_pixels=float(area.max()), pixels_per_step=float(area.sum() / F))


def timed(fn):
    for f in range(F):
        fn(f)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        fn(r % F)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


d1 = torch.empty(S, 256, device='cuda'); s1 = torch.empty_like(d1); i1 = torch.empty(S, 256, 4, device='cuda')
out['map_us'] = timed(lambda f: eng.box_depth(depth[f], boxes[f], counts[f], d1, s1, i1))
out['raw_us'] = timed(lambda f: eng.box_depth_raw(raw[f], (Hp, Wp), boxes[f], counts[f], d1, s1, i1))
ws = eng.k3_workspace[:8].cpu().tolist()
out['raw_lists_last_call'] = dict(medium=ws[0], overflow=ws[1], small=ws[3], huge=ws[4])
print(json.dumps(out))
