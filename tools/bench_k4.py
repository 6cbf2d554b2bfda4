"""K4 micro-benchmark at the bench workload (C2, 64 sequences): the tracker kernel alone, launches
back to back on one stream, on the detections K2 + K3 produce for consecutive frames.

    python tools/bench_k4.py [frames=40] [workload=C2] [sequences=64]
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

F = int(sys.argv[1]) if len(sys.argv) > 1 else 40
WL = sys.argv[2] if len(sys.argv) > 2 else 'C2'
cfg, S = CONFIGS[WL], int(sys.argv[3]) if len(sys.argv) > 3 else 64
CAPS = dict(C3=dict(max_dets=512, max_tracks=512, max_candidates=4096),
            C4=dict(max_dets=1024, max_tracks=512, max_candidates=4096)).get(WL, dict(max_candidates=2048))
gen = SyntheticStereo(cfg, S, F, seed=20240604 + 1000, device='cuda')
parts = [gen.frames(f, min(f + 2, F)) for f in range(0, F, 2)]
raw = torch.cat([p['raw'] for p in parts])
heads = [[torch.cat([p[k][l] for p in parts]) for l in range(3)] for k in ('cls', 'bbox', 'obj')]
eng = StereoTrackEngine(EngineConfig(n_seq=S, **CAPS))
Hp, Wp = cfg.padded_hw
dets = []
for f in range(F):
    eng.decode_nms([c[f] for c in heads[0]], [c[f] for c in heads[1]], [c[f] for c in heads[2]], (cfg.height, cfg.width))
    eng.box_depth_raw(raw[f], (Hp, Wp), eng.det_boxes, eng.det_counts, eng.det_depth, eng.det_scales, eng.det_inflated)
    dets.append([t.clone() for t in (eng.det_inflated, eng.det_scores, eng.det_labels, eng.det_scales, eng.det_depth,
                                     eng.det_counts)])
fids = [torch.full((S,), f, dtype=torch.int32, device='cuda') for f in range(F)]
torch.cuda.synchronize()
out = dict(workload=WL, sequences=S)
for rep in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(F + 1)]
    ev[0].record()
    for f in range(F):
        eng.track_step(fids[f], *dets[f])
        ev[f + 1].record()
    torch.cuda.synchronize()
    us = np.array([ev[f].elapsed_time(ev[f + 1]) * 1e3 for f in range(F)])
    out[f'rep{rep}'] = dict(total_us_per_frame=float(ev[0].elapsed_time(ev[F]) * 1e3 / F), median_us=float(np.median(us[8:])),
                            max_us=float(us[8:].max()))
# one more pass, synchronised per frame: the CTAs' own start / end stamps against the event time of the launch
span, evt, skew, longest = [], [], [], []
eng.lib.st_debug_set_skip(128)          # diagnostics bit 7: global-timer stamps
for f in range(F):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.track_step(fids[f], *dets[f])
    e1.record()
    torch.cuda.synchronize()
    d = eng.debug_words().astype(np.int64)
    st_ns, en_ns = d[:, 21] & 0xffffffff, d[:, 22] & 0xffffffff
    span.append(float((en_ns.max() - st_ns.min()) / 1e3)); evt.append(e0.elapsed_time(e1) * 1e3)
    skew.append(float((st_ns.max() - st_ns.min()) / 1e3)); longest.append(float((en_ns - st_ns).max() / 1e3))
out['per_launch_us_median'] = dict(event=float(np.median(evt[8:])), first_start_to_last_end=float(np.median(span[8:])),
                                  start_skew=float(np.median(skew[8:])), longest_cta=float(np.median(longest[8:])),
                                  total_cycles_max_last=int(d[:, 20].max()))
eng.lib.st_debug_set_skip(int(os.environ.get('ST_K4_SKIP', '0')))
worst_frames = []
refusals, last_cnt = {}, None
for f in range(F):
    eng.track_step(fids[f], *dets[f])
    if True:                                    # the slowest sequence of every frame
        torch.cuda.synchronize()
        d = eng.debug_words().astype(np.int64)
        w = int(np.argmax(d[:, 8:14].sum(1)))
        worst_frames.append(dict(frame=f, cycles=int(d[w, 8:14].sum()), stage1=int(d[w, 9]),
                                 sub=[int(x) for x in d[w, [19, 14, 15, 16, 17, 18]]], dets=int(dets[f][5][w]),
                                 fast_path=dict(zip(('code', 'refusals', 'start', 'search', 'certificate', 'searched', 'active', 'arcs'),
                                                    (int(x) for x in d[w, 24:32])))))
        # refusals of the certified whole-problem path of stage 1 (debug words 24 / 25), by reason code
        if last_cnt is not None:
            for q in np.nonzero(d[:, 25] > last_cnt)[0]:
                refusals[int(d[q, 24])] = refusals.get(int(d[q, 24]), 0) + 1
        last_cnt = d[:, 25].copy()
torch.cuda.synchronize()
if worst_frames:
    worst_frames.sort(key=lambda r: -r['cycles'])
    out['slowest_sequence_per_frame_top5'] = worst_frames[:5]
    out['stage1_fast_path_refusals_by_code'] = refusals
    out['slowest_sequence_cycles_median_over_frames'] = int(np.median([r['cycles'] for r in worst_frames]))
dbg = eng.debug_words()
worst = int(np.argmax(dbg[:, 8:14].sum(1)))
out['stage1_subphases'] = dict(names=['allpairs', 'gather', 'components', 'solve', 'fallback', 'scatter'],
                               median=[int(x) for x in np.median(dbg[:, [19, 14, 15, 16, 17, 18]], axis=0)],
                               slowest=[int(x) for x in dbg[worst, [19, 14, 15, 16, 17, 18]]])
out['phase_cycles_median'] = [int(x) for x in np.median(dbg[:, 8:14], axis=0)]
out['phase_cycles_slowest'] = [int(x) for x in dbg[int(np.argmax(dbg[:, 8:14].sum(1))), 8:14]]
out['sum_cycles_median_max'] = [int(np.median(dbg[:, 8:14].sum(1))), int(dbg[:, 8:14].sum(1).max())]
print(json.dumps(out))
