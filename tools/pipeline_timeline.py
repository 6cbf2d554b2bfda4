"""Timeline of the frame pipeline at the bench workload (C2, 64 sequences): begin / end of
every stage of a few steady-state frames, from CUDA events on the stage streams
(st_pipeline_trace).  Shows which stage waits for what.

    python tools/pipeline_timeline.py [frames=12] [depth=5] [depth_source=map|raw] [graph=0|1]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from stereotracking_b200.engine import EngineConfig, FramePipeline, StereoTrackEngine  # noqa: E402
from stereotracking_b200.synthetic import CONFIGS, SyntheticStereo  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 12
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 5
source = sys.argv[3] if len(sys.argv) > 3 else 'map'
graph = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cfg, S, W = CONFIGS['C2'], 64, 24
gen = SyntheticStereo(cfg, S, W + F, seed=20240604 + 1000, device='cuda')
parts = [gen.frames(f, min(f + 2, W + F)) for f in range(0, W + F, 2)]
raw = torch.cat([p['raw'] for p in parts])
heads = [[torch.cat([p[k][l] for p in parts]) for l in range(3)] for k in ('cls', 'bbox', 'obj')]
fid = torch.arange(W + F, dtype=torch.int32, device='cuda')[:, None].repeat(1, S).contiguous()
eng = StereoTrackEngine(EngineConfig(n_seq=S, max_candidates=2048))
pipe = FramePipeline(eng, (S, cfg.height, cfg.width), [c.shape[-2:] for c in heads[0]], depth=depth, depth_source=source)
if graph:
    pipe.use_graphs(True)
print(f'depth_source={source} depth={depth} graph={graph}' + (' (stages 0-2 = the whole pre segment)' if graph else ''))
for f in range(W + F):
    if f == W:
        pipe.trace(F)
    pipe.submit(fid[f], raw[f], [c[f] for c in heads[0]], [c[f] for c in heads[1]], [c[f] for c in heads[2]])
tr = pipe.read_trace()
t0 = tr[:, :, 0].min()
print('frame ' + ' '.join(f'{n[:14]:>22s}' for n in pipe.STAGES) + '   (begin-end, us since the first traced stage)')
for k in range(len(tr)):
    print(f'{W + k:5d} ' + ' '.join(f'{(tr[k, s, 0] - t0) * 1e3:10.0f}-{(tr[k, s, 1] - t0) * 1e3:<10.0f}' for s in range(5)))
dur = (tr[:, :, 1] - tr[:, :, 0]) * 1e3
print('stage duration us (median over frames): ' + ', '.join(f'{n}={np.median(dur[:, s]):.0f}' for s, n in enumerate(pipe.STAGES)))
gap = (tr[1:, 3, 0] - tr[:-1, 3, 1]) * 1e3
print(f'tracker idle between frames us (median): {np.median(gap):.0f};  wait of the tracker after its K3 finished: '
      f'{np.median((tr[:, 3, 0] - tr[:, 2, 1]) * 1e3):.0f}')
print(f'frame period us: {np.median(np.diff(tr[:, 4, 1])) * 1e3:.0f}')
