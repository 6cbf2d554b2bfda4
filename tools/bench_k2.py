"""K2 micro-benchmark: the score / threshold / compaction pass of st_decode_nms on random head
maps, at the bench shape (C2: 64 x 19 320 priors) and at a shape large enough to leave the
launch-latency regime (C3: 512 x 42 840 priors, 1080p).  Background logits only (no prior
passes the threshold), so the call is the streaming pass plus an idle sort/NMS launch;
algorithmic bytes = 4 per prior (objectness logits; class logits are fetched for items whose
objectness is above the cut only).  CUDA events over rotating inputs.

    python tools/bench_k2.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from stereotracking_b200.engine import EngineConfig, StereoTrackEngine  # noqa: E402


def run(name, S, hw, reps=30, with_objects=False):
    levels = [(hw[0] // s, hw[1] // s) for s in (8, 16, 32)]
    P = sum(h * w for h, w in levels)
    eng = StereoTrackEngine(EngineConfig(n_seq=S))
    g = torch.Generator(device='cuda').manual_seed(1)
    sets = []
    for _ in range(3):
        cls = [torch.randn(S, 1, h, w, generator=g, device='cuda') - 2.0 for h, w in levels]
        obj = [torch.randn(S, 1, h, w, generator=g, device='cuda') * 0.5 - 9.0 for h, w in levels]
        box = [torch.randn(S, 4, h, w, generator=g, device='cuda') for h, w in levels]
        if with_objects:                       # ~300 candidates per frame
            for o, c in zip(obj, cls):
                m = torch.rand(o.shape, generator=g, device='cuda') < 300.0 / P
                o[m] = 3.0
                c[m] = 3.0
        sets.append((cls, box, obj))
    for i in range(5):
        eng.decode_nms(*sets[i % 3], hw)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        eng.decode_nms(*sets[i % 3], hw)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = float(np.median([ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]))
    byt = 4 * P * S
    print(json.dumps(dict(case=name, sequences=S, priors=P, candidates_per_frame=int(eng.det_counts.float().mean()),
                          us_per_call=ms * 1e3, score_pass_bytes=byt, gbs_lower_bound=byt / (ms * 1e-3) / 1e9)))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'big':          # one case, for an ncu capture
        run('C3 512 x 1920x1080, background only', 512, (1088, 1920), reps=12)
        sys.exit(0)
    run('C2 64 x 1280x720, background only', 64, (736, 1280))
    run('C2 64 x 1280x720, ~300 candidates', 64, (736, 1280), with_objects=True)
    run('C3 256 x 1920x1080, background only', 256, (1088, 1920))
    run('C3 512 x 1920x1080, background only', 512, (1088, 1920))
    run('2048 x 1920x1080, background only', 2048, (1088, 1920), reps=10)
