"""Back-to-back timing of K1b (st_pack_image) at the C2 shape: achieved GB/s against the
measured HBM copy peak.  Algorithmic bytes per frame: 3*H*W read + 12*H'*W' written."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereotracking_b200.engine import EngineConfig, StereoTrackEngine  # noqa: E402

S, H, W = 64, 720, 1280
eng = StereoTrackEngine(EngineConfig(n_seq=S))
imgs = [torch.randint(0, 256, (S, H, W, 3), dtype=torch.uint8, device='cuda') for _ in range(3)]   # > L2 together
outs = [torch.empty(S, 3, 736, 1280, device='cuda') for _ in range(3)]
res = {}
for name, kw in (('plain', {}), ('norm_swap', dict(mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375), swap_rb=True))):
    for i in range(6):
        eng.pack_image(imgs[i % 3], out=outs[i % 3], **kw)
    torch.cuda.synchronize()
    n = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        eng.pack_image(imgs[i % 3], out=outs[i % 3], **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    nbytes = S * (3 * H * W + 12 * 736 * 1280)
    res[name] = dict(ms=round(ms, 4), GBps=round(nbytes / ms / 1e6, 1), bytes=nbytes)
peak = 6453.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
    peak = float(peak.get('hbm_gbs', 6453.4)) if isinstance(peak, dict) else 6453.4
except Exception:
    peak = 6453.4
for v in res.values():
    v['frac_of_peak'] = round(v['GBps'] / peak, 3)
print(json.dumps(dict(kernel='k1b_pack_image_v4', workload='C2 S=64', peak_GBps=peak, **res)))
