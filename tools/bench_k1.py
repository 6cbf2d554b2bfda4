"""K1 micro-benchmark: st_disp2depth at the C2 shape (64 x 1280x720 u16 -> 64 x 736x1280 f32),
CUDA events over rotating inputs (each launch reads a fresh 118 MB and writes 241 MB, so
nothing useful survives in the 126 MB L2).  Also feeds every u16 code through the kernel and
compares with the float32 expression of the reference, and times a torch memset / copy /
read of the same byte counts for context.

    python tools/bench_k1.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import numpy as np
    import torch
    from stereotracking_b200.engine import EngineConfig, StereoTrackEngine
    S, H, W = 64, 720, 1280
    eng = StereoTrackEngine(EngineConfig(n_seq=1), device='cuda:0')
    # every raw code, bit-exact
    codes = torch.arange(65536, dtype=torch.int32).to(torch.uint16).reshape(1, 64, 1024).cuda()
    depth, _, _ = eng.disp2depth(codes)
    raw = np.arange(65536, dtype=np.float32)
    disp = np.where(raw == 65535, np.float32(0), raw * np.float32(0.0625)).astype(np.float32)
    ref = (np.float32(1.0) / (disp + np.float32(1e-6))) * np.float32(160.0)
    got = depth.cpu().numpy().reshape(-1)
    bad = int((got.view(np.uint32) != ref.view(np.uint32)).sum())
    g = torch.Generator(device='cuda').manual_seed(1)
    bufs = [torch.randint(0, 4000, (S, H, W), generator=g, device='cuda', dtype=torch.int32).to(torch.uint16)
            for _ in range(4)]
    out = [torch.empty(S, 736, 1280, dtype=torch.float32, device='cuda') for _ in range(2)]
    for i in range(5):
        eng.disp2depth(bufs[i % 4], out=out[i % 2])
    torch.cuda.synchronize()
    n = 40
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        eng.disp2depth(bufs[i % 4], out=out[i % 2])
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(n)])
    byt = S * (2 * H * W + 4 * 736 * 1280)
    ref_times = {}
    if True:
        # what the memory system does on plain torch ops of the same sizes
        def timed(fn, reps=20):
            for _ in range(3):
                fn()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
            e[0].record()
            for r in range(reps):
                fn(); e[r + 1].record()
            torch.cuda.synchronize()
            return float(np.median([e[r].elapsed_time(e[r + 1]) for r in range(reps)]) * 1e3)
        a = torch.empty(byt // 8, dtype=torch.float32, device='cuda'); b2 = torch.empty_like(a)
        big = [torch.empty(S, 736, 1280, dtype=torch.float32, device='cuda') for _ in range(2)]
        k = [0]
        def fill():
            k[0] += 1; big[k[0] % 2].zero_()
        def copy():
            b2.copy_(a)
        def read():
            k[0] += 1; bufs[k[0] % 4].view(torch.int16).max()
        ref_times = dict(memset_241MB_us=timed(fill), copy_180MB_to_180MB_us=timed(copy), read_max_118MB_us=timed(read))
    print(json.dumps(dict(mismatching_codes=bad,
                          us_median=float(np.median(ms) * 1e3), us_min=float(ms.min() * 1e3),
                          gbs_median=byt / (np.median(ms) * 1e-3) / 1e9, bytes=byt, **ref_times)))


if __name__ == '__main__':
    one()
