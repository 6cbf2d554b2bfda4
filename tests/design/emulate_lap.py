"""Python emulation of csrc/lapjv.cuh solve_warp (same parallel formulation),
used to debug the device algorithm offline against oracle/lapjv."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
BIG = 1e300

def solve(cost, limit, max_iter=100000):
    R, C = cost.shape; n = R + C; half = limit / 2.0
    c = cost.astype(np.float64)
    def at(i, j):
        if i < R: return c[i, j] if j < C else half
        return half if j < C else 0.0
    x = [-1] * n; y = [0] * n; v = [0.0] * n; unique = [1] * n
    for j in range(n):
        best, arg = BIG, 0
        if j < C:
            for i in range(R):
                if c[i, j] < best: best, arg = c[i, j], i
            if n > R and half < best: best, arg = half, R
        else:
            if R > 0: best, arg = half, 0
            if n > R and 0.0 < best: best, arg = 0.0, R
        v[j], y[j] = best, arg
    for j in range(n): x[y[j]] = max(x[y[j]], j)
    for j in range(n):
        i = y[j]
        if x[i] != j: unique[i] = 0; y[j] = -1
    free = [0] * n; nfree = 0
    for i in range(n):
        if x[i] < 0: free[nfree] = i; nfree += 1
        elif unique[i]:
            j1 = x[i]
            m = min([at(i, j2) - v[j2] for j2 in range(n) if j2 != j1] + [BIG])
            v[j1] -= m
    for _pass in range(2):
        if nfree == 0: break
        cur = nxt = visits = 0
        while cur < nfree:
            visits += 1
            fr = free[cur]; cur += 1
            red = [(at(fr, j) - v[j], j) for j in range(n)]
            srt = sorted(red)
            (u1, j1) = srt[0]; (u2, j2) = srt[1] if n > 1 else (BIG, -1)
            i0 = y[j1]; v1n = v[j1] - (u2 - u1); lowers = v1n < v[j1]
            if visits < cur * n:
                if lowers: v[j1] = v1n
                elif i0 >= 0 and j2 >= 0: j1 = j2; i0 = y[j2]
                if i0 >= 0:
                    if lowers: cur -= 1; free[cur] = i0
                    else: free[nxt] = i0; nxt += 1
            elif i0 >= 0:
                free[nxt] = i0; nxt += 1
            x[fr] = j1; y[j1] = fr
            if visits > max_iter: raise RuntimeError('ARR runaway')
        nfree = nxt
    for f in range(nfree):
        start = free[f]
        cols = list(range(n)); pred = [start] * n
        d = [at(start, j) - v[j] for j in range(n)]
        lo = hi = n_ready = 0; final_j = -1; guard = 0
        while final_j == -1:
            guard += 1
            if guard > 2 * n + 8: raise RuntimeError(f'augment guard f={f} start={start} lo={lo} hi={hi}')
            if lo == hi:
                n_ready = lo
                run = d[cols[lo]]
                trig = []
                for k in range(lo + 1, n):
                    dv = d[cols[k]]
                    if dv < run: trig.append((k, True))
                    elif dv == run: trig.append((k, False))
                    run = min(run, dv)
                h2 = lo + 1
                for k, strict in trig:
                    if strict: h2 = lo
                    j = cols[k]; cols[k] = cols[h2]; cols[h2] = j; h2 += 1
                hi = h2
                cand = -1
                for k in range(lo, hi):
                    if y[cols[k]] < 0: cand = k
                if cand >= 0: final_j = cols[cand]
            if final_j == -1:
                while lo != hi:
                    js = cols[lo]; lo += 1
                    i = y[js]; mind = d[js]; h = at(i, js) - v[js] - mind
                    hits = []; first_free = None
                    for k in range(hi, n):
                        j = cols[k]; r = at(i, j) - v[j] - h
                        if r < d[j]:
                            d[j] = r; pred[j] = i
                            if r == mind:
                                hits.append(k)
                                if y[j] < 0 and first_free is None: first_free = k
                    if first_free is not None: final_j = cols[first_free]; break
                    h2 = hi
                    for k in hits:
                        j = cols[k]; cols[k] = cols[h2]; cols[h2] = j; h2 += 1
                    hi = h2
                if final_j != -1: lo = n_ready
        mind = d[cols[lo]]
        for k in range(n_ready):
            j = cols[k]; v[j] += d[j] - mind
        i, j, steps = -1, final_j, 0
        while i != start and steps <= n:
            steps += 1
            i = pred[j]; y[j] = i; j, x[i] = x[i], j
    xx = np.asarray(x[:R]); yy = np.asarray(y[:C])
    xx[xx >= C] = -1; yy[yy >= R] = -1
    return xx, yy

if __name__ == '__main__':
    from oracle.lapjv import lapjv
    import oracle.tracker as ot
    from stereotracking_b200.synthetic import SceneConfig, SyntheticStereo
    from oracle.model import StereoTrackOracle
    mats = []
    orig = ot.OCSortDisparityOracle._assign
    def spy(self, d):
        if d.size: mats.append(d.copy())
        return orig(self, d)
    ot.OCSortDisparityOracle._assign = spy
    cfg = SceneConfig(360, 640, 5, 30, invalid_prob=0.30, n_clutter=200, heavy_ego=True)
    fr = SyntheticStereo(cfg, 2, 8, seed=12).frames(0, 8)
    for s in range(2):
        o = StereoTrackOracle()
        for f in range(8):
            o.step(f, fr['raw'][f, s].numpy(), [c[f, s].numpy() for c in fr['cls']], [c[f, s].numpy() for c in fr['bbox']], [c[f, s].numpy() for c in fr['obj']])
    print(len(mats), 'matrices')
    for k, m in enumerate(mats):
        try:
            ex, ey = solve(m, 0.9)
        except RuntimeError as e:
            print(k, m.shape, 'EMULATION FAILED', e); continue
        rx, ry = lapjv(m, 0.9)
        ok = (ex == rx).all() and (ey == ry).all()
        print(k, m.shape, 'ok' if ok else 'MISMATCH')
