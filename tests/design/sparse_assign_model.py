"""CPU model of the sparse certified solver of csrc/lapjv.cuh (sparse_assign_warp + arcs_unique), statement by
statement, for analysing refused problems dumped from the GPU (tools/k4_tie_dump.py).  Analysis tool, not a test."""
import sys
import numpy as np

EPS = 1e-9


def sparse_assign(M, L, verbose=False):
    half = L / 2.0
    adm = M.astype(np.float64) < L
    rlist = np.nonzero(adm.any(1))[0]
    clist = np.nonzero(adm.any(0))[0]
    cmap = {int(c): k for k, c in enumerate(clist)}
    nR, nC = len(rlist), len(clist)
    edges = [[(cmap[int(j)], np.float32(M[r, j])) for j in np.nonzero(adm[r])[0]] for r in rlist]
    INF = 1e300
    VIRT = 1 << 30
    u = np.zeros(nR); v = np.zeros(nC); minv = np.full(nC, INF); p = np.full(nC, -1); way = np.full(nC, 0x7fffffff)
    used = np.zeros(nC, bool)
    want = np.zeros(nR, int)
    for i in range(nR):
        bc, bj = np.float32(3e38), 0x7fffffff
        for j, c in edges[i]:
            if c < bc or (c == bc and j < bj):
                bc, bj = c, j
        u[i] = float(bc) - half
        want[i] = bj
        way[bj] = min(way[bj], i)
    free = []
    for i in range(nR):
        if way[want[i]] == i:
            p[want[i]] = i
        else:
            free.append(i)
    for row in free:
        fcol = []
        j0, i0 = VIRT, row
        dmin = half - u[row]
        drow, dcol = row, VIRT
        to_dummy = False
        while True:
            ui0 = u[i0]
            for j, c in edges[i0]:
                if not used[j]:
                    cur = (float(c) - half) - ui0 - v[j]
                    mv = minv[j]
                    if not (mv < INF):
                        fcol.append(j)
                    if cur < mv:
                        minv[j] = cur; way[j] = j0
            best, bj = INF, -1
            for j in fcol:
                if used[j]:
                    continue
                if minv[j] < best or (minv[j] == best and j < bj):
                    best, bj = minv[j], j
            to_dummy = bj < 0 or dmin < best
            delta = dmin if to_dummy else best
            for j in fcol:
                if used[j]:
                    u[p[j]] += delta; v[j] -= delta
                else:
                    minv[j] -= delta
            u[row] += delta
            dmin -= delta
            if to_dummy:
                break
            j0 = bj
            used[j0] = True
            i0 = p[j0]
            if i0 < 0:
                break
            cand = half - u[i0]
            if cand < dmin:
                dmin, drow, dcol = cand, i0, j0
        jc = dcol if to_dummy else j0
        while jc != VIRT:
            jp = way[jc]
            p[jc] = row if jp == VIRT else p[jp]
            jc = jp
        for j in fcol:
            minv[j] = INF; used[j] = False
    sigma = np.array([nC + i for i in range(nR)])
    for j in range(nC):
        if p[j] >= 0:
            sigma[p[j]] = j
    arcs = []
    for i in range(nR):
        for j, c in edges[i]:
            if j != sigma[i] and (float(c) - half) - u[i] - v[j] <= EPS:
                arcs.append((int(sigma[i]), j, 'edge', int(rlist[i]), int(clist[j])))
    for i in range(nR):
        if sigma[i] != nC + i and half - u[i] <= EPS:
            arcs.append((int(sigma[i]), nC + i, 'dummy', int(rlist[i]), -1))
    assigned = lambda node: (p[node] >= 0) if node < nC else (sigma[node - nC] == node)
    vnode = lambda node: v[node] if node < nC else 0.0
    reach = [not assigned(a[1]) for a in arcs]
    changed = True
    while changed:
        changed = False
        for k, a in enumerate(arcs):
            if not reach[k]:
                for k2, b in enumerate(arcs):
                    if reach[k2] and b[0] == a[1]:
                        reach[k] = True; changed = True
    unsure1 = [a for k, a in enumerate(arcs) if reach[k] and vnode(a[0]) >= -EPS]
    alive = [assigned(a[1]) for a in arcs]
    changed = True
    while changed:
        changed = False
        for k, a in enumerate(arcs):
            if alive[k] and not any(alive[k2] and b[0] == a[1] for k2, b in enumerate(arcs)):
                alive[k] = False; changed = True
    unsure2 = [a for k, a in enumerate(arcs) if alive[k]]
    if verbose:
        print('  nR', nR, 'nC', nC, 'free rows', [int(rlist[i]) for i in free], 'arcs', arcs)
        print('  unsure by (1)', unsure1, ' by (2)', unsure2)
    pairs = sorted((int(rlist[p[j]]), int(clist[j])) for j in range(nC) if p[j] >= 0)
    return pairs, bool(unsure1 or unsure2), dict(u=u, v=v, rlist=rlist, clist=clist, sigma=sigma, p=p)


if __name__ == '__main__':
    d = np.load(sys.argv[1])
    for k in d.files:
        print(k)
        pairs, unsure, st = sparse_assign(d[k], 0.7, verbose=True)
        print('  pairs', len(pairs), 'refused', unsure)
