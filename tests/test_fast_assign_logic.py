"""Soundness of the assignment fast path (csrc/lapjv.cuh fast_assign_warp).

The CUDA kernel solves the association as a rectangular assignment with the
Hungarian method and accepts the result only with a uniqueness certificate
computed from the optimal duals (no re-routing along tight edges); otherwise
it falls back to the literal emulation of the reference's JV solver.  This file
restates that fast path in plain Python (test infrastructure, not product
code) and checks on the CPU that
  * a certified result always equals the oracle's lap.lapjv restatement, and
  * the certificate agrees with brute-force enumeration of all matchings:
    certified <=> the optimum is unique.
The GPU parity tests then check the kernel itself against the oracle.
"""
import numpy as np

from oracle.lapjv import lapjv


def fast(cost, limit, eps=1e-9):
    R,C = cost.shape; m=R+C; half=limit/2; L=limit; INF=1e300
    c=cost.astype(np.float64)
    u=np.zeros(R); v=np.zeros(m); p=-np.ones(m,int); way=np.zeros(m,int)
    active=[i for i in range(R) if c[i].min()<L]
    for i in range(R):
        if i not in active: u[i]=half
    def a(i,j):
        if j<C: return c[i,j]-half
        return half if j-C==i else INF
    for row in active:
        minv=np.full(m,INF); used=np.zeros(m,bool)
        j0=m
        while True:
            if j0<m: used[j0]=True
            i0=row if j0==m else p[j0]
            delta=INF; j1=-1
            for j in range(m):
                if used[j]: continue
                if j>=C and j-C!=i0:
                    if minv[j]<delta: delta=minv[j]; j1=j
                    continue
                cur=a(i0,j)-u[i0]-v[j]
                if cur<minv[j]: minv[j]=cur; way[j]=j0
                if minv[j]<delta: delta=minv[j]; j1=j
            for j in range(m):
                if used[j]: u[p[j]]+=delta; v[j]-=delta
                else: minv[j]-=delta
            u[row]+=delta
            j0=j1
            if p[j0]<0: break
        while j0!=m:
            jp=way[j0]; p[j0]= row if jp==m else p[jp]; j0=jp
    sigma=[C+i for i in range(R)]
    for j in range(C):
        if p[j]>=0: sigma[p[j]]=j
    arcs=[]
    for i in range(R):
        for j in range(C):
            if p[j]!=i and (c[i,j]-half)-u[i]-v[j] <= eps: arcs.append((sigma[i],j))
        if sigma[i]!=C+i and half-u[i]-v[C+i] <= eps: arcs.append((sigma[i],C+i))
    assigned=[ (p[j]>=0) if j<C else (sigma[j-C]==j) for j in range(m)]
    vac=[assigned[j] and v[j]>=-eps for j in range(m)]
    reach=[not a_ for a_ in assigned]
    ch=True
    while ch:
        ch=False
        for (f,t) in arcs:
            if reach[t] and not reach[f]: reach[f]=True; ch=True
    unsure=any(vac[j] and reach[j] for j in range(m))
    alive=list(assigned); ch=True
    while ch:
        ch=False
        ho=[False]*m
        for (f,t) in arcs:
            if alive[f] and alive[t]: ho[f]=True
        for j in range(m):
            if alive[j] and not ho[j]: alive[j]=False; ch=True
    unsure = unsure or any(alive)
    return p[:C].copy(), (not unsure), len(arcs)


def brute(cost, limit):
    R,C=cost.shape; half=limit/2
    best=[]; 
    cols=list(range(C))
    def rec(i,used,tot,assign):
        if i==R:
            best.append((tot+half*(C-len(used)), tuple(assign))); return
        rec(i+1,used,tot+half,assign+[-1])
        for j in cols:
            if j not in used: rec(i+1,used|{j},tot+float(cost[i,j]),assign+[j])
    rec(0,frozenset(),0.0,[])
    best.sort()
    return best


def _cases(rng, n, max_dim):
    for trial in range(n):
        R, C = rng.integers(1, max_dim, 2)
        kind = trial % 3
        if kind == 0:
            cost = rng.uniform(0, 1.3, (R, C)).astype(np.float32)
        elif kind == 1:
            cost = (rng.integers(0, 5, (R, C)) * 0.25).astype(np.float32)      # many exact ties
        else:
            cost = rng.uniform(0.5, 1.3, (R, C)).astype(np.float32)
            k = min(R, C)
            cost[np.arange(k), np.arange(k)] = rng.uniform(0.05, 0.5, k)
            if k > 2 and rng.random() < 0.5:
                cost[1] = cost[0]                                             # duplicate track
        yield cost


def test_certified_results_equal_the_reference_solver():
    rng = np.random.default_rng(1)
    certified = 0
    for cost in _cases(rng, 450, 14):
        y, cert, _ = fast(cost, 0.9)
        if cert:
            certified += 1
            np.testing.assert_array_equal(y, lapjv(cost, 0.9)[1])
    assert certified > 200


def test_certificate_is_equivalent_to_uniqueness():
    rng = np.random.default_rng(2)
    n_unique = 0
    for cost in _cases(rng, 240, 6):
        _, cert, _ = fast(cost, 0.9)
        b = brute(cost, 0.9)
        unique = len(b) == 1 or b[1][0] - b[0][0] > 1e-9
        assert cert == unique
        n_unique += unique
    assert 40 < n_unique < 230
