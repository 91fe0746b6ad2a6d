"""The tie rule of the device matching kernel (csrc/attnshift_assign.cu: ``better`` + the xor-shuffle reduction over 32 lanes),
restated in python and run against scipy's sequential scan on tie-heavy matrices: the parallel arg-min must pick the column
the published algorithm picks, or integer-valued costs would give a different (equally cheap) matching.  No GPU needed; the
kernel itself is tested against scipy in tests/test_gpu_assign.py."""
import numpy as np
from scipy.optimize import linear_sum_assignment
INF=float('inf')
def better(a,b):
    if a[2]<0: return False
    if b[2]<0: return True
    if a[0]!=b[0]: return a[0]<b[0]
    if a[1]!=b[1]: return a[1]>b[1]
    return a[2]>b[2] if a[1] else a[2]<b[2]
def solve(cost_gp, P):
    # cost_gp [G,P]
    G=cost_gp.shape[0]
    tr = G<P
    nr,nc=(G,P) if tr else (P,G)
    cst=(lambda i,j: float(cost_gp[i,j])) if tr else (lambda i,j: float(cost_gp[j,i]))
    u=[0.0]*nr; v=[0.0]*nc; col4row=[-1]*nr; row4col=[-1]*nc; path=[-1]*nc
    for cur in range(nr):
        SR=[0]*nr; SC=[0]*nc; spc=[INF]*nc; remaining=[nc-j-1 for j in range(nc)]
        num=nc; i=cur; sink=-1; mv=0.0
        while sink<0:
            SR[i]=1
            lanes=[(INF,0,-1)]*32
            for lane in range(32):
                best=(INF,0,-1)
                for it in range(lane,num,32):
                    j=remaining[it]
                    r=mv+cst(i,j)-u[i]-v[j]
                    if r<spc[j]: path[j]=i; spc[j]=r
                    c=(spc[j],1 if row4col[j]<0 else 0,it)
                    if better(c,best): best=c
                lanes[lane]=best
            o=16
            while o>0:
                new=list(lanes)
                for lane in range(32):
                    other=lanes[lane^o]
                    if better(other,lanes[lane]): new[lane]=other
                lanes=new; o>>=1
            assert all(l==lanes[0] for l in lanes)
            best=lanes[0]; mv=best[0]
            assert best[2]>=0 and mv<INF
            j=remaining[best[2]]
            if row4col[j]<0: sink=j
            else: i=row4col[j]
            SC[j]=1; remaining[best[2]]=remaining[num-1]; num-=1
        for r in range(nr):
            if r==cur: u[r]+=mv
            elif SR[r]: u[r]+=mv-spc[col4row[r]]
        for j in range(nc):
            if SC[j]: v[j]-=mv-spc[j]
        j=sink
        while True:
            r=path[j]; row4col[j]=r; t=col4row[r]; col4row[r]=j; j=t
            if r==cur: break
    out=[]
    for p in range(P):
        g = row4col[p] if tr else col4row[p]
        if g>=0: out.append((p,g))
    return out


def test_parallel_tie_rule_matches_scipy():
    rng = np.random.default_rng(0)
    for trial in range(150):
        P = int(rng.integers(1, 70)); G = int(rng.integers(1, 70))
        kind = trial % 3
        if kind == 0:
            c = rng.integers(0, 4, size=(P, G)).astype(np.float32)
        elif kind == 1:
            c = rng.standard_normal((P, G)).astype(np.float32)
        else:
            c = np.full((P, G), 2.0, np.float32)
        rows, cols = linear_sum_assignment(c)
        assert sorted(zip(rows.tolist(), cols.tolist())) == solve(np.ascontiguousarray(c.T), P), (trial, P, G, kind)
