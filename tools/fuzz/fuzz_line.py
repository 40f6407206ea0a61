"""Randomised comparison against the LIVE reference (this container only: needs /root/reference through oracle.refload).
Not part of the test suite -- the pinned subsets live in tests/golden; this is how the wider parity claim in DESIGN.md was
checked.  Run: python tools/fuzz/fuzz_line.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, warnings
warnings.simplefilter("ignore")
from oracle.refload import load_reference
import discretize_b200 as dz
ref=load_reference()
bad=0; n=0
for seed in range(60):
    rng=np.random.default_rng(seed)
    d=2+seed%2
    k=[2**int(rng.integers(2,5 if d==3 else 6)) for _ in range(d)]
    h=[rng.random(x)*.5+.5 for x in k] if seed%4<2 else [np.full(x,1.0/x) for x in k]
    o=rng.random(d); ext=np.array([x.sum() for x in h])
    ms=[]
    for lib in (ref,dz):
        m=lib.TreeMesh(h,origin=o,diagonal_balance=bool(seed%3==0))
        r2=np.random.default_rng(seed+7)
        m.refine_ball(o+ext*r2.random((2,d)), r2.random(2)*0.4*ext.min(), [-1,-2]); ms.append(m)
    mr,m=ms
    for trial in range(12):
        p0=o+ext*rng.random(d); p1=o+ext*rng.random(d)
        if trial%4==1: p1[0]=p0[0]                      # axis-parallel
        if trial%4==2:                                  # through nodes: diagonal between two mesh nodes
            nodes=mr.nodes; p0=nodes[rng.integers(0,len(nodes))].copy(); p1=nodes[rng.integers(0,len(nodes))].copy()
        if trial%4==3: p1=o+ext*(rng.random(d)*1.4-0.2) # leaves the mesh
        n+=1
        try: R=list(mr.get_cells_along_line(p0,p1))
        except Exception as e: R="EXC "+type(e).__name__
        try: M=list(m.get_cells_along_line(p0,p1))
        except Exception as e: M="EXC "+type(e).__name__+str(e)[:50]
        if R!=M:
            bad+=1
            if bad<6: print("MISMATCH",seed,d,trial,R if isinstance(R,str) else R[:12],M if isinstance(M,str) else M[:12])
print("lines",n,"mismatches",bad)
