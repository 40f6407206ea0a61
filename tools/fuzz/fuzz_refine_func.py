"""Randomised comparison against the LIVE reference (this container only: needs /root/reference through oracle.refload).
Not part of the test suite -- the pinned subsets live in tests/golden; this is how the wider parity claim in DESIGN.md was
checked.  Run: python tools/fuzz/fuzz_refine_func.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, warnings
warnings.simplefilter("ignore")
from oracle.refload import load_reference
import discretize_b200 as dz
ref=load_reference()
bad=0
for seed in range(80):
    rng=np.random.default_rng(seed)
    d=2+seed%2
    n=[2**int(rng.integers(2,5 if d==3 else 6)) for _ in range(d)]
    h=[rng.random(k)*.5+.5 for k in n]; o=rng.random(d)
    db=bool(seed%3==0)
    ext=np.array([x.sum() for x in h]); c0=o+ext*rng.random(d); r1,r2=sorted(rng.random(2)*0.5*ext.min())
    kind=seed%4
    out=[]
    for lib in (ref,dz):
        m=lib.TreeMesh(h,origin=o,diagonal_balance=db)
        calls=[0]
        def f(cell,m=m,calls=calls):
            calls[0]+=1
            r=np.linalg.norm(cell.center-c0)
            if kind==0: return m.max_level if r<r1 else (m.max_level-1 if r<r2 else 0)
            if kind==1: return -1 if r<r1 else (-2 if r<r2 else cell._level)
            if kind==2: return m.max_level if abs(cell.center[0]-c0[0])<r1 else 1
            return int(m.max_level*np.exp(-r/(r2+1e-9)))
        if seed%5==0: m.insert_cells(o+ext*rng.random((2,d)) if lib is ref else pts_saved, [m.max_level,m.max_level-1], finalize=False) if False else None
        m.refine(f); out.append((m,calls[0]))
    (mr,nr),(m,nm)=out
    ok = mr.n_cells==m.n_cells and np.array_equal(mr.cell_centers,m.cell_centers) and nr==nm
    if not ok: bad+=1; print("MISMATCH",seed,d,db,kind,mr.n_cells,m.n_cells,nr,nm)
print("done mismatches",bad)
