"""Randomised comparison against the LIVE reference (this container only: needs /root/reference through oracle.refload).
Not part of the test suite -- the pinned subsets live in tests/golden; this is how the wider parity claim in DESIGN.md was
checked.  Run: python tools/fuzz/fuzz_va.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, warnings, scipy.sparse as sp
warnings.simplefilter("ignore")
from oracle.refload import load_reference
import discretize_b200 as dz
ref=load_reference()
from importlib import import_module
rva = import_module(ref.__name__+".utils").volume_average
def canon(a):
    a=sp.csr_matrix(a.tocsr() if hasattr(a,'tocsr') else a).copy(); a.sum_duplicates(); a.eliminate_zeros(); a.sort_indices(); return a
bad=0
for seed in range(60):
    rng=np.random.default_rng(seed)
    d=2+seed%2
    def mk(lib, r):
        kind=r.integers(0,2)
        n=[2**r.integers(2,5) for _ in range(d)]
        h=[r.random(k)*.5+.5 for k in n] if r.integers(0,2) else [np.full(k,4.0/k) for k in n]
        o=r.random(d)*2-1 if r.integers(0,3) else np.zeros(d)
        if kind==0:
            return lib.TensorMesh(h,origin=o)
        m=lib.TreeMesh(h,origin=o)
        ext=np.array([x.sum() for x in h])
        m.refine_ball(o+ext*r.random((2,d)), r.random(2)*0.4*ext.min(), [int(r.integers(1,m.max_level+1)) for _ in range(2)], finalize=True)
        return m
    pair=[]
    for lib in (ref,dz):
        r=np.random.default_rng(seed+500)
        pair.append((mk(lib,r),mk(lib,r)))
    (ri,ro),(mi,mo)=pair
    if ri._meshType=="TENSOR" and ro._meshType=="TENSOR": continue
    try:
        B=canon(rva(ri,ro))
    except Exception as e:
        print("ref failed",seed,type(e).__name__,str(e)[:60]); continue
    A=canon(dz.utils.volume_average(mi,mo))
    pat=A.shape==B.shape and np.array_equal(A.indptr,B.indptr) and np.array_equal(A.indices,B.indices)
    err=abs(A-B).max() if pat else -1
    if not pat or err>1e-13:
        bad+=1; print("MISMATCH",seed,d,ri._meshType,ro._meshType,A.shape,A.nnz,B.nnz,err)
print("done, mismatches",bad)
