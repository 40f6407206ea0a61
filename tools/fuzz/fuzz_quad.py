"""Randomised comparison against the LIVE reference (this container only: needs /root/reference through oracle.refload).
Not part of the test suite -- the pinned subsets live in tests/golden; this is how the wider parity claim in DESIGN.md was
checked.  Run: python tools/fuzz/fuzz_quad.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, warnings
warnings.simplefilter("ignore")
from oracle.refload import load_reference
import discretize_b200 as dz
ref=load_reference()
bad=0
for seed in range(120):
    rng=np.random.default_rng(seed)
    nx,ny=2**rng.integers(2,6),2**rng.integers(2,6)
    h=[rng.random(nx)*.5+.5, rng.random(ny)*.5+.5] if seed%2 else [np.full(nx,1.0/nx), np.full(ny,1.0/ny)]
    o=rng.random(2)-0.5
    db=bool(seed%3==0)
    ext=np.array([x.sum() for x in h])
    ops=[]
    for k in range(rng.integers(1,4)):
        kind=rng.integers(0,5)
        if kind==0: ops.append(("insert_cells",(o+ext*(rng.random((rng.integers(1,6),2))*1.2-0.1), None)))
        elif kind==1: ops.append(("refine_ball",(o+ext*rng.random((2,2)), rng.random(2)*0.3*ext.min())))
        elif kind==2: ops.append(("refine_box",(o+ext*rng.random((1,2))*0.5, o+ext*(0.5+rng.random((1,2))*0.5))))
        elif kind==3: ops.append(("refine_points",(o+ext*rng.random((4,2)), None)))
        else: ops.append(("refine_bounding_box",(o+ext*(0.3+0.4*rng.random((5,2))), None)))
    meshes=[]
    for lib in (ref,dz):
        m=lib.TreeMesh(h,origin=o,diagonal_balance=db)
        r2=np.random.default_rng(seed+1000)
        for name,(a,b) in ops:
            lv=int(r2.integers(1,m.max_level+1))
            if name=="insert_cells":
                lvls=r2.integers(1,m.max_level+1,size=a.shape[0])
                m.insert_cells(a,lvls,finalize=False)
            elif name=="refine_ball": m.refine_ball(a,b,[lv,max(lv-1,1)],finalize=False)
            elif name=="refine_box": m.refine_box(a,b,[lv],finalize=False)
            elif name=="refine_points": m.refine_points(a,lv,padding_cells_by_level=[1,1],finalize=False)
            else: m.refine_bounding_box(a,max(lv,2),padding_cells_by_level=[1,2],finalize=False)
        m.finalize(); meshes.append(m)
    mr,m=meshes
    ok = mr.n_cells==m.n_cells and np.array_equal(mr.cell_centers,m.cell_centers) and mr.n_hanging_edges==m.n_hanging_edges and mr.n_nodes==m.n_nodes and np.array_equal(mr.nodes,m.nodes) and np.array_equal(mr.edges_x,m.edges_x) and np.array_equal(mr.hanging_edges_y,m.hanging_edges_y)
    if ok and m.n_cells>1:
        A=m._host_matrix("edge_curl"); B=mr.edge_curl
        ok = abs(A-B).max()<1e-12*max(1,abs(B).max())
        A=m._host_matrix("average_cell_to_face"); B=mr.average_cell_to_face
        ok = ok and A.shape==B.shape and abs(A-B).max()<1e-13
    if not ok:
        bad+=1; print("MISMATCH seed",seed,(nx,ny),db,[n for n,_ in ops], mr.n_cells, m.n_cells)
print("fuzz done, mismatches:",bad)
