"""Randomised comparison against the LIVE reference (this container only: needs /root/reference through oracle.refload).
Not part of the test suite -- the pinned subsets live in tests/golden; this is how the wider parity claim in DESIGN.md was
checked.  Run: python tools/fuzz/fuzz_geom.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, warnings
warnings.simplefilter("ignore")
from oracle.refload import load_reference
import discretize_b200 as dz
ref=load_reference()
bad=0; ncase=0
for seed in range(150):
    rng=np.random.default_rng(seed)
    d=2+seed%2
    n=[2**rng.integers(2,5 if d==3 else 6) for _ in range(d)]
    h=[rng.random(k)*.5+.5 for k in n] if seed%4<2 else [np.full(k,1.0/k) for k in n]
    o=rng.random(d)-0.5
    ext=np.array([x.sum() for x in h])
    P=lambda *shape: o+ext*(rng.random(shape+(d,))*1.2-0.1)
    kind=seed%7
    db=bool(rng.integers(0,2))
    meshes=[]
    args=None
    if kind==0: args=("refine_line",(P(4),[-1,-2,-1]))
    elif kind==1: args=("refine_plane",(P(2),rng.random((2,d))-0.5,[-1,-2]))
    elif kind==2: args=("refine_triangle",(P(3,3),[-1,-2,-1]))
    elif kind==3 and d==3: args=("refine_vertical_trianglular_prism",(P(2,3),rng.random(2)*0.3*ext[2],[-1,-2]))
    elif kind==4 and d==3: args=("refine_tetrahedron",(P(2,4),[-1,-1]))
    elif kind==5:
        pts=P(30)
        args=("refine_surface",(pts,-1,[[1]*d,[1]*d]))
    elif kind==6: args=("refine_line",(P(2),-1))
    if args is None: continue
    ncase+=1
    for lib in (ref,dz):
        m=lib.TreeMesh(h,origin=o,diagonal_balance=db)
        getattr(m,args[0])(*args[1])
        meshes.append(m)
    mr,m=meshes
    ok = mr.n_cells==m.n_cells and np.array_equal(mr.cell_centers,m.cell_centers)
    # queries on the finished mesh
    q=[]
    try:
        c=P(1)[0]; r=0.2*ext.min()
        q.append(("ball", np.array_equal(np.sort(mr.get_cells_in_ball(c,r)), m.get_cells_in_ball(c,r))))
        seg=P(2); q.append(("line", np.array_equal(np.sort(mr.get_cells_on_line(seg)), m.get_cells_on_line(seg))))
        x0=P(1)[0]; x1=x0+0.3*ext; q.append(("aabb", np.array_equal(np.sort(mr.get_cells_in_aabb(x0,x1)), m.get_cells_in_aabb(x0,x1))))
        og=P(1)[0]; nm=rng.random(d)-0.5; q.append(("plane", np.array_equal(np.sort(mr.get_cells_on_plane(og,nm)), m.get_cells_on_plane(og,nm))))
        tri=P(3); q.append(("tri", np.array_equal(np.sort(mr.get_cells_in_triangle(tri)), m.get_cells_in_triangle(tri))))
        if d==3:
            q.append(("prism", np.array_equal(np.sort(mr.get_cells_in_vertical_trianglular_prism(tri,0.2)), m.get_cells_in_vertical_trianglular_prism(tri,0.2))))
            tet=P(4); q.append(("tet", np.array_equal(np.sort(mr.get_cells_in_tetrahedron(tet)), m.get_cells_in_tetrahedron(tet))))
        # sortedness of the reference's own answer (claim: traversal order == cell order)
        a=np.asarray(mr.get_cells_in_aabb(x0,x1)); q.append(("ref sorted", np.array_equal(a,np.sort(a))))
    except Exception as e:
        q.append(("EXC "+type(e).__name__+str(e)[:80], False))
    if not ok or not all(v for _,v in q):
        bad+=1; print("MISMATCH seed",seed,d,args[0],db,mr.n_cells,m.n_cells,[k for k,v in q if not v])
print("cases",ncase,"mismatches",bad)
