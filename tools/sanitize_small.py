"""Small run of every kernel family for compute-sanitizer (memcheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import discretize_b200 as dz

rng = np.random.default_rng(0)
for shape in ((37, 5, 9), (64, 3, 4), (301, 2, 3)):
    mesh = dz.TensorMesh([rng.random(n) + 0.5 for n in shape])
    sigma, mui = np.exp(rng.standard_normal(mesh.nC)), np.exp(rng.standard_normal(mesh.nC))
    G, C, D = mesh.nodal_gradient, mesh.edge_curl, mesh.face_divergence
    Me, Mf = mesh.get_edge_inner_product(sigma), mesh.get_face_inner_product(mui)
    phi, e, u = rng.standard_normal(mesh.nN), rng.standard_normal(mesh.nE), rng.standard_normal(mesh.nF)
    (G.T @ Me @ G) @ phi
    (C.T @ Mf @ C + Me) @ e
    D @ u, C @ e, G @ phi, C.T @ u, mesh.aveE2CC @ e, mesh.aveN2CC.T @ rng.standard_normal(mesh.nC)
    mesh.get_face_inner_product(np.c_[rng.random((mesh.nC, 3)) + 2, rng.random((mesh.nC, 3)) * 0.1]) @ u
    mesh.get_edge_inner_product_deriv(sigma)(e) @ rng.standard_normal(mesh.nC)
    D.tocsr()
    ext = np.array([h.sum() for h in mesh.h])
    Q = mesh.get_interpolation_matrix(rng.random((1000, 3)) * ext, "Ey")
    Q @ e, Q.T @ rng.standard_normal(1000)
t = dz.TreeMesh([16, 16, 16], diagonal_balance=True)
t.refine_points(rng.random((4, 3)), -1, [1, 1], finalize=True)
t.face_divergence @ rng.standard_normal(t.nF), t.edge_curl.T @ rng.standard_normal(t.nF)
t.get_edge_inner_product(rng.random(t.nC) + 1) @ rng.standard_normal(t.nE)
print("sanitize_small done")
