"""An OcTree DC-resistivity style workflow, written exactly as it would be for the reference -- only the import differs.

    from discretize import TreeMesh; from discretize.utils import mesh_builder_xyz, active_from_xyz
becomes
    from discretize_b200 import TreeMesh; from discretize_b200.utils import mesh_builder_xyz, active_from_xyz

1. build an OcTree around a survey (points on a topography surface), refine below the surface and around the electrodes;
2. flag the cells below the topography;
3. assemble A = G^T Me(sigma) G as an operator and solve A phi = q with scipy's conjugate gradients -- every ``A @ x`` runs
   on the GPU (numpy in, numpy out);
4. read the potential at the electrodes with the interpolation operator.

Run: python examples/octree_dc_workflow.py        (needs a GPU for steps 3-4)
"""
import numpy as np
import scipy.sparse.linalg as spla

import discretize_b200 as dz
from discretize_b200.utils import active_from_xyz, mesh_builder_xyz


def topography(xy):
    return 10.0 * np.exp(-((xy[:, 0] - 50.0) ** 2 + (xy[:, 1] + 20.0) ** 2) / 80.0 ** 2)


def build_mesh():
    gx, gy = np.meshgrid(np.linspace(-200.0, 200.0, 21), np.linspace(-200.0, 200.0, 21))
    surface = np.c_[gx.reshape(-1), gy.reshape(-1)]
    surface = np.c_[surface, topography(surface)]
    electrodes = np.c_[np.linspace(-100.0, 100.0, 9), np.zeros(9)]
    electrodes = np.c_[electrodes, topography(electrodes) - 2.5]          # buried a little below the surface
    mesh = mesh_builder_xyz(surface, [10.0, 10.0, 5.0], padding_distance=[[300, 300], [300, 300], [300, 100]],
                            depth_core=100.0, mesh_type="tree", tree_diagonal_balance=True)
    mesh.refine_surface(surface, level=-2, padding_cells_by_level=[[0, 0, 2], [0, 0, 2]], finalize=False)
    mesh.refine_points(electrodes, level=-1, padding_cells_by_level=[2, 2, 2], finalize=True)
    active = active_from_xyz(mesh, surface, grid_reference="CC")
    return mesh, electrodes, active


def main(verbose=True):
    mesh, electrodes, active = build_mesh()
    sigma = np.where(active, 1e-2, 1e-8)                       # air is (almost) an insulator
    block = active & (np.linalg.norm(mesh.cell_centers - np.r_[0.0, 0.0, -30.0], axis=1) < 25.0)
    sigma[block] = 1.0                                         # a conductive body
    G = mesh.nodal_gradient
    Me = mesh.get_edge_inner_product(sigma)
    A = G.T @ Me @ G                                           # operator algebra; nothing is assembled
    Q = mesh.get_interpolation_matrix(electrodes, "N")         # electrodes <- nodes
    q = Q.T @ np.r_[1.0, np.zeros(7), -1.0]                    # a dipole: +1 A at the first, -1 A at the last electrode
    calls = [0]

    def count(_):
        calls[0] += 1

    Gm = G.tocsr()
    diag = Gm.multiply(Gm).T @ Me.diagonal()                   # diag(G^T Me G) = (G o G)^T diag(Me): Jacobi preconditioner
    jacobi = spla.LinearOperator(A.shape, matvec=lambda r: r / diag, dtype=np.float64)
    phi, info = spla.cg(A, q, rtol=1e-8, maxiter=4000, M=jacobi, callback=count)
    residual = np.linalg.norm(A @ phi - q) / np.linalg.norm(q)
    voltages = Q @ phi
    if verbose:
        print(f"{mesh.n_cells} cells ({active.sum()} below the surface), {mesh.n_nodes} nodes, "
              f"{mesh.n_hanging_nodes} hanging; CG iterations {calls[0]}, relative residual {residual:.2e}")
        print("potential differences to the first electrode:", np.round(voltages - voltages[0], 4))
    return mesh, residual, voltages, info


if __name__ == "__main__":
    main()
