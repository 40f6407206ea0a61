"""QuadTree (2-D TreeMesh) cases shared by the golden generator (run on the UNMODIFIED reference) and
tests/test_quadtree_cpu.py (run on discretize_b200): meshes built through the public refinement calls, then every
count, geometry array, operator, mass matrix / derivative and interpolation matrix, entry by entry."""
import numpy as np
import scipy.sparse as sp

OPERATORS = ["face_divergence", "edge_curl", "nodal_gradient", "average_face_to_cell", "average_face_to_cell_vector",
             "average_face_x_to_cell", "average_face_y_to_cell", "average_edge_to_cell", "average_edge_to_cell_vector",
             "average_edge_x_to_cell", "average_edge_y_to_cell", "average_node_to_cell", "average_node_to_edge",
             "average_node_to_face", "average_edge_to_face", "face_x_divergence", "face_y_divergence",
             "stencil_cell_gradient", "stencil_cell_gradient_x", "stencil_cell_gradient_y", "cell_gradient",
             "cell_gradient_x", "cell_gradient_y", "average_cell_to_face", "average_cell_to_face_x",
             "average_cell_to_face_y", "average_cell_vector_to_face", "average_cell_to_total_face_x",
             "average_cell_to_total_face_y"]
COUNTS = ["n_cells", "n_nodes", "n_total_nodes", "n_hanging_nodes", "n_edges_x", "n_edges_y", "n_total_edges_x",
          "n_total_edges_y", "n_hanging_edges_x", "n_hanging_edges_y", "n_faces_x", "n_faces_y", "n_total_faces_x",
          "n_total_faces_y", "n_hanging_faces_x", "n_hanging_faces_y", "n_edges", "n_faces", "n_total_edges",
          "n_total_faces", "n_hanging_edges", "n_hanging_faces", "max_level", "n_edges_z", "n_faces_z",
          "n_edges_per_direction", "n_faces_per_direction", "vntE", "vntF"]
ARRAYS = ["cell_centers", "h_gridded", "cell_volumes", "edge_lengths", "face_areas", "nodes", "hanging_nodes", "edges_x",
          "edges_y", "faces_x", "faces_y", "hanging_edges_x", "hanging_edges_y", "hanging_faces_x", "hanging_faces_y",
          "faces", "edges", "total_nodes", "face_normals", "edge_tangents", "cell_bounds", "nodes_x", "nodes_y",
          "cell_centers_x", "cell_centers_y"]
LOCS = ["N", "Ex", "Ey", "Fx", "Fy", "CC"]


def specs():
    rng = np.random.default_rng(7)
    return [
        ("uniform_ball", [np.full(16, 1 / 16)] * 2, np.zeros(2), "ball"),
        ("stretched_ball", [rng.random(16) * .2 + .1, rng.random(8) * .3 + .2], np.array([-1.0, 2.0]), "ball"),
        ("stretched_box_points", [rng.random(32) * .2 + .1, rng.random(16) * .3 + .2], np.array([-1.0, 2.0]), "box"),
        ("uniform_level", [np.full(8, 1 / 8)] * 2, np.array([0.5, 0.5]), "level"),
    ]


def build(lib, h, origin, kind, diagonal_balance=False):
    m = lib.TreeMesh(h, origin=origin, diagonal_balance=diagonal_balance)
    ext = np.array([x.sum() for x in h])
    if kind == "ball":
        m.refine_ball(np.array([[0.3, 0.4], [0.8, 0.2]]) * ext + origin, [0.15 * ext[0], 0.1 * ext[0]], [-1, -2], finalize=False)
        m.insert_cells(np.array([[0.05, 0.9]]) * ext + origin, [-1], finalize=True)
    elif kind == "box":
        m.refine_box(origin + 0.2 * ext, origin + 0.5 * ext, -1, finalize=False)
        pts = origin + ext * np.random.default_rng(5).random((5, 2))
        m.refine_points(pts, -1, padding_cells_by_level=[1, 2], finalize=True)
    elif kind == "level":
        m.refine(2)
    return m


def _dense(a):
    if hasattr(a, "tocsr"):
        a = a.tocsr()
    if sp.issparse(a):
        return np.asarray(a.toarray(), dtype=np.float64)
    return np.asarray(a)


def _guard(fn):
    try:
        out = fn()
    except Exception as e:
        return np.array("EXC:" + type(e).__name__)
    return np.array("NONE") if out is None else _dense(out)


def points_for(mesh_ref_like, h, origin):
    """Random points (some outside) plus points exactly on cell centres / nodes / edges; the order is part of the case."""
    rng = np.random.default_rng(11)
    ext = np.array([x.sum() for x in h])
    pts = rng.random((300, 2)) * ext * 1.2 - 0.1 * ext + origin
    m = mesh_ref_like
    adv = np.r_[m.cell_centers[rng.integers(0, m.n_cells, 40)], m.nodes[rng.integers(0, m.n_nodes, 40)],
                m.edges_x[rng.integers(0, m.n_edges_x, 20)]]
    return np.r_[pts, adv][rng.permutation(400)]


def evaluate(mesh, pts, host_matrix=None):
    """name -> array for every case.  ``host_matrix(name)``: how to get an operator's matrix (default: the property)."""
    rng = np.random.default_rng(3)
    def reference_matrix(name):
        m = getattr(mesh, name)
        return m() if callable(m) else m        # average_cell_to_total_face_* are plain methods in the reference

    get = host_matrix or reference_matrix
    out = {}
    for n in COUNTS:
        out["count:" + n] = np.asarray(getattr(mesh, n), dtype=np.int64)
    for n in ARRAYS:
        out["array:" + n] = np.asarray(getattr(mesh, n), dtype=np.float64)
    out["levels"] = np.asarray(mesh._cell_levels_by_indexes(), dtype=np.int64)
    for n in OPERATORS:
        out["op:" + n] = _dense(get(n))
    nC = mesh.n_cells
    models = {"none": None, "scalar": 2.5, "iso": rng.random(nC) + .5, "aniso": rng.random((nC, 2)) + .5,
              "tensor": rng.random((nC, 3)) + np.array([2.0, 2.0, 0.0])}
    for proj, name in (("F", "face"), ("E", "edge")):
        v = rng.random(mesh.n_faces if proj == "F" else mesh.n_edges)
        for mk, model in models.items():
            for im in (False, True):
                for ix in (False, True):
                    tag = f"{name}:{mk}:{int(im)}{int(ix)}"
                    out["mass:" + tag] = _guard(
                        lambda: getattr(mesh, f"get_{name}_inner_product")(model, invert_model=im, invert_matrix=ix))

                    def deriv():
                        f = getattr(mesh, f"get_{name}_inner_product_deriv")(model, invert_model=im, invert_matrix=ix)
                        return None if f is None else f(v)
                    out["deriv:" + tag] = _guard(deriv)
    out["point2index"] = np.asarray(mesh.point2index(pts), dtype=np.int64)
    for lt in LOCS:
        out["interp:" + lt] = _dense(mesh.get_interpolation_matrix(pts, lt))
    out["interp:N:zeros_outside"] = _dense(mesh.get_interpolation_matrix(pts, "N", zeros_outside=True))
    out["interp:Fz"] = _guard(lambda: mesh.get_interpolation_matrix(pts, "Fz"))
    for n in ("cell_boundary_indices", "face_boundary_indices"):
        for i, idx in enumerate(getattr(mesh, n)):
            out[f"index:{n}#{i}"] = np.asarray(idx, dtype=np.int64)
    # boundary projections / integrals, Robin pieces, surface and line inner products: the recipe of boundary_cases.py
    import boundary_cases

    for k, v in boundary_cases.evaluate(mesh, tree=True).items():
        out["boundary:" + k] = v
    return out
