"""The boundary / weak-form / surface-product cases shared by the golden generator (run on the UNMODIFIED reference)
and by tests/test_boundary_ops.py (run on discretize_b200): one recipe, two libraries, compared entry by entry."""
import copy

import numpy as np
import scipy.sparse as sp


def tensor_specs():
    rng = np.random.default_rng(77)
    return [("t3", [rng.random(4) + .5, rng.random(3) + .5, rng.random(5) + .5], np.array([-1.0, -2.0, -3.0])),
            ("t2", [rng.random(5) + .5, rng.random(4) + .5], np.array([0.5, -2.0])),
            ("t1", [rng.random(6) + .5], np.array([-1.0]))]


def tree_specs():
    rng = np.random.default_rng(78)
    return [("o_uniform", [np.full(8, 0.5)] * 3, np.zeros(3)),
            ("o_stretched", [rng.random(8) * .3 + .3, rng.random(4) * .3 + .4, rng.random(8) * .2 + .2], np.array([-1.0, 2.0, 0.5]))]


def build_tree(cls, h, origin):
    m = cls(h, origin=origin)
    m.refine_ball(np.array([[0.3, 0.4, 0.5], [2.2, 1.1, 0.7]]) + origin, [0.6, 0.5], [3, 3], finalize=False)
    m.insert_cells(np.array([[0.1, 0.1, 0.1]]) + origin, [m.max_level], finalize=True)
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
    except Exception as e:  # the exception TYPE is part of the interface
        return np.array("EXC:" + type(e).__name__)
    if out is None:
        return np.array("NONE")
    if isinstance(out, (tuple, list)) and not isinstance(out, np.ndarray):
        return [_dense(o) for o in out]
    return _dense(out)


def evaluate(mesh, tree=False):
    """name -> dense array (or 'EXC:<type>' / 'NONE' markers) for every case, on either library's mesh."""
    rng = np.random.default_rng(123)
    d = mesh.dim
    out = {}

    def put(name, fn):
        val = _guard(fn)
        if isinstance(val, list):
            for i, v in enumerate(val):
                out[f"{name}#{i}"] = v
        else:
            out[name] = val

    geom = ["faces", "edges", "face_normals", "edge_tangents", "cell_bounds", "cell_nodes", "boundary_faces",
            "boundary_nodes", "boundary_face_outward_normals", "project_face_to_boundary_face",
            "project_node_to_boundary_node", "boundary_face_scalar_integral", "boundary_node_vector_integral"]
    if d > 1:
        geom += ["boundary_edges", "project_edge_to_boundary_edge", "boundary_edge_vector_integral"]
    if tree:
        geom += ["hanging_edges_x", "hanging_edges_y", "hanging_faces_x", "hanging_faces_y",
                 "total_nodes", "vntF", "vntE", "nodes_x", "max_used_level",
                 "permute_cells", "permute_faces", "permute_edges", "edge_nodes",
                 "average_node_to_edge_x", "average_node_to_edge_y", "average_node_to_face_x", "average_node_to_face_y"]
        if d == 3:
            geom += ["hanging_edges_z", "hanging_faces_z", "cell_centers_z", "average_node_to_edge_z",
                     "average_node_to_face_z"]
    else:
        geom += ["h_gridded", "face_x_areas", "face_y_areas", "face_z_areas", "edge_x_lengths", "edge_y_lengths",
                 "edge_z_lengths", "cell_boundary_indices", "face_boundary_indices"]
    for name in geom:
        put(name, lambda name=name: getattr(mesh, name))

    nbf = mesh.project_face_to_boundary_face.shape[0]
    nbn = mesh.project_node_to_boundary_node.shape[0]
    robin = [(0.0, 1.0, 0.0), (0.5, 2.0, 1.5),
             (rng.random(nbf), rng.random(nbf) + .5, rng.random(nbf)),
             (rng.random(nbn), rng.random(nbn) + .5, rng.random(nbn)),
             (0.3, rng.random(nbf) + .5, rng.random((nbf, 2))),
             (rng.random(nbn), 1.0, rng.random((nbn, 3))),
             (1.0, 0.0, 1.0)]
    for i, (al, be, ga) in enumerate(robin):
        put(f"edge_div_robin{i}", lambda: mesh.edge_divergence_weak_form_robin(al, be, ga))
        if not (np.ndim(al) and len(al) == nbn):
            put(f"cell_grad_robin{i}", lambda: mesh.cell_gradient_weak_form_robin(al, be, ga))

    if not tree:
        bcs = ["neumann", "dirichlet", [["neumann", "dirichlet"]] + ["dirichlet"] * (d - 1), [["dirichlet", "neumann"]] * d,
               "robin", ["neumann"] * (d + 1)]
        for i, bc in enumerate(bcs):
            put(f"bc_proj{i}", lambda: mesh.get_BC_projections(copy.deepcopy(bc)))
        put("bc_proj_simple", lambda: mesh.get_BC_projections_simple())
        put("bc_proj_nodal", lambda: mesh.get_BC_projections("neumann", discretization="N"))
        pts = rng.random((20, d)) * 4 - 3
        if d == 1:
            pts = pts[:, 0]
        put("point2index", lambda: mesh.point2index(pts))
    else:
        pts = rng.random((30, d)) * 4 - 1 + np.asarray(mesh.origin)
        put("is_inside", lambda: mesh.is_inside(pts))
    for loc in ("CC", "N", "Fx", "Ex"):
        put("closest_" + loc, lambda: mesh.closest_points_index(pts, loc))
    if d > 1:
        fv, ev = rng.random((mesh.n_faces, d)), rng.random((mesh.n_edges, d))
        put("project_face_vector", lambda: mesh.project_face_vector(fv))
        put("project_edge_vector", lambda: mesh.project_edge_vector(ev))
        put("project_face_vector_bad", lambda: mesh.project_face_vector(fv[:-1]))

    kinds = ["get_face_inner_product_surface", "get_edge_inner_product_line"]
    if d > 1:
        kinds.append("get_edge_inner_product_surface")
    for meth in kinds:
        n = mesh.n_edges if meth.endswith("line") else mesh.n_faces
        v = rng.random(mesh.n_faces if meth.startswith("get_face") else mesh.n_edges)
        models = {"none": None, "float": 2.5, "array1": np.array([2.5]), "full": rng.random(n) + .5, "bad": np.ones(n + 1)}
        for mk, model in models.items():
            for im in (False, True):
                for ix in (False, True):
                    tag = f"{meth}:{mk}:{int(im)}{int(ix)}"
                    put(tag, lambda: getattr(mesh, meth)(model, invert_model=im, invert_matrix=ix))

                    def deriv():
                        f = getattr(mesh, meth + "_deriv")(model, invert_model=im, invert_matrix=ix)
                        return None if f is None else f(v)
                    put(tag + ":deriv", deriv)
    return out
