"""mesh_builder_xyz / refine_tree_xyz / active_from_xyz / extract_core_mesh cases shared by the golden generator
(reference) and tests/test_mesh_utils.py (discretize_b200).  ``evaluate(utils_module)`` -> {name: array}."""
import warnings

import numpy as np


def evaluate(u):
    out = {}
    rng = np.random.default_rng(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for d in (2, 3):
            xyz = rng.random((40, d)) * np.r_[400., 300., 50.][:d] + np.r_[1000., 2000., 100.][:d]
            h = [20., 20., 10.][:d]
            pad = [[100, 150], [80, 80], [200, 0]][:d]

            def put_mesh(tag, m):
                for a, hh in enumerate(m.h):
                    out[f"{tag}:h{a}"] = np.asarray(hh)
                out[f"{tag}:origin"] = np.asarray(m.origin, dtype=np.float64)
                if m._meshType == "TREE" and m.finalized:
                    out[f"{tag}:cell_centers"] = m.cell_centers

            for mt in ("tensor", "tree"):
                put_mesh(f"{d}d:builder_{mt}", u.mesh_builder_xyz(xyz, h, padding_distance=pad, depth_core=100., mesh_type=mt))
            base = u.mesh_builder_xyz(xyz, h, padding_distance=pad, mesh_type="tensor")
            put_mesh(f"{d}d:builder_on_base", u.mesh_builder_xyz(xyz[:10], h, padding_distance=pad, base_mesh=base, mesh_type="tree"))
            for method, kw in (("radial", dict(octree_levels=[2, 2, 1])),
                               ("box", dict(octree_levels=[1, 2, 2], octree_levels_padding=[1, 1, 2])),
                               ("surface", dict(octree_levels=[1, 2, 1], octree_levels_padding=[0, 1, 1], max_distance=150.))):
                m = u.mesh_builder_xyz(xyz, h, padding_distance=pad, depth_core=100., mesh_type="tree")
                m = u.refine_tree_xyz(m, xyz, method=method, finalize=True, **kw)
                put_mesh(f"{d}d:refine_{method}", m)
                if method == "surface":
                    for gr in ("CC", "N"):
                        for meth in ("linear", "nearest"):
                            out[f"{d}d:active_tree_{gr}_{meth}"] = u.active_from_xyz(m, xyz, gr, meth)
            for gr in ("CC", "N"):
                out[f"{d}d:active_tensor_{gr}"] = u.active_from_xyz(base, xyz, gr)
            lim = np.c_[xyz.min(axis=0) + 20, xyz.max(axis=0) - 20]
            act, core = u.extract_core_mesh(lim, base)
            out[f"{d}d:core_active"] = act
            put_mesh(f"{d}d:core", core)
    x = rng.random((3, 4))
    out["mkvc"], out["mkvc2"] = u.mkvc(x), u.mkvc(x, 2)
    out["ndgrid"] = u.ndgrid(np.r_[1., 2.], np.r_[3., 4., 5.])
    out["unpack_widths"] = u.unpack_widths([(10.0, 3, -1.5), (10.0, 2), (10.0, 3, 1.5)])
    for shape in (20, (12, 9), (6, 7, 5)):
        out[f"random_model_{np.size(shape)}d"] = u.random_model(shape, random_seed=3, its=30, bounds=[1, 5])
    return out
