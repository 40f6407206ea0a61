"""Geometric refinement (line, plane, triangle, vertical prism, tetrahedron, surface) and cell queries on QuadTrees and
OcTrees; shared by the golden generator (reference) and tests/test_tree_geometry_cpu.py (discretize_b200)."""
import numpy as np


def cases():
    out = []
    for seed in range(28):
        rng = np.random.default_rng(seed)
        d = 2 + seed % 2
        n = [2 ** int(rng.integers(2, 5 if d == 3 else 6)) for _ in range(d)]
        h = [rng.random(k) * .5 + .5 for k in n] if seed % 4 < 2 else [np.full(k, 1.0 / k) for k in n]
        o = rng.random(d) - 0.5
        ext = np.array([x.sum() for x in h])

        def pts(*shape, rng=rng, o=o, ext=ext, d=d):
            return o + ext * (rng.random(shape + (d,)) * 1.2 - 0.1)

        kind = seed % 7
        db = bool(rng.integers(0, 2))
        if kind == 0:
            call = ("refine_line", (pts(4), [-1, -2, -1]))
        elif kind == 1:
            call = ("refine_plane", (pts(2), rng.random((2, d)) - 0.5, [-1, -2]))
        elif kind == 2:
            call = ("refine_triangle", (pts(3, 3), [-1, -2, -1]))
        elif kind == 3:
            call = ("refine_vertical_trianglular_prism", (pts(2, 3), rng.random(2) * 0.3 * ext[-1], [-1, -2])) if d == 3 \
                else ("refine_triangle", (pts(1, 3), -1))
        elif kind == 4:
            call = ("refine_tetrahedron", (pts(2, d + 1), [-1, -1]))
        elif kind == 5:
            call = ("refine_surface", (pts(30), -1, [[1] * d, [1] * d]))
        else:
            call = ("refine_line", (pts(2), -1))
        queries = {"ball": (pts(1)[0], 0.2 * ext.min()), "line": pts(2), "aabb": (lambda x0: (x0, x0 + 0.3 * ext))(pts(1)[0]),
                   "plane": (pts(1)[0], rng.random(d) - 0.5), "tri": pts(3), "tet": pts(d + 1), "points": pts(6)}
        out.append((f"s{seed}_{call[0]}_{d}d", h, o, db, call, queries))
    return out


def evaluate(lib, case):
    tag, h, o, db, (method, args), q = case
    m = lib.TreeMesh(h, origin=o, diagonal_balance=db)
    getattr(m, method)(*args)
    d = len(h)
    out = {"cell_centers": m.cell_centers, "levels": np.asarray(m._cell_levels_by_indexes(), dtype=np.int64),
           "ball": np.sort(m.get_cells_in_ball(*q["ball"])), "line": np.sort(m.get_cells_on_line(q["line"])),
           "aabb": np.sort(m.get_cells_in_aabb(*q["aabb"])), "plane": np.sort(m.get_cells_on_plane(*q["plane"])),
           "tri": np.sort(m.get_cells_in_triangle(q["tri"])), "tet": np.sort(m.get_cells_in_tetrahedron(q["tet"])),
           "containing": np.asarray(m.get_containing_cells(q["points"]), dtype=np.int64),
           "overlapping": np.sort(m.get_overlapping_cells(np.stack(q["aabb"], axis=1).reshape(-1)))}
    if d == 3:
        out["prism"] = np.sort(m.get_cells_in_vertical_trianglular_prism(q["tri"], 0.2))
    # ordered walks: a generic segment, one that leaves the mesh, one from node to node (crosses edges / corners exactly)
    out["along_line"] = np.asarray(m.get_cells_along_line(q["line"][0], q["line"][1]), dtype=np.int64)
    out["along_line_reversed"] = np.asarray(m.get_cells_along_line(q["line"][1], q["line"][0]), dtype=np.int64)
    nodes = m.nodes
    out["along_line_nodes"] = np.asarray(m.get_cells_along_line(nodes[len(nodes) // 7], nodes[(5 * len(nodes)) // 7]), dtype=np.int64)
    return {k: np.asarray(v) for k, v in out.items()}


def image_cases():
    out = []
    for seed in range(8):
        rng = np.random.default_rng(100 + seed)
        d = 2 + seed % 2
        n = [2 ** int(rng.integers(2, 5 if d == 3 else 6)) for _ in range(d)]
        h = [rng.random(k) * .5 + .5 for k in n]
        o = rng.random(d)
        img = np.zeros(n)
        for _ in range(3):
            lo = [int(rng.integers(0, k)) for k in n]
            hi = [int(rng.integers(a + 1, k + 1)) for a, k in zip(lo, n)]
            img[tuple(slice(a, b) for a, b in zip(lo, hi))] = rng.integers(1, 4)
        out.append((f"img{seed}_{d}d", h, o, bool(seed % 3 == 0), img if seed % 2 else img.reshape(-1, order="F"), seed))
    return out


def evaluate_image(lib, case):
    tag, h, o, db, img, seed = case
    m = lib.TreeMesh(h, origin=o, diagonal_balance=db)
    m.refine_image(img)
    act = np.random.default_rng(seed).random(m.n_cells) > 0.3
    out = {"cell_centers": m.cell_centers, "levels": np.asarray(m._cell_levels_by_indexes(), dtype=np.int64),
           "fill": np.array(m.fill)}
    for direction in ("xd", "xu", "yd", "yu", "zd", "zu"):
        out["boundary_" + direction] = np.asarray(m.get_boundary_cells(direction=direction))
        out["active_boundary_" + direction] = np.asarray(m.get_boundary_cells(act, direction)[0])
    for key in ("CC", "N", "Fx", "Ey"):
        for a, t in enumerate(m.get_tensor(key)):
            out[f"tensor_{key}_{a}"] = np.asarray(t)
    return out
