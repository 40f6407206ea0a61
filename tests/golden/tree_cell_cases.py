"""refine(function(cell)) and the TreeCell objects (mesh[i]); shared by the golden generator (reference) and
tests/test_tree_cell.py (discretize_b200).  ``evaluate(lib)`` -> {name: array}."""
import numpy as np


def evaluate(lib):
    out = {}
    for seed in range(12):
        rng = np.random.default_rng(seed)
        d = 2 + seed % 2
        n = [2 ** int(rng.integers(2, 5 if d == 3 else 6)) for _ in range(d)]
        h = [rng.random(k) * .5 + .5 for k in n]
        o = rng.random(d)
        ext = np.array([x.sum() for x in h])
        c0 = o + ext * rng.random(d)
        r1, r2 = sorted(rng.random(2) * 0.5 * ext.min())
        kind = seed % 4
        m = lib.TreeMesh(h, origin=o, diagonal_balance=bool(seed % 3 == 0))
        if seed >= 8:            # the pass also descends into cells that are divided already
            m.refine_ball(o + ext * rng.random((1, d)), [0.2 * ext.min()], [-1], finalize=False)
        calls = [0]

        def f(cell, m=m, calls=calls, c0=c0, r1=r1, r2=r2, kind=kind):
            calls[0] += 1
            r = np.linalg.norm(cell.center - c0)
            if kind == 0:
                return m.max_level if r < r1 else (m.max_level - 1 if r < r2 else 0)
            if kind == 1:
                return -1 if r < r1 else (-2 if r < r2 else cell._level)
            if kind == 2:
                return m.max_level if abs(cell.center[0] - c0[0]) < r1 else 1
            return int(m.max_level * np.exp(-r / (r2 + 1e-9)))

        m.refine(f)
        tag = f"s{seed}_{d}d"
        out[f"{tag}|cell_centers"] = m.cell_centers
        out[f"{tag}|levels"] = np.asarray(m._cell_levels_by_indexes(), dtype=np.int64)
        out[f"{tag}|n_calls"] = np.array(calls[0])
        picks = rng.integers(0, m.n_cells, 12)
        out[f"{tag}|picks"] = picks
        for k, i in enumerate(picks):
            cell = m[int(i)]
            out[f"{tag}|cell{k}|geometry"] = np.r_[cell.center, cell.origin, cell.x0, cell.h, cell.bounds]
            out[f"{tag}|cell{k}|ints"] = np.r_[cell.index, cell._level, cell.dim, cell._index_loc, cell.nodes, cell.edges,
                                              cell.faces].astype(np.int64)
            flat = []
            for nb in cell.neighbors:
                flat += [-7] + (list(nb) if isinstance(nb, (list, np.ndarray)) else [nb])
            out[f"{tag}|cell{k}|neighbors"] = np.array(flat, dtype=np.int64)
        out[f"{tag}|len"] = np.array(len(m))
        out[f"{tag}|slice"] = np.array([c.index for c in m[1:4]], dtype=np.int64)
    return out
