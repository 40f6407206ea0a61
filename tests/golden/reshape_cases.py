"""Every combination of (input, x_type, out_type, return_format) for TensorMesh.reshape on a 3-D, a 2-D and a 1-D mesh,
reduced to a digest per combination (exception type, or shapes + SHA-1 of the values); shared by the golden generator
(reference) and tests/test_reshape.py (discretize_b200)."""
import hashlib
import itertools

import numpy as np

TYPES = ["CC", "N", "F", "Fx", "Fy", "Fz", "E", "Ex", "Ey", "Ez", "cell_centers", "faces_x", "G"]


def _digest(out):
    if out is None:
        return "NONE"
    parts = out if isinstance(out, tuple) else (out,)
    sha = hashlib.sha1()
    shapes = []
    for p in parts:
        if p is None:
            shapes.append("None")
            continue
        p = np.ascontiguousarray(p, dtype=np.float64)
        shapes.append(str(p.shape))
        sha.update(p.tobytes())
    return ("T" if isinstance(out, tuple) else "A") + "|".join(shapes) + ":" + sha.hexdigest()[:12]


def evaluate(lib):
    keys, digests = [], []
    for h in ([3, 4, 2], [3, 4], [5]):
        rng = np.random.default_rng(len(h))
        m = lib.TensorMesh(h)
        d = m.dim
        inputs = {"cc": rng.random(m.n_cells), "n": rng.random(m.n_nodes), "f": rng.random(m.n_faces),
                  "e": rng.random(m.n_edges), "fx": rng.random(m.n_faces_x), "ccM": rng.random(m.shape_cells),
                  "vec": rng.random((m.n_cells, d)), "list": [rng.random(m.n_cells)], "bad": rng.random(7)}
        if d > 1:
            inputs["ey"] = rng.random(m.n_edges_y)
        for (k, x), xt, ot, fmt in itertools.product(inputs.items(), TYPES, TYPES, ("V", "M", "Q")):
            try:
                dig = _digest(m.reshape(x, xt, ot, fmt))
            except Exception as e:
                dig = "EXC:" + type(e).__name__
            keys.append(f"{d}d|{k}|{xt}|{ot}|{fmt}")
            digests.append(dig)
    return np.array(keys), np.array(digests)
