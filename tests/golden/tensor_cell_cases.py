"""TensorCell objects (mesh[i], tuple and slice indexing, iteration) for 1-D / 2-D / 3-D meshes; shared by the golden
generator (reference) and tests/test_tensor_cell.py (discretize_b200)."""
import numpy as np


def evaluate(lib):
    rng = np.random.default_rng(0)
    out = {}
    for h in ([rng.random(4) + .5, rng.random(3) + .5, rng.random(5) + .5], [rng.random(4) + .5, rng.random(3) + .5],
              [rng.random(6) + .5]):
        d = len(h)
        m = lib.TensorMesh(h, origin=-np.arange(1., d + 1))
        tag = f"{d}d"
        cells = list(m)
        out[f"{tag}|len"] = np.array(len(m))
        out[f"{tag}|geometry"] = np.array([np.r_[c.h, c.origin, c.center, c.bounds] for c in cells])
        out[f"{tag}|index"] = np.array([np.r_[c.index, c.index_unraveled, c.mesh_shape, c.dim] for c in cells], dtype=np.int64)
        for prop in ("nodes", "edges", "faces"):
            out[f"{tag}|{prop}"] = np.array([getattr(c, prop) for c in cells], dtype=np.int64)
        out[f"{tag}|neighbors"] = np.array([x for c in cells for x in [-7] + list(c.neighbors)], dtype=np.int64)
        out[f"{tag}|repr0"] = np.array(repr(cells[0]))
        picks = {"neg": m[-1], "slice": m[1:5], "rev": m[::-2]}
        if d == 3:
            picks.update({"tuple": m[1, 2, 3], "tuple_neg": m[-1, 0, -2], "mixed": m[1:3, :, 2], "mixed2": m[:, -1, ::2],
                          "nb": m[2].get_neighbors(m)})
            out[f"{tag}|empty"] = np.array(str(m[5, 0:0, 1]))
        if d == 2:
            picks.update({"tuple": m[1, 2], "mixed": m[1:3, -1]})
        for k, v in picks.items():
            v = v if isinstance(v, list) else [v]
            out[f"{tag}|pick_{k}"] = np.array([c.index for c in v], dtype=np.int64)
        for k, bad in (("str", "a"), ("float", 1.5), ("short", (0,) * (d + 1))):
            try:
                m[bad]
                out[f"{tag}|bad_{k}"] = np.array("ok")
            except Exception as e:
                out[f"{tag}|bad_{k}"] = np.array("EXC:" + type(e).__name__)
    return out
