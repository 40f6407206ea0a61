"""Generates tests/golden/volume_average_small.npz from the UNMODIFIED reference (discretize.utils.volume_average)."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.refload import load_reference  # noqa: E402


def cases():
    rng = np.random.default_rng(31)
    return {
        "aligned": ([8, 8, 8], None, [4, 4, 4], None),
        "random": ([rng.random(7) + .5, rng.random(5) + .5, rng.random(6) + .5], None,
                   [rng.random(4) + .8, rng.random(9) + .3, rng.random(5) + .6], None),
        "sticks_out": ([np.ones(6) / 6, np.ones(5) / 5, np.ones(4) / 4], np.array([-0.2, 0.1, 0.0]),
                       [np.ones(5) / 5, np.ones(7) / 7, np.ones(3) / 3], np.array([0.1, -0.3, 0.2])),
    }


def main():
    ref = load_reference()
    assert ref is not None
    from discretize.utils import volume_average

    out = {}
    for tag, (hin, oin, hout, oout) in cases().items():
        hin = [np.ones(h) / h if np.isscalar(h) else np.asarray(h) for h in hin]
        hout = [np.ones(h) / h if np.isscalar(h) else np.asarray(h) for h in hout]
        m = sp.csr_matrix(volume_average(ref.TensorMesh(hin, origin=oin), ref.TensorMesh(hout, origin=oout)))
        m.sum_duplicates()
        m.sort_indices()
        for a, nm in enumerate("xyz"):
            out[f"{tag}_hin{nm}"], out[f"{tag}_hout{nm}"] = hin[a], hout[a]
        out[f"{tag}_oin"] = np.zeros(3) if oin is None else oin
        out[f"{tag}_oout"] = np.zeros(3) if oout is None else oout
        out[f"{tag}__indptr"], out[f"{tag}__indices"] = m.indptr.astype(np.int64), m.indices.astype(np.int64)
        out[f"{tag}__data"], out[f"{tag}__shape"] = m.data, np.array(m.shape, dtype=np.int64)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "volume_average_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
