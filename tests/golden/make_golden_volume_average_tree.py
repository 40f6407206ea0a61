"""Generates tests/golden/volume_average_tree_small.npz from the UNMODIFIED reference: discretize.utils.volume_average for
every pairing that involves a TreeMesh (same tensor base, different bases, partially and fully outside the input)."""
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle.refload import load_reference  # noqa: E402
import volume_average_tree_cases as cases  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    ref = load_reference()
    assert ref is not None
    from discretize.utils import volume_average

    out = {}
    rng = np.random.default_rng(5)
    for tag, (mesh_in, mesh_out) in cases.build(ref).items():
        m = sp.csr_matrix(volume_average(mesh_in, mesh_out)).copy()
        m.sum_duplicates()
        m.eliminate_zeros()
        m.sort_indices()
        out[f"{tag}|indptr"], out[f"{tag}|indices"] = m.indptr.astype(np.int32), m.indices.astype(np.int32)
        out[f"{tag}|data"], out[f"{tag}|shape"] = m.data, np.array(m.shape, dtype=np.int64)
        v = rng.random(mesh_in.n_cells)
        out[f"{tag}|values"], out[f"{tag}|averaged"] = v, volume_average(mesh_in, mesh_out, v)
    path = os.path.join(HERE, "volume_average_tree_small.npz")
    np.savez_compressed(path, **out)
    print(len(out), "entries ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
