"""Generates tests/golden/mesh_utils_small.npz with the UNMODIFIED reference's discretize.utils (cases in
mesh_utils_cases.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle.refload import load_reference  # noqa: E402
import mesh_utils_cases as mc  # noqa: E402


def main():
    ref = load_reference()
    assert ref is not None
    import discretize.utils as ru

    out = mc.evaluate(ru)
    path = os.path.join(HERE, "mesh_utils_small.npz")
    np.savez_compressed(path, **out)
    print(len(out), "entries ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
