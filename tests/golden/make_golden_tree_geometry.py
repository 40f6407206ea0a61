"""Generates tests/golden/tree_geometry_small.npz from the UNMODIFIED reference: the trees produced by the geometric
refinement calls and the answers of the cell queries (cases in tree_geometry_cases.py)."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle.refload import load_reference  # noqa: E402
import tree_geometry_cases as gc  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    ref = load_reference()
    assert ref is not None
    out = {}
    for case in gc.cases():
        for k, v in gc.evaluate(ref, case).items():
            out[f"{case[0]}|{k}"] = v
    for case in gc.image_cases():
        for k, v in gc.evaluate_image(ref, case).items():
            out[f"{case[0]}|{k}"] = v
    path = os.path.join(HERE, "tree_geometry_small.npz")
    np.savez_compressed(path, **out)
    print(len(out), "entries ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
