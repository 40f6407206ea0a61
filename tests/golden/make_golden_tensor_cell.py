"""Generates tests/golden/tensor_cell_small.npz with the UNMODIFIED reference (cases in tensor_cell_cases.py)."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle.refload import load_reference  # noqa: E402
import tensor_cell_cases as tc  # noqa: E402

if __name__ == "__main__":
    warnings.simplefilter("ignore")
    ref = load_reference()
    assert ref is not None
    out = tc.evaluate(ref)
    path = os.path.join(HERE, "tensor_cell_small.npz")
    np.savez_compressed(path, **out)
    print(len(out), "entries ->", path, os.path.getsize(path) // 1024, "KiB")
