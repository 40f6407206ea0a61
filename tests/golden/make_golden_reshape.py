"""Generates tests/golden/reshape_digests.npz with the UNMODIFIED reference (cases in reshape_cases.py)."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle.refload import load_reference  # noqa: E402
import reshape_cases as rc  # noqa: E402

if __name__ == "__main__":
    warnings.simplefilter("ignore")
    ref = load_reference()
    assert ref is not None
    keys, digests = rc.evaluate(ref)
    path = os.path.join(HERE, "reshape_digests.npz")
    np.savez_compressed(path, keys=keys, digests=digests)
    print(len(keys), "combinations ->", path, os.path.getsize(path) // 1024, "KiB")
