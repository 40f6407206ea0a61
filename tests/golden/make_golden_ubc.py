"""Writes tests/golden/ubc/*.msh / *.mod with the UNMODIFIED reference's write_UBC / write_model_UBC."""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle.refload import load_reference  # noqa: E402
import ubc_cases as uc  # noqa: E402


def main():
    warnings.simplefilter("ignore")
    ref = load_reference()
    assert ref is not None
    out = os.path.join(HERE, "ubc")
    os.makedirs(out, exist_ok=True)
    for tag, mesh in {**uc.tensor_cases(ref), **uc.tree_cases(ref)}.items():
        mesh.write_UBC(os.path.join(out, tag + ".msh"), models={os.path.join(out, tag + ".mod"): uc.model_for(mesh)})
        print(tag, mesh.n_cells, "cells")


if __name__ == "__main__":
    main()
