"""ctypes binding of libdzb.so (the C ABI declared in include/dzb.h).

There is no CPU fallback: if the library is missing, or no CUDA device is present
when an operator is applied, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdzb.so")

ABI_VERSION = 1


class DzbTensorMesh(C.Structure):
    _fields_ = [
        ("nx", C.c_int64), ("ny", C.c_int64), ("nz", C.c_int64),
        ("hx", C.c_void_p), ("hy", C.c_void_p), ("hz", C.c_void_p),
        ("rhx", C.c_void_p), ("rhy", C.c_void_p), ("rhz", C.c_void_p),
    ]


# operator ids (enum DzbOp)
OP_DIV, OP_CURL, OP_GRAD = 0, 1, 2
OP_AVE_F2CC, OP_AVE_F2CCV, OP_AVE_FX2CC, OP_AVE_FY2CC, OP_AVE_FZ2CC = 3, 4, 5, 6, 7
OP_AVE_E2CC, OP_AVE_E2CCV, OP_AVE_EX2CC, OP_AVE_EY2CC, OP_AVE_EZ2CC = 8, 9, 10, 11, 12
OP_AVE_N2CC = 13
OP_AVE_CC2F, OP_AVE_CCV2F, OP_AVE_CC2E, OP_AVE_E2F, OP_AVE_N2E, OP_AVE_N2F = 14, 15, 16, 17, 18, 19
OP_CELL_GRAD, OP_CELL_GRAD_X, OP_CELL_GRAD_Y, OP_CELL_GRAD_Z = 20, 21, 22, 23
OP_DIV_X, OP_DIV_Y, OP_DIV_Z = 24, 25, 26
OP_GRAD_SQ, OP_CURL_SQ = 27, 28
CG_SCALE = 64  # DzbCellGradFlags: bits 0..5 = Dirichlet on x-lo, x-hi, y-lo, y-hi, z-lo, z-hi

MODEL_SCALAR, MODEL_ISO, MODEL_ANISO, MODEL_TENSOR = 0, 1, 2, 3
FLAG_INVERT_MODEL, FLAG_INVERT_MATRIX, FLAG_EDGES = 1, 2, 4

_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_D = C.c_double
_MP = C.POINTER(DzbTensorMesh)

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "dzb_version": [],
    "dzb_last_error": [],
    "dzb_tm_apply": [_MP, _I, _I, _P, _P, _P],
    "dzb_tm_shape": [_MP, _I, C.POINTER(_L), C.POINTER(_L)],
    "dzb_tm_csr_count": [_MP, _I, _I, _P, _P],
    "dzb_tm_csr_fill": [_MP, _I, _I, _P, _P, _P, _P],
    "dzb_tm_apply_flags": [_MP, _I, _I, _I, _P, _P, _P],
    "dzb_tm_csr_count_flags": [_MP, _I, _I, _I, _P, _P],
    "dzb_tm_csr_fill_flags": [_MP, _I, _I, _I, _P, _P, _P, _P],
    "dzb_tm_mass_diag": [_MP, _I, _I, _P, _D, _P, _P],
    "dzb_tm_tensor_invert": [_L, _P, _P, _P],
    "dzb_tm_mass_apply": [_MP, _I, _I, _P, _D, _P, _P, _P],
    "dzb_tm_mass_tensor_csr_count": [_MP, _I, _P, _P],
    "dzb_tm_mass_tensor_csr_fill": [_MP, _I, _P, _P, _P, _P, _P],
    "dzb_tm_mass_deriv_apply": [_MP, _I, _I, _P, _D, _P, _I, _P, _P, _P],
    "dzb_tm_dc_nodal": [_MP, _P, _P, _P, _P],
    "dzb_tm_dc_nodal_range": [_MP, _P, _P, _P, _L, _L, _P],
    "dzb_tm_dc_nodal_multi": [_MP, _P, _P, _P, _I, _L, _P],
    "dzb_tm_dc_cell": [_MP, _P, _I, _P, _P, _P],
    "dzb_tm_dc_cell_mi": [_MP, _P, _I, _P, _P, _P],
    "dzb_tm_maxwell": [_MP, _P, _P, _P, _P, _P],
    "dzb_tm_maxwell_range": [_MP, _P, _P, _P, _P, _L, _L, _P],
    "dzb_tm_maxwell_range2": [_MP, _P, _P, _P, _P, _L, _L, _L, _L, _P],
    "dzb_tm_maxwell_cplx": [_MP, _P, _P, _D, _D, _P, _P, _P],
    "dzb_tm_dc_nodal_dot": [_MP, _P, _P, _P, _P, _L, _P, _P],
    "dzb_tm_dc_nodal_range_peer": [_MP, _P, _P, _P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _P, _P, _L, _L, _P],
    "dzb_tm_maxwell_range_peer_sync": [_MP, _P, _P, _P, _P, _L, _L, _L, _P, _L, _L, _L, _P, _P, _P, _P, _L, _L, _P],
    "dzb_tm_maxwell_range_peer": [_MP, _P, _P, _P, _P, _L, _L, _L, _P, _L, _L, _L, _P, _L, _L, _P],
    "dzb_interp_locate": [_P, _L, _P, _L, _P, _L, _P, _L, _L, _P, _I, _P, _P],
    "dzb_interp_apply": [_P, _I, _P, _P, _L, _P, _P, _P],
    "dzb_interp_apply_t": [_P, _I, _P, _P, _L, _P, _P, _P],
    "dzb_interp_apply_perm": [_P, _I, _P, _P, _P, _L, _P, _P, _P],
    "dzb_interp_apply_t_perm": [_P, _I, _P, _P, _P, _L, _P, _P, _P],
    "dzb_gather_rows": [_L, _P, _P, _P, _P],
    "dzb_interp_locate_compact": [_P, _L, _P, _L, _P, _L, _P, _L, _L, _P, _I, _P, _P, _P, _P],
    "dzb_interp_apply_compact": [_P, _I, _P, _P, _P, _L, _L, _P, _L, _P, _P, _P],
    "dzb_interp_apply_t_compact": [_P, _I, _P, _P, _P, _L, _L, _P, _P, _L, _P, _P, _P],
    "dzb_interp_expand": [_P, _I, _P, _P, _P, _L, _L, _L, _P, _I, _P, _P],
    "dzb_interp_fused": [_P, _L, _P, _L, _P, _L, _P, _L, _L, _P, _P, _P, _P],
    "dzb_tree_spmv": [_L, _P, _I, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _I, _P],
    "dzb_tree_ell_spmv": [_L, _I, _L, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _P, _P, _P],
    "dzb_tree_locate": [_P, _P, _P, _P, _I, _P, _L, _P, _L, _P, _P],
    "dzb_tree_rows_add": [_L, _P, _P, _I, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P],
    "dzb_tree_sell_spmv": [_L, _P, _I, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _P, _P],
    "dzb_tree_ell_spmv_perm": [_L, _I, _L, _P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _P, _P, _P, _P],
    "dzb_csr_spmv": [_L, _P, _I, _P, _P, _P, _P, _I, _P],
    "dzb_diag_scale": [_L, _P, _P, _P, _I, _P],
    "dzb_tree_mass_tensor": [_L, _I, _P, _P, _P, _P, _P, _P],
    "dzb_tree_mass_tensor_adj": [_L, _I, _P, _P, _P, _P, _P, _P],
    "dzb_solver_workspace_bytes": [],
    "dzb_dot": [_L, _P, _P, _P, _P, _I, _P],
    "dzb_cg_update": [_L, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "dzb_cg_direction": [_L, _P, _P, _P, _P, _P],
    "dzb_points_inside": [_P, _L, C.POINTER(_D), C.POINTER(_D), C.POINTER(_D), _P, _P],
}
_RESTYPES = {"dzb_last_error": C.c_char_p, "dzb_solver_workspace_bytes": C.c_int64}

_lib = None


class DzbError(RuntimeError):
    pass


def load():
    """Load libdzb.so (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DzbError(
            f"{LIB_PATH} is missing: build it with `python -m discretize_b200._build` "
            "(there is no CPU fallback for the operator path)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:  # never let this masquerade as a missing python attribute
            raise DzbError(f"{LIB_PATH} does not export {name}; rebuild it") from exc
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if lib.dzb_version() != ABI_VERSION:
        raise DzbError(f"libdzb ABI {lib.dzb_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = "libdzb"):
    if rc != 0:
        msg = load().dzb_last_error()
        raise DzbError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
