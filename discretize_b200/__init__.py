"""discretize_b200: B200-native operator application for simpeg/discretize meshes.

Keeps the reference's property and method names (``face_divergence``, ``edge_curl``,
``nodal_gradient``, ``average_*_to_cell``, ``get_face_inner_product``,
``get_edge_inner_product``, their ``*_deriv`` and ``get_interpolation_matrix``) on
``TensorMesh`` (and ``TreeMesh``); every operator is a
``scipy.sparse.linalg.LinearOperator``-compatible object applied by hand-written
sm_100a CUDA kernels behind the C ABI in ``include/dzb.h``.
"""
from .tensor_mesh import TensorMesh
from .tree_mesh import TreeMesh, TreeMeshNotFinalizedError
from .operators import Operator
from ._lib import DzbError
from . import utils

__all__ = ["TensorMesh", "TreeMesh", "TreeMeshNotFinalizedError", "Operator", "DzbError", "utils"]
__version__ = "0.1.0"
