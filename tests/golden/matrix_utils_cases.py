"""discretize.utils matrix / index / coordinate helpers: one recipe for the golden generator (reference) and
tests/test_matrix_utils.py (discretize_b200).  ``evaluate(utils_module, lib)`` -> {name: array or marker string}."""
import numpy as np
import scipy.sparse as sp


def _dense(a):
    if sp.issparse(a):
        return np.asarray(a.toarray(), dtype=np.float64)
    if isinstance(a, tuple):
        return np.stack([np.asarray(x, dtype=np.float64) for x in a])
    return np.asarray(a)


def _guard(fn):
    try:
        out = fn()
    except Exception as e:
        return np.array("EXC:" + type(e).__name__)
    if out is None:
        return np.array("NONE")
    if type(out).__name__ in ("Zero", "Identity"):
        return np.array("SYM:" + repr(out))
    if isinstance(out, (bool, np.bool_)):
        return np.array(bool(out))
    return _dense(out)


def evaluate(u, lib):
    rng = np.random.default_rng(12)
    out = {}

    def put(name, fn):
        out[name] = _guard(fn)

    v = rng.random(5) + 0.5
    put("sdiag", lambda: u.sdiag(v))
    put("sdiag_zero", lambda: u.sdiag(u.Zero()))
    put("sdinv", lambda: u.sdinv(u.sdiag(v)))
    put("speye", lambda: u.speye(4))
    put("spzeros", lambda: u.spzeros(3, 5))
    put("kron3", lambda: u.kron3(u.ddx(2), u.av(3), u.speye(2)))
    for n in (1, 2, 5):
        put(f"ddx{n}", lambda: u.ddx(n))
        put(f"av{n}", lambda: u.av(n))
        put(f"av_extrap{n}", lambda: u.av_extrap(n))
    for shape in ((4,), (3, 4), (3, 2, 4)):
        for bdir in ("xyz", "x", "yz", "z"):
            put(f"boundary_bool_{len(shape)}_{bdir}", lambda: u.make_boundary_bool(shape, bdir))
    put("ind2sub", lambda: np.stack(u.ind2sub((7, 6), np.array([31, 41, 13]))))
    put("sub2ind", lambda: u.sub2ind((7, 6), np.array([[3, 4], [6, 5], [6, 1]])))
    put("sub2ind_1d", lambda: u.sub2ind((7,), np.array([3, 4])))
    put("sub2ind_bad", lambda: u.sub2ind((7, 6), np.array([[3, 4, 1]])))
    A = rng.random((3, 4, 5))
    put("get_subarray3", lambda: u.get_subarray(A, [np.array([0, 2]), np.array([1, 3]), np.array([0, 4])]))
    put("get_subarray2", lambda: u.get_subarray(A[0], [np.array([0, 2]), np.array([1, 3])]))
    put("get_subarray_bad", lambda: u.get_subarray(A, [np.array([0])]))
    put("get_subarray_type", lambda: u.get_subarray(A, np.array([0])))
    a = [rng.random(6) + (3.0 if i in (0, 4, 8) else 0.0) for i in range(9)]
    put("inv3", lambda: u.inverse_3x3_block_diagonal(*a, return_matrix=False))
    put("inv3_matrix", lambda: u.inverse_3x3_block_diagonal(*a))
    put("inv2", lambda: u.inverse_2x2_block_diagonal(a[0], a[1], a[3], a[4], return_matrix=False))
    put("inv2_matrix", lambda: u.inverse_2x2_block_diagonal(a[0], a[1], a[3], a[4]))
    B = rng.random((4, 2, 3, 3)) + 2 * np.eye(3)
    put("invert_blocks3", lambda: u.invert_blocks(B))
    put("invert_blocks2", lambda: u.invert_blocks(B[..., :2, :2]))
    put("invert_blocks_bad", lambda: u.invert_blocks(B[..., :2, :]))
    for d, h in ((2, [3, 2]), (3, [3, 2, 2])):
        mesh = lib.TensorMesh(h)
        nC = mesh.n_cells
        models = {"none": None, "scalar": 2.5, "iso": rng.random(nC) + .5, "aniso": rng.random((nC, d)) + .5,
                  "tensor": rng.random((nC, 3 if d == 2 else 6)) + np.r_[np.full(d, 2.0), np.zeros(3 if d == 3 else 1)],
                  "bad": np.ones(nC + 1)}
        for mk, model in models.items():
            put(f"{d}d_tensor_type_{mk}", lambda: str(u.TensorType(mesh, model)))
            put(f"{d}d_tensor_type_cmp_{mk}", lambda: np.array([u.TensorType(mesh, model) == 1, u.TensorType(mesh, model) < 3,
                                                               u.TensorType(mesh, model) >= 2]))
            put(f"{d}d_make_{mk}", lambda: u.make_property_tensor(mesh, model))
            if model is not None:
                put(f"{d}d_inverse_{mk}", lambda: u.inverse_property_tensor(mesh, model))
                put(f"{d}d_inverse_matrix_{mk}", lambda: u.inverse_property_tensor(mesh, model, return_matrix=True))
    put("cross2d", lambda: u.cross2d(rng.random((4, 2)), rng.random((4, 2))))
    x, y, z = np.cumsum(rng.random(5)), np.cumsum(rng.random(4)), np.cumsum(rng.random(6))
    pts = np.c_[rng.random(9) * 3, rng.random(9) * 2.5, rng.random(9) * 3.5]
    put("interp1", lambda: u.interpolation_matrix(pts[:, 0], x))
    put("interp2", lambda: u.interpolation_matrix(pts[:, :2], x, y))
    put("interp3", lambda: u.interpolation_matrix(pts, x, y, z))
    cyl = np.c_[rng.random(5) + .5, rng.random(5) * 6 - 3, rng.random(5)]
    vecs = rng.random((5, 3))
    put("cyl2cart", lambda: u.cylindrical_to_cartesian(cyl))
    put("cyl2cart_vec", lambda: u.cylindrical_to_cartesian(cyl, vecs))
    put("cyl2cart_vec_flat", lambda: u.cyl2cart(cyl, vecs.reshape(-1, order="F")))
    put("cart2cyl", lambda: u.cartesian_to_cylindrical(cyl))
    put("cart2cyl_vec", lambda: u.cart2cyl(cyl, vecs))
    v0, v1 = rng.random(3), rng.random(3)
    put("rotation", lambda: u.rotation_matrix_from_normals(v0, v1))
    put("rotation_same", lambda: u.rotation_matrix_from_normals(v0, 2 * v0))
    put("rotation_bad", lambda: u.rotation_matrix_from_normals(v0[:2], v1))
    put("rotate_points", lambda: u.rotate_points_from_normals(vecs, v0, v1, x0=np.r_[1.0, 2.0, 3.0]))
    # symbolic zero / identity
    Z, I, w = u.Zero(), u.Identity(), np.arange(4.0) + 1
    S = sp.identity(3, format="csr") * 2.0
    sym = {"Z+w": lambda: Z + w, "w+Z": lambda: w + Z, "Z-w": lambda: Z - w, "w-Z": lambda: w - Z, "Z*w": lambda: Z * w,
           "w*Z": lambda: w * Z, "Z@w": lambda: Z @ w, "Z/w": lambda: Z / w, "w/Z": lambda: w / Z, "-Z": lambda: -Z,
           "Z<1": lambda: Z < 1, "Z==0": lambda: Z == 0, "Z!=0": lambda: Z != 0, "Z.T": lambda: Z.T, "Z[0]": lambda: Z[0],
           "bool(Z)": lambda: bool(Z),
           "I+w": lambda: I + w, "w+I": lambda: w + I, "I-w": lambda: I - w, "w-I": lambda: w - I, "I*w": lambda: I * w,
           "w*I": lambda: w * I, "I@w": lambda: I @ w, "w@I": lambda: w @ I, "I/w": lambda: I / w, "w/I": lambda: w / I,
           "-I": lambda: -I, "(-I)*w": lambda: (-I) * w, "(-I)+w": lambda: (-I) + w, "I+S": lambda: I + S,
           "(-I)+S": lambda: (-I) + S, "I/S": lambda: I / S, "I==1": lambda: I == 1, "-I==-1": lambda: (-I) == -1,
           "I>0": lambda: I > 0, "I//w": lambda: I // w, "w//I": lambda: w // I, "I@Z": lambda: I @ Z, "Z@I": lambda: Z @ I,
           "bool(I)": lambda: bool(I), "I.T": lambda: I.T, "repr(-I)": lambda: repr(-I)}
    # (.shape of Zero / Identity is left out: the reference's helper tuple cannot be constructed and raises TypeError)
    for k, fn in sym.items():
        put("sym:" + k, fn)
    return {k: (np.array(str(v)) if isinstance(v, str) else v) for k, v in out.items()}
