"""Mesh-construction utilities that sit in front of the operator path (reference: ``discretize/utils/mesh_utils.py``,
``matrix_utils.py``, ``code_utils.py``): building a tensor or tree mesh around a cloud of points, refining a tree around
points / a surface / a box, flagging the cells below a topography surface, cutting the core out of a padded tensor mesh,
plus the small array helpers user scripts import next to them.  Host-side; results identical to the reference's
(tests/test_mesh_utils.py against a fixture generated from it)."""
from __future__ import annotations

import warnings

import numpy as np

from .tensor_mesh import TensorMesh, as_array_n_by_dim, is_scalar, unpack_widths  # noqa: F401  (re-exported)


def mkvc(x, n_dims=1, **kwargs):
    """Flatten in column-major order to a vector with ``n_dims`` dimensions (matrix_utils.py:35-110)."""
    if isinstance(x, np.matrix):
        x = np.array(x)
    if hasattr(x, "tovec"):
        x = x.tovec()
    if not isinstance(x, np.ndarray):
        raise TypeError("Vector must be a numpy array")
    if n_dims == 1:
        return x.flatten(order="F")
    if n_dims == 2:
        return x.flatten(order="F")[:, np.newaxis]
    if n_dims == 3:
        return x.flatten(order="F")[:, np.newaxis, np.newaxis]
    raise ValueError("n_dims must be 1, 2 or 3")


def ndgrid(*args, vector=True, order="F"):
    """Gridded locations of 1-D vectors, first axis fastest (matrix_utils.py:307-400)."""
    if not isinstance(vector, bool):
        raise TypeError("'vector' must be a boolean")
    xin = args[0] if isinstance(args[0], list) else list(args)
    if len(xin) == 1:
        return xin[0]
    grids = np.meshgrid(*xin, indexing="ij")
    if vector:
        return np.stack([g.reshape(-1, order=order) for g in grids], axis=1)
    return tuple(grids)


def random_model(shape, random_seed=None, anisotropy=None, its=100, bounds=None, seed=None):
    """A smooth random model: uniform noise passed ``its`` times through a small averaging kernel (1-4-1 per axis by
    default), rescaled to ``bounds`` (mesh_utils.py:21-114).  Same generator and kernel as the reference, so a seed gives
    the same model."""
    import scipy.ndimage as ndi

    if bounds is None:
        bounds = [0, 1]
    if seed is not None:
        warnings.warn("Deprecated in version 0.11.0. The `seed` keyword argument has been renamed to `random_seed` for "
                      "consistency across the package. Please update your code to use the new keyword argument.",
                      FutureWarning, stacklevel=2)
        random_seed = seed
    rng = np.random.default_rng(random_seed)
    if random_seed is None:
        print("Using a seed of: ", rng.bit_generator.seed_seq)
    if isinstance(shape, (int, np.integer)):
        shape = (shape,)
    model = rng.random(shape)
    if anisotropy is None:
        if len(shape) == 1:
            kernel = np.array([1, 10.0, 1], dtype=float)
        elif len(shape) == 2:
            kernel = np.array([[1, 7, 1], [2, 10, 2], [1, 7, 1]], dtype=float)
        else:
            k = np.array([1.0, 4.0, 1.0])
            kernel = np.kron(np.kron(k.reshape(1, 3), k.reshape(3, 1)), k.reshape(1, 3)).reshape((3, 3, 3))
    else:
        if len(anisotropy.shape) != len(shape):
            raise ValueError("Anisotropy must be the same shape.")
        kernel = np.array(anisotropy, dtype=float)
    kernel = kernel / kernel.sum()
    for _ in range(its):
        model = ndi.convolve(model, kernel)
    model = (model - model.min()) / (model.max() - model.min())
    return model * (bounds[1] - bounds[0]) + bounds[0]


def closest_points_index(mesh, pts, grid_loc="CC", **kwargs):
    if "gridLoc" in kwargs:
        raise TypeError("The gridLoc keyword argument has been removed, please use grid_loc. "
                        "This message will be removed in discretize 1.0.0")
    warnings.warn("The closest_points_index utilty function has been moved to be a method of a class object. Please access "
                  "it as mesh.closest_points_index(). This will be removed in a future version of discretize",
                  DeprecationWarning, stacklevel=2)
    return mesh.closest_points_index(pts, grid_loc=grid_loc, discard=True)


def extract_core_mesh(xyzlim, mesh, mesh_type="tensor"):
    """(mask of the cells strictly inside the limits, TensorMesh of those cells) -- mesh_utils.py:246-393."""
    if not isinstance(mesh, TensorMesh):
        raise Exception("Only implemented for class TensorMesh")
    xyzlim = np.asarray(xyzlim, dtype=np.float64)
    if mesh.dim == 1:
        xyzlim = xyzlim.reshape(1, 2) if xyzlim.ndim == 1 else xyzlim
    if mesh.dim not in (1, 2, 3):
        raise Exception("Not implemented!")
    centers_1d = [mesh.cell_centers_x, mesh.cell_centers_y, mesh.cell_centers_z][:mesh.dim]
    cc = mesh.cell_centers.reshape(mesh.n_cells, -1)
    hs, origin = [], []
    active = np.ones(mesh.n_cells, dtype=bool)
    for d in range(mesh.dim):
        lo, hi = xyzlim[d, 0], xyzlim[d, 1]
        keep = (centers_1d[d] > lo) & (centers_1d[d] < hi)
        h = mesh.h[d][keep]
        hs.append(h)
        origin.append(centers_1d[d][keep][0] - h[0] * 0.5)
        active &= (cc[:, d] > lo) & (cc[:, d] < hi)
    if mesh.dim == 1:
        active = (mesh.cell_centers > xyzlim[0, 0]) & (mesh.cell_centers < xyzlim[0, 1])
    return active, TensorMesh(hs, origin=origin)


def mesh_builder_xyz(xyz, h, padding_distance=None, base_mesh=None, depth_core=None, expansion_factor=1.3,
                     mesh_type="tensor", tree_diagonal_balance=None):
    """Tensor or tree mesh around a cloud of points with minimum cell widths ``h`` (mesh_utils.py:396-560)."""
    from .tree_mesh import TreeMesh

    if mesh_type.lower() not in ["tensor", "tree"]:
        raise ValueError("Revise mesh_type. Only TENSOR | TREE mesh are implemented")
    if padding_distance is None:
        padding_distance = [[0, 0], [0, 0], [0, 0]]
    xyz = np.asarray(xyz)
    dim = xyz.shape[1]
    limits, center, n_core = [], [], []
    for d in range(dim):
        hi_lo = np.r_[xyz[:, d].max(), xyz[:, d].min()]
        limits.append(hi_lo)
        center.append(np.mean(hi_lo))
        n_core.append(int((hi_lo[0] - hi_lo[1]) / h[d]))
    if depth_core is not None:
        n_core[-1] += int(depth_core / h[-1])
        limits[-1][1] -= depth_core
    n_before = []
    if mesh_type.lower() == "tensor":
        def n_padding(dx, pad):      # cells growing by expansion_factor until they cover the padding distance
            length, nc = 0, 0
            while length < pad:
                nc += 1
                length = np.sum(dx * expansion_factor ** (np.asarray(range(nc)) + 1))
            return nc

        widths = []
        for d in range(dim):
            widths.append([(h[d], n_padding(h[d], padding_distance[d][0]), -expansion_factor), (h[d], n_core[d]),
                           (h[d], n_padding(h[d], padding_distance[d][1]), expansion_factor)])
            n_before.append(widths[-1][0][1])
        mesh = TensorMesh(widths)
    else:
        widths = []
        for d in range(dim):
            extent = limits[d][0] - limits[d][1] + np.sum(padding_distance[d])
            widths.append(np.ones(2 ** (int(np.log2(extent / h[d])) + 1)) * h[d])
        mesh = TreeMesh(widths, diagonal_balance=tree_diagonal_balance)
        for d in range(dim):
            core = limits[d][0] - limits[d][1]
            n_before.append(int(np.ceil((mesh.h[d].sum() - core) / h[d] / 2)))
    origin = [limits[d][1] - np.sum(mesh.h[d][: n_before[d]]) for d in range(dim)]
    mesh.origin = np.hstack(origin)
    if base_mesh is not None:          # shift onto the base mesh's cell-centre lattice, nearest to the centroid
        for d in range(base_mesh.dim):
            cc_base = getattr(base_mesh, "cell_centers_" + "xyz"[d])
            cc_local = getattr(mesh, "cell_centers_" + "xyz"[d])
            shift = (cc_base[np.max([np.searchsorted(cc_base, center[d]) - 1, 0])]
                     - cc_local[np.max([np.searchsorted(cc_local, center[d]) - 1, 0])])
            origin[d] += shift
            mesh.origin = np.hstack(origin)
    return mesh


def refine_tree_xyz(mesh, xyz, method="radial", octree_levels=(1, 1, 1), octree_levels_padding=None, finalize=False,
                    min_level=0, max_distance=np.inf):
    """Refine a TreeMesh around points ("radial"), below a surface ("surface") or inside their bounding box ("box");
    deprecated in the reference in favour of the TreeMesh methods but still what most scripts call
    (mesh_utils.py:562-935)."""
    if octree_levels_padding is not None:
        if len(octree_levels_padding) != len(octree_levels):
            raise ValueError("'octree_levels_padding' must be the length %i" % len(octree_levels))
    else:
        octree_levels_padding = np.zeros_like(octree_levels)
    octree_levels = np.asarray(octree_levels)
    octree_levels_padding = np.asarray(octree_levels_padding)
    xyz = np.asarray(xyz)
    kind = method.lower()
    if kind == "radial":
        warnings.warn("The radial option is deprecated as of `0.9.0` please update your code to use the "
                      "`TreeMesh.refine_points` functionality. It will be removed in a future version of discretize.",
                      DeprecationWarning, stacklevel=2)
        r_max = np.cumsum(mesh.h[0].min() * octree_levels * 2 ** np.arange(len(octree_levels)))
        for ii in range(len(octree_levels)):
            if r_max[ii] > 0:
                mesh.refine_ball(xyz, np.ones(xyz.shape[0]) * r_max[ii],
                                 np.ones(xyz.shape[0], dtype=np.int64) * (mesh.max_level - ii), finalize=False)
        if finalize:
            mesh.finalize()
    elif kind == "surface":
        from scipy import interpolate
        from scipy.spatial import Delaunay, cKDTree

        warnings.warn("The surface option is deprecated as of `0.9.0` please update your code to use the "
                      "`TreeMesh.refine_surface` functionality. It will be removed in a future version of discretize.",
                      DeprecationWarning, stacklevel=2)
        centroid = np.mean(xyz, axis=0)
        if mesh.dim == 2:
            r_out = np.abs(centroid[0] - xyz).max()
            hz = mesh.h[1].min()
        else:
            r_out = np.linalg.norm(np.r_[np.abs(centroid[0] - xyz[:, 0]).max(), np.abs(centroid[1] - xyz[:, 1]).max()])
            hz = mesh.h[2].min()
        z_max = np.cumsum(hz * octree_levels * 2 ** np.arange(len(octree_levels)))
        pad_width = np.cumsum(mesh.h[0].min() * octree_levels_padding * 2 ** np.arange(len(octree_levels_padding)))
        xy_pad = -1
        depth = z_max[-1]
        nearest = cKDTree(xyz)
        for ii in range(len(octree_levels) - 1, -1, -1):      # coarsest level first
            dx = mesh.h[0].min() * 2 ** ii
            dz = mesh.h[-1].min() * 2 ** ii
            if xy_pad != pad_width[ii]:                       # widen the surface horizontally for this level
                xy_pad = pad_width[ii]
                x_loc = (xyz - centroid) * ((r_out + xy_pad) / r_out) + centroid
                if mesh.dim == 3:
                    tri = Delaunay(x_loc[:, :2])
                    elevation = interpolate.LinearNDInterpolator(tri, x_loc[:, 2])
                else:
                    elevation = interpolate.interp1d(x_loc[:, 0], x_loc[:, 1], fill_value="extrapolate")
            lim_x = np.r_[x_loc[:, 0].max(), x_loc[:, 0].min()]
            n_x = int(np.ceil((lim_x[0] - lim_x[1]) / dx))
            if mesh.dim == 3:
                dy = mesh.h[1].min() * 2 ** ii
                lim_y = np.r_[x_loc[:, 1].max(), x_loc[:, 1].min()]
                n_y = int(np.ceil((lim_y[0] - lim_y[1]) / dy))
                gx, gy = np.meshgrid(np.linspace(lim_x[1], lim_x[0], n_x), np.linspace(lim_y[1], lim_y[0], n_y))
                xy = np.c_[gx.reshape(-1), gy.reshape(-1)]
                inside = tri.find_simplex(xy)
            else:
                xy = np.linspace(lim_x[1], lim_x[0], n_x)
                inside = np.ones_like(xy, dtype="bool")
            new_loc = np.c_[xy[inside != -1], elevation(xy[inside != -1])]
            r, _ = nearest.query(new_loc)
            z_offset = 0
            while z_offset < depth:                           # stack the sample points downward, one cell at a time
                keep = r < (max_distance + pad_width[ii])
                nnz = int(np.sum(keep))
                if nnz > 0:
                    mesh.insert_cells(np.c_[new_loc[keep, :mesh.dim - 1], new_loc[keep, -1] - z_offset],
                                      np.ones(nnz) * mesh.max_level - ii, finalize=False)
                z_offset += dz
            depth -= dz * octree_levels[ii]
        if finalize:
            mesh.finalize()
    elif kind == "box":
        warnings.warn("The box option is deprecated as of `0.9.0` please update your code to use the "
                      "`TreeMesh.refine_bounding_box` functionality. It will be removed in a future version of discretize.",
                      DeprecationWarning, stacklevel=2)
        bsw, tne = np.min(xyz, axis=0), np.max(xyz, axis=0)
        hs = np.asarray([h.min() for h in mesh.h])
        scale = 2 ** np.arange(len(octree_levels))
        pad = np.cumsum(hs[0] * octree_levels_padding * scale)
        if mesh.dim == 3:
            pad = np.c_[pad, np.cumsum(hs[1] * octree_levels_padding * scale)]
        pad = np.c_[pad, np.cumsum(hs[-1] * np.maximum(octree_levels - 1, 0) * scale)]
        levels, lows, highs = [], [], []
        for ii, n_cells in enumerate(octree_levels):
            if n_cells > 0:
                levels.append(mesh.max_level - ii)
                lows.append(bsw - pad[ii])
                highs.append(tne + pad[ii])
        mesh.refine_box(lows, highs, levels, finalize=finalize)
    else:
        raise NotImplementedError("Only method= 'radial', 'surface' or 'box' have been implemented")
    return mesh


def active_from_xyz(mesh, xyz, grid_reference="CC", method="linear"):
    """Boolean per cell: below the surface through ``xyz`` (cell centre for "CC", the whole cell top for "N");
    linear (or nearest) interpolation of the surface, nearest neighbour outside its hull (mesh_utils.py:938-1110)."""
    from scipy import interpolate
    from scipy.spatial import Delaunay, cKDTree

    try:
        if not mesh.is_symmetric:
            raise NotImplementedError("Unsymmetric CylindricalMesh is not yet supported")
    except AttributeError:
        pass
    if grid_reference not in ["N", "CC"]:
        raise ValueError("Value of grid_reference must be 'N' (nodal) or 'CC' (cell center)")
    xyz = np.asarray(xyz)
    dim = mesh.dim - 1
    if mesh.dim == 3:
        if xyz.shape[1] != 3:
            raise ValueError("xyz locations of shape (*, 3) required for 3D mesh")
        if method == "linear":
            surface = interpolate.LinearNDInterpolator(Delaunay(xyz[:, :2]), xyz[:, 2])
        else:
            surface = interpolate.NearestNDInterpolator(xyz[:, :2], xyz[:, 2])
    elif mesh.dim == 2:
        if xyz.shape[1] != 2:
            raise ValueError("xyz locations of shape (*, 2) required for 2D mesh")
        surface = interpolate.interp1d(xyz[:, 0], xyz[:, 1], bounds_error=False, fill_value=np.nan, kind=method)
    elif xyz.ndim != 1:
        raise ValueError("xyz locations of shape (*, ) required for 1D mesh")
    if mesh.dim == 1:
        grid = mesh.cell_centers_x if grid_reference == "CC" else mesh.nodes_x
        active = np.zeros(mesh.n_cells, dtype="bool")
        active[np.searchsorted(grid, xyz).max():] = True
        return active
    cc, hg = mesh.cell_centers, mesh.h_gridded
    if grid_reference == "CC":
        locations = cc
    elif mesh.dim == 3:      # the four top corners of every cell
        locations = np.vstack([cc + np.r_[sx, sy, 1.0] * hg / 2.0 for sx, sy in ((-1, 1), (-1, -1), (1, 1), (1, -1))])
    else:
        locations = np.vstack([cc + np.r_[sx, 1.0] * hg / 2.0 for sx in (-1, 1)])
    z_surface = np.asarray(surface(locations[:, :-1])).squeeze()
    missing = np.isnan(z_surface)
    if missing.any():
        _, ind = cKDTree(xyz).query(locations[missing, :])
        z_surface[missing] = xyz[ind, dim]
    return np.all((locations[:, dim] < z_surface).reshape((mesh.n_cells, -1), order="F"), axis=1).ravel()
