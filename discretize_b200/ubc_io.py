"""UBC-GIF mesh and model files for TensorMesh (2-D, 3-D) and TreeMesh (QuadTree / OcTree).

Reference: ``discretize/mixins/mesh_io.py`` (TensorMeshIO :15-420, TreeMeshIO :440-615) and the UBC cell ordering of
``_extensions/tree_ext.pyx:6536-6565``.  The format conventions that matter:

* the vertical axis of the files points DOWN from the top of the mesh: mesh files give the top south-west corner and list
  the vertical cell widths top first; 3-D tensor models are stored z fastest (top first), then x, then y;
* 2-D tensor mesh files describe each axis as a run-length list of node positions ``x0 x1 n`` / ``x n``;
* OcTree files list every cell as the 1-based index of its top south-west fine cell plus its width in fine cells, sorted by
  x fastest, then y, then z from the top; models follow that order.

Files written here are read back by the reference and vice versa (tests/test_ubc_io.py, fixtures written by the reference).
"""
from __future__ import annotations

import os

import numpy as np


def _path(file_name, directory):
    return os.path.join(directory or "", file_name)


def _lines(path):
    """Non-empty lines with everything after a ``!`` removed."""
    out = []
    with open(path, "r") as f:
        for line in f:
            line = line.split("!")[0].strip()
            if line:
                out.append(line)
    return out


# ---- TensorMesh --------------------------------------------------------------------------------------------------------


def _expand_widths(line):
    """``"3*10.0 5.0"`` -> [10, 10, 10, 5]."""
    out = []
    for token in line.split():
        if "*" in token:
            count, width = token.split("*")
            out.extend([float(width)] * int(count))
        else:
            out.append(float(token))
    return np.array(out, dtype=np.float64)


def _read_segments(lines, pos):
    """One axis of a 2-D mesh file starting at ``lines[pos]`` (the number of segments): (first node, widths, next pos)."""
    nseg = int(lines[pos].split()[0])
    widths = []
    x0 = end = None
    for k in range(nseg):
        vals = lines[pos + 1 + k].split()
        if k == 0:
            x0, x1, n = float(vals[0]), float(vals[1]), int(float(vals[2]))
            widths.append(np.full(n, (x1 - x0) / n))
            end = x1
        else:
            x1, n = float(vals[0]), int(float(vals[1]))
            widths.append(np.full(n, (x1 - end) / n))
            end = x1
    return x0, np.concatenate(widths), pos + 1 + nseg


def read_tensor_mesh(cls, file_name, directory=None):
    path = _path(file_name, directory)
    lines = _lines(path)
    n_first = len(lines[0].split())
    if n_first == 3:
        origin = np.array(lines[1].split(), dtype=np.float64)
        hx, hy, hz_top_down = (_expand_widths(lines[k]) for k in (2, 3, 4))
        hz = hz_top_down[::-1]
        origin[2] -= hz.sum()                      # the file gives the TOP south-west corner
        return cls([hx, hy, hz], origin=origin)
    if n_first == 1:
        x0, dx, pos = _read_segments(lines, 0)
        z0, dz, _ = _read_segments(lines, pos)
        return cls([dx, dz[::-1]], origin=(x0, -(z0 + dz.sum())))      # depths are positive down
    raise Exception("File format not recognized")


def _node_runs(nodes):
    """Run-length description of a node vector: maximal runs of (bitwise) equal spacing."""
    out = []
    i = 0
    while i + 1 < nodes.size:
        step = nodes[i + 1] - nodes[i]
        j = i + 1
        while j + 1 < nodes.size and nodes[j + 1] - nodes[j] == step:
            j += 1
        out.append((nodes[j], j - i))
        i = j
    text = f"{len(out):d}\n"
    for k, (end, n) in enumerate(out):
        text += f"{nodes[0]:.10f} {end:.10f} {n:d} \n" if k == 0 else f"{end:.10f} {n:d} \n"
    return text


def write_tensor_mesh(mesh, file_name, models=None, directory=None, comment_lines=""):
    path = _path(file_name, directory)
    if mesh.dim == 3:
        top = np.asarray(mesh.origin, dtype=np.float64) + np.array([0.0, 0.0, mesh.h[2].sum()])
        text = comment_lines + "{0:d} {1:d} {2:d}\n".format(*mesh.shape_cells)
        text += "{0:.6f} {1:.6f} {2:.6f}\n".format(*top)
        for widths in (mesh.h[0], mesh.h[1], mesh.h[2][::-1]):
            text += "".join("%.6f " % w for w in widths) + "\n"
    elif mesh.dim == 2:
        text = comment_lines + _node_runs(mesh.nodes_x) + "\n" + _node_runs(-mesh.nodes_y[::-1])
    else:
        raise Exception("mesh must be a Tensor Mesh 2D or 3D")
    with open(path, "w") as f:
        f.write(text)
    if models is None:
        return
    if not isinstance(models, dict):
        raise TypeError("models must be a dict")
    for key, model in models.items():
        if not isinstance(key, str):
            raise TypeError("The dict key must be a string representing the file name")
        write_tensor_model(mesh, key, model, directory=directory)


def read_tensor_model(mesh, file_name, directory=None):
    path = _path(file_name, directory)
    if mesh.dim == 3:
        values = np.loadtxt(path, dtype=np.float64).reshape(-1)
        nx, ny, nz = mesh.shape_cells
        cube = values.reshape((nz, nx, ny), order="F")[::-1]            # file: z (top first) fastest, then x, then y
        return cube.transpose(1, 2, 0).reshape(-1, order="F")
    if mesh.dim == 2:
        lines = _lines(path)
        if tuple(int(v) for v in lines[0].split()) != tuple(mesh.shape_cells):
            raise Exception("Dimension of the model and mesh mismatch")
        values = np.array([float(v) for line in lines[1:] for v in line.split()])
        if values.size != mesh.n_cells:
            raise Exception(f"Something is not right, expected size is {mesh.n_cells:d} but unwrap vector is size {values.size:d}")
        return values.reshape(mesh.shape_cells, order="F")[:, ::-1].reshape(-1, order="F")
    raise Exception("mesh must be a Tensor Mesh 2D or 3D")


def write_tensor_model(mesh, file_name, model, directory=None):
    path = _path(file_name, directory)
    model = np.asarray(model)
    if mesh.dim == 3:
        cube = model.reshape(mesh.shape_cells, order="F").transpose(2, 0, 1)[::-1]
        np.savetxt(path, cube.reshape(-1, order="F"))
    elif mesh.dim == 2:
        with open(path, "w") as f:
            f.write("{:d} {:d}\n".format(*mesh.shape_cells))
        with open(path, "ab") as f:
            np.savetxt(f, model.reshape(mesh.shape_cells, order="F").T[::-1])
    else:
        raise Exception("mesh must be a Tensor Mesh 2D or 3D")


# ---- TreeMesh ----------------------------------------------------------------------------------------------------------------


def tree_ubc_cells(mesh):
    """(1-based top south-west fine index per axis, width in fine cells) of every cell, in the mesh's cell order."""
    centers, levels = mesh.__getstate__()
    centers = np.array(centers, dtype=np.int64)
    width = np.left_shift(np.int64(1), (mesh.max_level - np.asarray(levels)).astype(np.int64))
    n_last = mesh._xs[mesh.dim - 1].size - 1
    centers[:, -1] = n_last - centers[:, -1]                  # count the last axis from the top
    return (centers - width[:, None]) // 2 + 1, width


def tree_ubc_order(mesh):
    """Permutation from the mesh's cell order to the file order: x fastest, then y, (then z from the top)."""
    key = "ubc_order"
    if key not in mesh._cache:
        idx, _ = tree_ubc_cells(mesh)
        mesh._cache[key] = np.lexsort(tuple(idx[:, d] for d in range(mesh.dim)))
    return mesh._cache[key]


def read_tree_mesh(cls, file_name, directory=None):
    lines = _lines(_path(file_name, directory))
    n_base = np.array(lines[0].split(), dtype=np.int64)
    top_corner = np.array(lines[1].split(), dtype=np.float64)
    small = np.array(lines[2].split(), dtype=np.float64)
    table = np.array([[int(v) for v in line.split()] for line in lines[4:]], dtype=np.int64)
    n_base = n_base[: top_corner.size]
    hs = [np.ones(n) * w for n, w in zip(n_base, small)]
    origin = top_corner.copy()
    origin[-1] -= hs[-1].sum()
    ls = np.log2(n_base).astype(int)
    max_level = ls[0] if ls.min() == ls.max() else ls.min() + 1
    mesh = cls(hs, origin=origin, diagonal_balance=False)
    width = table[:, -1]
    centers = 2 * (table[:, :-1] - 1) + width[:, None]
    centers[:, -1] = 2 * n_base[-1] - centers[:, -1]
    mesh.__setstate__((centers, (max_level - np.log2(width)).astype(np.int64)))
    return mesh


def write_tree_mesh(mesh, file_name, models=None, directory=None):
    if not all(np.allclose(h, h[0]) for h in mesh.h):
        raise Exception("UBC form does not support variable cell widths")
    top_corner = np.array(mesh.origin, dtype=np.float64)
    top_corner[-1] += mesh.h[-1].sum()
    idx, width = tree_ubc_cells(mesh)
    order = tree_ubc_order(mesh)
    head = " ".join(f"{int(h.size)}" for h in mesh.h) + " \n"
    head += " ".join(f"{v:.4f}" for v in top_corner) + " \n"
    head += " ".join(f"{h[0]:.3f}" for h in mesh.h) + " \n"
    head += f"{int(mesh.n_cells)}"
    np.savetxt(_path(file_name, directory), np.c_[idx[order], width[order]], fmt="%i", header=head, comments="")
    if models is None:
        return
    if not isinstance(models, dict):
        raise TypeError("models must be a dict")
    for key, model in models.items():
        write_tree_model(mesh, key, model, directory=directory)


def read_tree_model(mesh, file_name, directory=None):
    if isinstance(file_name, list):
        return {f: read_tree_model(mesh, f, directory=directory) for f in file_name}
    values = np.loadtxt(file_name if not directory else _path(file_name, directory))
    order = tree_ubc_order(mesh)
    back = np.empty_like(order)
    back[order] = np.arange(order.size)
    return values[back].copy()


def write_tree_model(mesh, file_name, model, directory=None):
    if isinstance(file_name, list):
        for f, m in zip(file_name, model):
            write_tree_model(mesh, f, m)
        return
    np.savetxt(_path(file_name, directory), np.asarray(model)[tree_ubc_order(mesh)])
