"""Compact storage for golden dictionaries: mostly-zero 2-D float arrays are stored as COO triplets, everything else as
is.  ``save(path, dict)`` / ``load(path)``; the loaded object offers ``.files`` and ``[key]`` like ``numpy.load`` and
returns the original dense arrays."""
import numpy as np

_SEP = "::"


def save(path, arrays):
    out = {}
    for k, v in arrays.items():
        v = np.asarray(v)
        if v.ndim == 2 and v.dtype.kind == "f" and v.size > 400 and np.count_nonzero(v) * 3 < v.size:
            r, c = np.nonzero(v)
            out[k + _SEP + "rc"] = np.stack([r, c]).astype(np.int32)
            out[k + _SEP + "val"] = v[r, c]
            out[k + _SEP + "shape"] = np.array(v.shape, dtype=np.int64)
        else:
            out[k] = v
    np.savez_compressed(path, **out)


class load:
    def __init__(self, path):
        self._z = np.load(path)
        names = set()
        for k in self._z.files:
            names.add(k.split(_SEP)[0] if _SEP in k and k.rsplit(_SEP, 1)[1] in ("rc", "val", "shape") else k)
        self.files = sorted(names)

    def __getitem__(self, k):
        if k in self._z.files:
            return self._z[k]
        shape = tuple(self._z[k + _SEP + "shape"])
        out = np.zeros(shape)
        rc = self._z[k + _SEP + "rc"]
        out[rc[0], rc[1]] = self._z[k + _SEP + "val"]
        return out
