"""Dataset wire formats of the reference (src/read/fvecs_read.jl, ivecs_read.jl, bvecs_read.jl):
every vector is stored as a little-endian int32 dimension header followed by d components
(float32 / int32 / uint8).  Readers return (n, d) arrays == the reference's d-by-n Julia matrices.
`bounds` mirrors the readers' `nvectors::Union{Integer,UnitRange}` argument: an int n -> the first n
vectors; a (first, last) pair -> that 1-based inclusive range (fvecs_read.jl:9-16)."""
import os

import numpy as np


def _read(path, comp_dtype, bounds):
    comp = np.dtype(comp_dtype)
    with open(path, "rb") as f:
        head = np.fromfile(f, dtype="<i4", count=1)
        if head.size == 0:
            return np.zeros((0, 0), comp)
        d = int(head[0])
    rec = 4 + d * comp.itemsize
    total = os.path.getsize(path) // rec
    if bounds is None:
        first, last = 1, total
    elif isinstance(bounds, (tuple, list, range)):
        first, last = (bounds[0], bounds[-1])
    else:
        first, last = 1, int(bounds)
    assert 1 <= first and last <= total and first <= last + 1, "requested range exceeds the file"
    n = last - first + 1
    raw = np.memmap(path, dtype=np.uint8, mode="r", offset=(first - 1) * rec, shape=(n, rec))
    dims = raw[:, :4].copy().view("<i4").reshape(-1)
    assert np.all(dims == d), "inconsistent dimension headers"
    return np.ascontiguousarray(raw[:, 4:]).view(comp.newbyteorder("<")).reshape(n, d).astype(comp, copy=False)


def fvecs_read(path, bounds=None):
    return _read(path, np.float32, bounds)


def ivecs_read(path, bounds=None):
    return _read(path, np.int32, bounds)


def bvecs_read(path, bounds=None):
    return _read(path, np.uint8, bounds)


def _write(path, a, comp_dtype):
    a = np.ascontiguousarray(a, dtype=np.dtype(comp_dtype).newbyteorder("<"))
    n, d = a.shape
    rec = np.empty((n, 4 + d * a.itemsize), np.uint8)
    rec[:, :4] = np.full(n, d, "<i4").view(np.uint8).reshape(n, 4)
    rec[:, 4:] = a.view(np.uint8).reshape(n, -1)
    rec.tofile(path)


def fvecs_write(path, a):
    _write(path, a, np.float32)


def ivecs_write(path, a):
    _write(path, a, np.int32)


def bvecs_write(path, a):
    _write(path, a, np.uint8)
