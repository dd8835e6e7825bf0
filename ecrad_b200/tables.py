"""Packed read-only optics tables ("ETB1" blob).

One named-array directory holds everything `setup_radiation` leaves in Fortran module storage for the RRTMG
path (reference: radiation/radiation_interface.F90:37-156; module storage ifsrrtm/yoerrta*.F90,
yoesrta*.F90, yoerrtwn.F90, yoerrtrf.F90, yoesrtwn.F90; config%cloud_optics, config%pdf_sampler).
Arrays are stored in *Fortran element order* with their Fortran dimensions, so that a Fortran host can hand
over the same bytes by pointer (`ecrad_b200_tables_add`, include/ecrad_b200.h) and the C side indexes them
identically in both cases.

Layout (little endian):
    char[4]  magic "ETB1"
    uint32   n
    n x { char name[48]; int32 dtype (0 = float64, 1 = int32); int32 ndim; int64 dims[4]; int64 offset }
    data, each array 64-byte aligned, `offset` counted from the start of the file
"""
from __future__ import annotations

import struct

import numpy as np

_ENTRY = struct.Struct("<48sii4qq")
MAGIC = b"ETB1"


def write_blob(path, arrays):
    names = sorted(arrays)
    head = 8 + _ENTRY.size * len(names)
    off = (head + 63) // 64 * 64
    entries, chunks = [], []
    for nm in names:
        a = np.asarray(arrays[nm])
        if a.dtype.kind in "iu":
            a = a.astype("<i4")
            code = 1
        else:
            a = a.astype("<f8")
            code = 0
        if a.ndim == 0:
            a = a.reshape(1)
        dims = list(a.shape) + [1] * (4 - a.ndim)
        raw = a.ravel(order="F").tobytes()
        entries.append(_ENTRY.pack(nm.encode(), code, a.ndim, *dims, off))
        chunks.append((off, raw))
        off = (off + len(raw) + 63) // 64 * 64
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<I", len(names)))
        for e in entries:
            f.write(e)
        for o, raw in chunks:
            f.seek(o)
            f.write(raw)
        f.truncate(off)


def read_blob(path):
    """Return {name: ndarray} with the Fortran logical shape (element [i,j] == Fortran (i+1,j+1))."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:4] != MAGIC:
        raise ValueError(f"{path}: not an ETB1 table blob")
    (n,) = struct.unpack_from("<I", buf, 4)
    out = {}
    for i in range(n):
        nm, code, ndim, d0, d1, d2, d3, off = _ENTRY.unpack_from(buf, 8 + i * _ENTRY.size)
        dims = [d0, d1, d2, d3][:ndim]
        dt = "<f8" if code == 0 else "<i4"
        cnt = int(np.prod(dims))
        a = np.frombuffer(buf, dtype=dt, count=cnt, offset=off).reshape(dims, order="F")
        out[nm.rstrip(b"\0").decode()] = a
    return out
