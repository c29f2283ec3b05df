"""Minimal PLY / bnpts readers and writers for tests and benches (numpy, host side).

Formats follow the reference's file contract: oriented points as a PLY
``vertex`` element with float ``x y z nx ny nz`` (ascii / binary_little_endian), or ``.bnpts``
(raw float32 x6 per point, PointStream.inl:72-92); meshes as ASCII PLY with ``%g`` floats and
``list uchar int vertex_indices`` faces (plyfile.cu:2769-2837).  The production readers /
writers are the C++ ones behind the CLI (csrc/ply_io.cpp); these exist so the Python tests can
create inputs and parse outputs independently of that code.
"""
from __future__ import annotations

import numpy as np


def write_points_ply(path: str, pts: np.ndarray, nrm: np.ndarray, binary: bool = True) -> None:
    n = pts.shape[0]
    hdr = "ply\nformat %s 1.0\nelement vertex %d\n" % ("binary_little_endian" if binary else "ascii", n)
    for name in ("x", "y", "z", "nx", "ny", "nz"):
        hdr += "property float %s\n" % name
    hdr += "end_header\n"
    data = np.concatenate([pts.astype("<f4"), nrm.astype("<f4")], axis=1)
    with open(path, "wb") as fh:
        fh.write(hdr.encode())
        if binary:
            fh.write(np.ascontiguousarray(data).tobytes())
        else:
            for row in data:
                fh.write((" ".join("%.9g" % v for v in row) + "\n").encode())


def write_bnpts(path: str, pts: np.ndarray, nrm: np.ndarray) -> None:
    np.concatenate([pts.astype("<f4"), nrm.astype("<f4")], axis=1).tofile(path)


def read_mesh_ply(path: str):
    """Parse a triangle mesh PLY (ascii or binary_little_endian, float xyz + uchar/int list)."""
    with open(path, "rb") as fh:
        raw = fh.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode().splitlines()
    fmt = [l for l in header if l.startswith("format")][0].split()[1]
    nv = nf = 0
    for l in header:
        t = l.split()
        if t[:2] == ["element", "vertex"]:
            nv = int(t[2])
        if t[:2] == ["element", "face"]:
            nf = int(t[2])
    body = raw[end:]
    if fmt == "ascii":
        tok = body.split()
        v = np.array(tok[: 3 * nv], dtype=np.float64).reshape(nv, 3).astype(np.float32)
        f = np.array(tok[3 * nv: 3 * nv + 4 * nf], dtype=np.int64).reshape(nf, 4)
        assert nf == 0 or (f[:, 0] == 3).all()
        return v, f[:, 1:].astype(np.int32)
    v = np.frombuffer(body, dtype="<f4", count=3 * nv).reshape(nv, 3)
    rec = np.dtype([("n", "u1"), ("idx", "<i4", (3,))])
    f = np.frombuffer(body, dtype=rec, count=nf, offset=12 * nv)
    return v.copy(), f["idx"].copy()
