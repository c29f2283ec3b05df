"""File boundary (include/prb_io.h): readers for .ply (ascii / LE / BE, properties by name),
.bnpts and ASCII records, and the mesh writer whose ASCII output must be byte-identical to the
reference's PlyWriteTriangles ("%g " per item, plyfile.cu:2136-2141, 2769-2837).  Host-only."""
import ctypes
import os

import numpy as np
import pytest

from poissonrecon_gpu_b200 import api, plyio


def _lib():
    lib = api.load_library()
    lib.prbio_read_points.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64)]
    lib.prbio_free.argtypes = [ctypes.c_void_p]
    lib.prbio_free.restype = None
    lib.prbio_write_mesh.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_float, ctypes.c_int]
    return lib


def read_points(path):
    lib = _lib()
    px, pn, n = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
    rc = lib.prbio_read_points(str(path).encode(), ctypes.byref(px), ctypes.byref(pn), ctypes.byref(n))
    if rc != 0:
        raise RuntimeError(lib.prb_last_error().decode())
    if n.value == 0:
        lib.prbio_free(px); lib.prbio_free(pn)
        return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32)
    a = np.ctypeslib.as_array(ctypes.cast(px, ctypes.POINTER(ctypes.c_float)), (n.value, 3)).copy()
    b = np.ctypeslib.as_array(ctypes.cast(pn, ctypes.POINTER(ctypes.c_float)), (n.value, 3)).copy()
    lib.prbio_free(px); lib.prbio_free(pn)
    return a, b


@pytest.fixture()
def cloud():
    g = np.random.default_rng(5)
    return g.standard_normal((1234, 3)).astype(np.float32), g.standard_normal((1234, 3)).astype(np.float32)


def test_binary_le_ply(cloud, tmp_path):
    p, n = cloud
    plyio.write_points_ply(str(tmp_path / "a.ply"), p, n, binary=True)
    a, b = read_points(tmp_path / "a.ply")
    assert np.array_equal(a, p) and np.array_equal(b, n)


def test_ascii_ply(cloud, tmp_path):
    p, n = cloud
    plyio.write_points_ply(str(tmp_path / "a.ply"), p, n, binary=False)   # %.9g round-trips float32
    a, b = read_points(tmp_path / "a.ply")
    assert np.array_equal(a, p) and np.array_equal(b, n)


def test_bnpts_and_ascii_records(cloud, tmp_path):
    p, n = cloud
    plyio.write_bnpts(str(tmp_path / "a.bnpts"), p, n)
    a, b = read_points(tmp_path / "a.bnpts")
    assert np.array_equal(a, p) and np.array_equal(b, n)
    with open(tmp_path / "a.xyz", "w") as fh:
        for i in range(p.shape[0]):
            fh.write(" ".join("%.9g" % v for v in list(p[i]) + list(n[i])) + "\n")
        fh.write("1 2 3\n")     # incomplete trailing record: ignored like fscanf != 6 (PointStream.inl:47)
    a, b = read_points(tmp_path / "a.xyz")
    assert np.array_equal(a, p) and np.array_equal(b, n)


def test_big_endian_mixed_types_any_order(cloud, tmp_path):
    """Properties are matched by NAME, any scalar type, any order, extras skipped
    (plyfile.cu:1008-1039); binary_big_endian is byte-swapped (plyfile.cu:782-792)."""
    p, n = cloud
    rec = np.dtype([("nz", ">f4"), ("q", ">u1"), ("x", ">f8"), ("ny", ">f4"), ("y", ">f4"), ("r", ">i2"), ("nx", ">f8"), ("z", ">f4")])
    a = np.zeros(p.shape[0], rec)
    a["x"], a["y"], a["z"] = p[:, 0], p[:, 1], p[:, 2]
    a["nx"], a["ny"], a["nz"] = n[:, 0], n[:, 1], n[:, 2]
    a["q"], a["r"] = 7, -3
    hdr = "ply\nformat binary_big_endian 1.0\ncomment made by test\nelement vertex %d\n" % p.shape[0]
    hdr += "property float nz\nproperty uchar q\nproperty double x\nproperty float ny\nproperty float y\nproperty short r\nproperty float64 nx\nproperty float32 z\n"
    hdr += "element face 0\nproperty list uchar int vertex_indices\nend_header\n"
    with open(tmp_path / "b.ply", "wb") as fh:
        fh.write(hdr.encode()); fh.write(a.tobytes())
    x, y = read_points(tmp_path / "b.ply")
    assert np.array_equal(x, p) and np.array_equal(y, n)


def test_reader_errors(tmp_path):
    with pytest.raises(RuntimeError, match="Failed to open"):
        read_points(tmp_path / "missing.ply")
    with open(tmp_path / "c.ply", "w") as fh:
        fh.write("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nproperty float y\nproperty float z\nend_header\n0 0 0\n")
    with pytest.raises(RuntimeError, match="Failed to find property in ply file: nx"):
        read_points(tmp_path / "c.ply")
    with open(tmp_path / "d.ply", "w") as fh:
        fh.write("ply\nformat ascii 1.0\nelement face 0\nproperty list uchar int vertex_indices\nelement vertex 0\nproperty float x\nend_header\n")
    with pytest.raises(RuntimeError, match="Could not find vertices"):
        read_points(tmp_path / "d.ply")
    with open(tmp_path / "e.ply", "wb") as fh:
        fh.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 10\nproperty float x\nproperty float y\nproperty float z\nproperty float nx\nproperty float ny\nproperty float nz\nend_header\n1234")
    with pytest.raises(RuntimeError, match="truncated"):
        read_points(tmp_path / "e.ply")
    open(tmp_path / "empty.bnpts", "wb").close()
    a, b = read_points(tmp_path / "empty.bnpts")
    assert a.shape == (0, 3)


def test_mesh_writer_matches_reference_format(tmp_path):
    lib = _lib()
    g = np.random.default_rng(3)
    v = g.uniform(0, 1, (70001, 3)).astype(np.float32)
    v[0] = [0, 1e-7, 123456.789]
    t = g.integers(0, v.shape[0], (90001, 3)).astype(np.int32)
    c = np.array([-1.25, 0.5, 3.0], np.float32)
    s = np.float32(2.5)
    out = tmp_path / "mesh"            # ".ply" is appended (plyfile.cu:251-257)
    assert lib.prbio_write_mesh(str(out).encode(), v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], c.ctypes.data, s, 0) == 0
    raw = open(str(out) + ".ply", "rb").read()
    hdr = ("ply\nformat ascii 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
           "element face %d\nproperty list uchar int vertex_indices\nend_header\n" % (v.shape[0], t.shape[0]))
    assert raw.startswith(hdr.encode())
    w = v * s + c                       # float32 arithmetic like plyfile.cu:2801-2803
    lines = raw[len(hdr):].decode().split("\n")
    assert lines[-1] == "" and len(lines) == v.shape[0] + t.shape[0] + 1
    for i in (0, 1, 2, 65535, 65536, 70000):
        assert lines[i] == "%g %g %g " % tuple(float(x) for x in w[i])
    for i in (0, 65535, 65536, 90000):
        assert lines[v.shape[0] + i] == "3 %d %d %d " % tuple(t[i])
    v2, t2 = plyio.read_mesh_ply(str(out) + ".ply")
    assert np.array_equal(t2, t) and np.allclose(v2, w, rtol=1e-5)
    # binary fast path
    assert lib.prbio_write_mesh(str(tmp_path / "m2.ply").encode(), v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], c.ctypes.data, s, 1) == 0
    v3, t3 = plyio.read_mesh_ply(str(tmp_path / "m2.ply"))
    assert np.array_equal(t3, t) and np.array_equal(v3, w)
