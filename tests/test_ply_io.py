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


def test_weld_merges_identical_positions_and_keeps_order():
    """prbio_weld_mesh (poisson_recon --weld): duplicate seam vertices collapse onto their first occurrence, triangles follow."""
    lib = _lib()
    lib.prbio_weld_mesh.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
    g = np.random.default_rng(9)
    base = g.random((500, 3)).astype(np.float32)
    dup = g.integers(0, 500, 300)
    v = np.concatenate([base, base[dup]]).astype(np.float32)          # 300 seam copies appended, like a second pass
    v[7] = [0.0, 0.5, 0.25]
    v = np.concatenate([v, np.array([[-0.0, 0.5, 0.25]], np.float32)])  # -0.0 welds with +0.0
    t = g.integers(0, v.shape[0], (2000, 3)).astype(np.int32)
    pos_before = v[t].copy()
    v2, t2 = v.copy(), t.copy()
    nv_out = ctypes.c_int64()
    assert lib.prbio_weld_mesh(v2.ctypes.data, v2.shape[0], t2.ctypes.data, t2.shape[0], ctypes.byref(nv_out)) == 0
    assert nv_out.value == 500
    assert np.array_equal(v2[:500], v[:500])                           # first occurrences, original order
    assert t2.max() < 500 and np.array_equal(v2[t2], pos_before)       # every triangle still has the same corner positions (+0 == -0)
    bad = np.array([[0, 1, 900]], np.int32)
    assert lib.prbio_weld_mesh(v2.ctypes.data, 500, bad.ctypes.data, 1, ctypes.byref(nv_out)) != 0


def test_truncated_or_hostile_headers_are_rejected(tmp_path):
    """Counts from the header are untrusted: absurd vertex counts and list lengths must fail cleanly, not overrun."""
    f = tmp_path / "huge.ply"
    f.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 4611686018427387904\nproperty float x\nproperty float y\nproperty float z\n"
                  b"property float nx\nproperty float ny\nproperty float nz\nend_header\n" + b"\0" * 48)
    with pytest.raises(RuntimeError):
        read_points(f)
    f = tmp_path / "list.ply"
    body = np.zeros(6, "<f4").tobytes() + np.array([-5], "<i4").tobytes() + b"\0" * 64
    f.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
                  b"property float nx\nproperty float ny\nproperty float nz\nproperty list int int junk\nend_header\n" + body)
    with pytest.raises(RuntimeError):
        read_points(f)
    f = tmp_path / "ascii.ply"
    f.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 1000000000\nproperty float x\nproperty float y\nproperty float z\n"
                  b"property float nx\nproperty float ny\nproperty float nz\nend_header\n1 2 3 4 5 6\n")
    with pytest.raises(RuntimeError):
        read_points(f)
