"""B-spline tables: (1) the oracle's restatement is byte-identical to the tables produced by the
REFERENCE's own host code (tests/golden/tables_d*.bin, written by oracle/_ref/ref_tables =
oracle/ref_tables_harness.cpp linked against the reference's FunctionData / PPolynomial sources);
(2) the product's compact translation-invariant tables (csrc/bspline_host.cpp, prb_host_tables)
carry exactly the same numbers."""
import ctypes
import os
import struct
import subprocess

import numpy as np
import pytest

from poissonrecon_gpu_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def parse(path):
    raw = open(path, "rb").read()
    D, res = struct.unpack_from("<ii", raw, 0)
    o = 8
    g = np.frombuffer(raw, "<f4", 16, o).reshape(4, 4); o += 64
    m = np.frombuffer(raw, "<f4", 16, o).reshape(4, 4); o += 64
    b = np.frombuffer(raw, "<f4", res * 20, o).reshape(res, 4, 5); o += res * 80
    full = None
    if D <= 6:
        full = np.frombuffer(raw, "<f8", 3 * res * res, o).reshape(3, res, res); o += 24 * res * res
    (npb,) = struct.unpack_from("<i", raw, o); o += 4
    pr = np.frombuffer(raw, np.dtype([("a", "<i4"), ("b", "<i4"), ("ff", "<f8"), ("df", "<f8"), ("d2", "<f8")]), npb, o)
    return dict(D=D, res=res, gauss=g, maxfn=m, base=b, full=full, probes=pr)


def host_table(depth, name, dtype):
    lib = api.load_library()
    lib.prb_host_tables.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64]
    lib.prb_host_tables.restype = ctypes.c_int64
    nb = lib.prb_host_tables(depth, name.encode(), None, 0)
    assert nb > 0, (name, nb)
    a = np.empty(nb // np.dtype(dtype).itemsize, dtype)
    assert lib.prb_host_tables(depth, name.encode(), a.ctypes.data, nb) == nb
    return a


@pytest.mark.parametrize("depth", [5, 8, 10])
def test_oracle_tables_match_reference_host_code(depth, tmp_path):
    out = tmp_path / "t.bin"
    subprocess.run([os.path.join(ROOT, "oracle", "build", "tables_dump"), str(depth), str(out)], check=True, stdout=subprocess.DEVNULL)
    assert open(out, "rb").read() == open(os.path.join(GOLD, f"tables_d{depth}.bin"), "rb").read()


def test_reference_kats():
    """Known-answer values extracted from the reference's FunctionData code."""
    g = parse(os.path.join(GOLD, "tables_d8.bin"))
    # the stored function is F/F(0) (FunctionData.inl:134-135, F(0) = 0.75); starts are unscaled
    kat = np.array([[1.125, 1.5, 0.5, -1.5], [-0.375, -1.5, -1.5, -0.5], [0.375, -1.5, 1.5, 0.5], [-1.125, 1.5, -0.5, 1.5]], np.float64)
    kat[:, :3] /= 0.75
    np.testing.assert_allclose(g["gauss"], kat, rtol=2e-7)
    b0 = g["base"][0]
    np.testing.assert_allclose(b0[0], [0.666666687, 1.33333325, 0.666666687, 0, -1], rtol=1e-7)
    np.testing.assert_allclose(b0[2], [2, -4, 2, 0, 1], rtol=1e-7)
    pr = g["probes"]
    same = pr[(pr["a"] == pr["b"])]
    for row in same:
        d = int(np.log2(row["a"] + 1))
        w = 2.0 ** -d
        assert abs(row["ff"] / w - 0.977781296) < 1e-6
        assert abs(row["d2"] * w - 1.7777853) < 1e-5
        assert abs(row["df"] + 1.33514404e-05) < 1e-9


@pytest.mark.parametrize("depth", [5, 8, 10])
def test_product_tables_match_reference_host_code(depth):
    g = parse(os.path.join(GOLD, f"tables_d{depth}.bin"))
    assert (host_table(depth, "gauss", "<f4").reshape(4, 4) == g["gauss"]).all()
    assert (host_table(depth, "max_depth_fn", "<f4").reshape(4, 4) == g["maxfn"]).all()
    assert (host_table(depth, "base_fn", "<f4").reshape(-1, 4, 5) == g["base"]).all()
    # same-depth values and the 27-point stencil (main.cu:1143-1158: double products, float store)
    pr = {(int(r["a"]), int(r["b"])): r for r in g["probes"]}
    ff0, ff1 = host_table(depth, "ff0", "<f8"), host_table(depth, "ff1", "<f8")
    d20, d21 = host_table(depth, "d20", "<f8"), host_table(depth, "d21", "<f8")
    st = host_table(depth, "stencil", "<f4").reshape(depth + 1, 27)
    for d in range(depth + 1):
        n = 1 << d
        a = (n - 1) + n // 2
        assert ff0[d] == pr[(a, a)]["ff"] and d20[d] == pr[(a, a)]["d2"]
        if d >= 1:
            b = a - 1
            assert ff1[d] == pr[(a, b)]["ff"] and d21[d] == pr[(a, b)]["d2"]
        if d >= 2:
            for j in range(27):
                t = [j // 9 - 1, (j // 3) % 3 - 1, j % 3 - 1]
                f = [ff0[d] if x == 0 else ff1[d] for x in t]
                s = [d20[d] if x == 0 else d21[d] for x in t]
                want = np.float32(s[0] * f[1] * f[2] + f[0] * s[1] * f[2] + f[0] * f[1] * s[2])
                assert st[d, j] == want, (d, j)
    # divergence rows: dfT[d][t] = <dF_o, F_s>, t = off_s - k*(off_o - 1)
    off = host_table(depth, "df_offset", "<i4")
    dft = host_table(depth, "df_table", "<f4")
    nD = 1 << depth
    checked = 0
    for r in g["probes"]:
        a, b = int(r["a"]), int(r["b"])
        if b < nD - 1 or r["df"] == 0:
            continue
        d = int(np.log2(a + 1))
        k = 1 << (depth - d)
        t = (b - (nD - 1)) - k * ((a - ((1 << d) - 1)) - 1)
        if 0 <= t < 3 * k:
            assert dft[off[d] + t] == np.float32(r["df"]), (d, t)
            checked += 1
    assert checked > 3 * depth


def test_full_tables_translation_invariance():
    """D=5 golden holds the full res x res tables: the compact rows reproduce every entry the
    divergence reads -- slots s under the 27 neighbours of o, i.e. t = off_s - k*(off_o-1) in [0,3k)
    (main.cu:1023-1056 only visits those)."""
    depth = 5
    g = parse(os.path.join(GOLD, "tables_d5.bin"))
    full = g["full"]
    off = host_table(depth, "df_offset", "<i4")
    dft = host_table(depth, "df_table", "<f4")
    nD = 1 << depth
    for d in range(depth + 1):
        k = 1 << (depth - d)
        for oo in range(1 << d):
            a = (1 << d) - 1 + oo
            for s in range(nD):
                b = nD - 1 + s
                want = np.float32(full[1, b, a])   # flat index a + res*b
                t = s - k * (oo - 1)
                if 0 <= t < 3 * k:
                    assert dft[off[d] + t] == want, (d, oo, s)


def test_cross_depth_integrals_against_quadrature():
    """Tables of the opt-in cascadic mode (prb_host_tables ff_cross / d2_cross; not in the reference): <F_o, F_n> and <F_o', F_n'> for a
    depth-d node o and a coarser depth-e node n, against Gauss-Legendre quadrature of the closed-form B-spline between the knots."""
    import ctypes
    from poissonrecon_gpu_b200 import api
    lib = api.load_library()

    def table(D, name, dt):
        nb = lib.prb_host_tables(D, name.encode(), None, 0)
        a = np.empty(nb // np.dtype(dt).itemsize, dt)
        lib.prb_host_tables(D, name.encode(), a.ctypes.data, nb)
        return a

    D = 6
    ff, d2, off = table(D, "ff_cross", "<f8"), table(D, "d2_cross", "<f8"), table(D, "cross_offset", "<i4").reshape(D + 1, D + 1)
    assert ff.size == d2.size == sum(3 << (d - e) for d in range(D + 1) for e in range(d))

    def B(t):
        t = np.abs(t)
        return np.where(t < 0.5, 1 - (4 / 3) * t * t, np.where(t < 1.5, (2 / 3) * (1.5 - t) ** 2, 0.0))

    def dB(t):
        a = np.abs(t)
        return np.where(a < 0.5, -(8 / 3) * t, np.where(a < 1.5, -(4 / 3) * (1.5 - a) * np.sign(t), 0.0))

    xs, ws = np.polynomial.legendre.leggauss(8)

    def integ(f, knots):
        pts = np.unique(knots)
        return sum(0.5 * (b - a) * (ws * f(0.5 * (b - a) * xs + 0.5 * (a + b))).sum() for a, b in zip(pts[:-1], pts[1:]))

    worst = 0.0
    for d in range(1, D + 1):
        for e in range(d):
            k, w1, w2 = 1 << (d - e), 2.0 ** -d, 2.0 ** -e
            for u in range(3 * k):
                c1, c2 = (u + 0.5) * w1, 1.5 * w2                     # o at fine offset u, n at coarse offset 1: u = off_o - k (off_n - 1)
                knots = np.concatenate([c1 + w1 * np.array([-1.5, -.5, .5, 1.5]), c2 + w2 * np.array([-1.5, -.5, .5, 1.5])])
                F = integ(lambda x: B((x - c1) / w1) * B((x - c2) / w2), knots)
                G = integ(lambda x: dB((x - c1) / w1) / w1 * dB((x - c2) / w2) / w2, knots)
                worst = max(worst, abs(ff[off[d, e] + u] - F) / w1, abs(d2[off[d, e] + u] - G) * w1)
    assert worst < 5e-5, worst          # the table code does its polynomial algebra in float, like the reference's (DF(0) = -1.3e-5)
