"""Size-independent structural properties of the octree arrays and the mesh:
used on the oracle (CPU) and, at BASELINE.json's full sizes where the oracle would take minutes,
on the CUDA path alone."""
import numpy as np


def check_octree(a, n_points, D):
    base, cnt = a["base"].astype(np.int64), a["count"].astype(np.int64)
    M = int(base[D + 1])
    assert cnt[0] == 1 and base[0] == 0
    assert np.array_equal(np.cumsum(cnt), base[1:D + 2])
    key, parent, pnum, pidx = a["key"].astype(np.int64), a["parent"], a["pnum"], a["pidx"]
    dnum, didx = a["dnum"], a["didx"]
    ch = a["children"].reshape(M, 8)
    nb = a["neighs"].reshape(M, 27)
    sk = a["sorted_key"].astype(np.int64)
    assert (np.diff(sk) >= 0).all(), "sample keys are not sorted"
    assert parent[0] == -1 and pnum[0] == n_points and dnum[0] == cnt[D]
    for d in range(1, D + 1):
        sl = slice(int(base[d]), int(base[d + 1]))
        k = key[sl]
        assert (np.diff(k) > 0).all(), f"node keys of depth {d} are not strictly ascending"
        # sibling groups: 8 consecutive slots share the parent, slot = child code
        shift = 3 * (D - d)
        assert np.array_equal((k >> shift) & 7, np.tile(np.arange(8), cnt[d] // 8))
        par = parent[sl]
        assert (par.reshape(-1, 8) == par.reshape(-1, 8)[:, :1]).all()
        assert (par >= base[d - 1]).all() and (par < base[d]).all()
        # M_d = 8 * (number of non-empty nodes of depth d-1) and children link back
        prev = slice(int(base[d - 1]), int(base[d]))
        nonempty_prev = pnum[prev] > 0
        assert cnt[d] == 8 * int(nonempty_prev.sum())
        c0 = ch[prev][:, 0]
        assert ((c0 >= 0) == nonempty_prev).all()
        assert np.array_equal(parent[c0[nonempty_prev]], np.arange(int(base[d - 1]), int(base[d]))[nonempty_prev])
        assert pnum[sl].sum() == n_points
        # sample ranges are contiguous prefix sums in node order
        assert np.array_equal(pidx[sl], np.concatenate([[0], np.cumsum(pnum[sl])[:-1]]))
        assert dnum[sl].sum() == cnt[D]
        assert np.array_equal(didx[sl], np.concatenate([[0], np.cumsum(dnum[sl])[:-1]]))
    # neighbour tables: self at slot 13, symmetric, same depth
    assert np.array_equal(nb[:, 13], np.arange(M))
    depth_of = np.repeat(np.arange(D + 1), cnt[:D + 1])
    for j in range(27):
        q = nb[:, j]
        ok = q >= 0
        assert np.array_equal(nb[q[ok], 26 - j], np.nonzero(ok)[0]), f"neighbour slot {j} is not symmetric"
        assert np.array_equal(depth_of[q[ok]], depth_of[ok])
    # every sample maps to a non-empty depth-D leaf that contains it
    p2n = a["p2n"].astype(np.int64)
    leaf = base[D] + p2n
    assert (pnum[leaf] > 0).all()
    i = np.arange(n_points)
    assert ((pidx[leaf] <= i) & (i < pidx[leaf] + pnum[leaf])).all()
    assert np.array_equal(key[leaf], sk)


def check_mesh(v, t, passes, finite=True):
    nv, nt = v.shape[0], t.shape[0]
    assert passes[:, 1].sum() == nv and passes[:, 2].sum() == nt
    if nt:
        assert t.min() >= 0 and t.max() < nv
    if nv and finite:
        assert np.isfinite(v).all() and v.min() >= 0.0 and v.max() <= 1.0
    # triangles of a pass only use that pass's vertices (insertTriangle offsets, main.cu:3224-3243)
    av = at = 0
    for kind, pv, pt in passes:
        if pt:
            tri = t[at:at + pt]
            assert tri.min() >= av and tri.max() < av + pv
            # a marching-cubes triangle never repeats a vertex
            assert ((tri[:, 0] != tri[:, 1]) & (tri[:, 1] != tri[:, 2]) & (tri[:, 0] != tri[:, 2])).all()
        av += pv
        at += pt
