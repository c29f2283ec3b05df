"""Seeded input cases shared by the oracle property tests (CPU) and the parity tests (GPU)."""
import numpy as np

from poissonrecon_gpu_b200 import synth


def _unit(v):
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def make_case(name):
    g = np.random.default_rng(abs(hash(name)) % (2**31) if False else sum(map(ord, name)))
    if name == "sphere20k_d6":
        p, n = synth.sphere(20_000, seed=7)
        return p, n, 6
    if name == "sphere3k_d5":
        p, n = synth.sphere(3_000, seed=8)
        return p, n, 5
    if name == "torus60k_d7":          # 32 empty depth-2 leaves -> coarse single-root passes are exercised
        p, n = synth.torus(60_000)
        return p, n, 7
    if name == "scan80k_d7":           # 20:1 density skew + radial noise
        p, n = synth.nonuniform_scan(80_000)
        return p, n, 7
    if name == "multi120k_d7":
        p, n = synth.multi_object(120_000)
        return p, n, 7
    if name == "sphere8k_d8":          # sparse deep tree: long empty-sibling runs, large refinement passes
        p, n = synth.sphere(8_000, seed=9)
        return p, n, 8
    if name == "one_point_d5":
        return np.array([[0.3, -0.2, 0.9]], np.float32), np.array([[0, 0, 1]], np.float32), 5
    if name == "two_points_d4":
        return np.array([[0, 0, 0], [1, 2, 3]], np.float32), np.array([[0, 0, 1], [1, 0, 0]], np.float32), 4
    if name == "duplicates_d6":        # many coincident samples: one leaf with a large pnum
        p, n = synth.sphere(500, seed=3)
        p = np.concatenate([p, np.repeat(p[:5], 200, axis=0)])
        n = np.concatenate([n, np.repeat(n[:5], 200, axis=0)])
        return p, n, 6
    if name == "lattice_d5":           # samples exactly on cell boundaries after normalisation (Q5: strict >)
        k = np.arange(17, dtype=np.float32)
        x, y = np.meshgrid(k, k, indexing="ij")
        p = np.stack([x.ravel(), y.ravel(), np.zeros(x.size, np.float32)], 1)
        p = np.concatenate([p, p + np.array([0, 0, 16], np.float32)])
        n = np.concatenate([np.tile([0, 0, -1], (x.size, 1)), np.tile([0, 0, 1], (x.size, 1))]).astype(np.float32)
        return p.astype(np.float32), n, 5
    if name == "zero_normals_d5":      # |n| <= 1e-6 is not normalised (main.cu:561-568)
        p, n = synth.sphere(2_000, seed=4)
        n = n.copy()
        n[::7] = 0
        n[1::7] *= 1e-3
        return p, n, 5
    if name == "cluster_plus_outlier_d8":
        p = (g.standard_normal((4000, 3)) * 1e-3).astype(np.float32)
        p = np.concatenate([p, np.array([[5, 5, 5]], np.float32)])
        return p, _unit(g.standard_normal((4001, 3))), 8
    if name == "sphere2k_d2":
        p, n = synth.sphere(2_000, seed=5)
        return p, n, 2
    if name == "sphere2k_d3":
        p, n = synth.sphere(2_000, seed=6)
        return p, n, 3
    # ---- depth 10 / 11 clouds whose refinement passes the CPU oracle can still materialise (tens of millions of virtual cells):
    # roots 5 to 8 levels above maxDepth (the certified super-brick path, single-root coarse passes) and 64-bit keys at depth 11
    if name == "tiny_sphere60k_d10":   # dense small object + two far samples fixing the cube: one depth-2 root (8^8 virtual cells), none of the big passes emits
        s, _ = synth.sphere(60_000, seed=11)
        p = np.concatenate([s * np.float32(0.03) + np.array([0.31, 0.27, 0.36], np.float32), np.array([[0, 0, 0], [1, 1, 1]], np.float32)]).astype(np.float32)
        n = np.concatenate([s, np.array([[0, 0, -1], [0, 0, 1]], np.float32)]).astype(np.float32)
        return p, n, 10
    if name == "sparse80_d10":         # 80 samples: roots at depths 3..9, every batched pass emits triangles
        p, n = synth.sphere(80, seed=13)
        return p, n, 10
    if name == "sparse300_d11":        # depth 11 (33-bit keys): roots at depths 3..10, the depth-3 pass has 8^8 virtual cells and emits
        p, n = synth.sphere(300, seed=12)
        return p, n, 11
    raise KeyError(name)


SMALL_CASES = ["sphere3k_d5", "sphere20k_d6", "torus60k_d7", "scan80k_d7", "multi120k_d7", "sphere8k_d8"]
EDGE_CASES = ["one_point_d5", "two_points_d4", "duplicates_d6", "lattice_d5", "zero_normals_d5", "cluster_plus_outlier_d8", "sphere2k_d2", "sphere2k_d3"]
DEEP_CASES = ["tiny_sphere60k_d10", "sparse80_d10", "sparse300_d11"]
