"""Deterministic synthetic oriented point clouds for the five BASELINE.json configs.

SURVEY.md §8(d) defines the shapes; everything is generated from numpy's PCG64 seeded with
0xC0FFEE00+k so the same cloud is produced here and on the GPU box.  Points are float32
``(N,3)`` positions and float32 ``(N,3)`` normals (unit length; the pipeline rescales them
exactly like the reference does, main.cu:561-568).
"""
from __future__ import annotations

import numpy as np

SEED0 = 0xC0FFEE00


def _rng(k: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(SEED0 + k))


def _unit(v: np.ndarray) -> np.ndarray:
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def sphere(n: int = 100_000, seed: int = 0):
    """Config 1: uniform unit sphere, normal = position."""
    r = _rng(seed)
    p = _unit(r.standard_normal((n, 3)))
    return p.astype(np.float32), p.astype(np.float32)


def torus(n: int = 1_000_000, R: float = 1.0, r: float = 0.35, sigma: float = 0.05, seed: int = 1):
    """Config 2: area-uniform torus, analytic normals + N(0, sigma^2) noise, renormalised."""
    g = _rng(seed)
    # rejection sampling of the minor angle for area uniformity (density ~ R + r cos v)
    m = int(n * 1.6) + 1024
    v = g.uniform(0, 2 * np.pi, m)
    keep = g.uniform(0, R + r, m) < (R + r * np.cos(v))
    v = v[keep][:n]
    while v.shape[0] < n:  # pragma: no cover - extremely unlikely
        vv = g.uniform(0, 2 * np.pi, m)
        kk = g.uniform(0, R + r, m) < (R + r * np.cos(vv))
        v = np.concatenate([v, vv[kk]])[:n]
    u = g.uniform(0, 2 * np.pi, n)
    cu, su, cv, sv = np.cos(u), np.sin(u), np.cos(v), np.sin(v)
    p = np.stack([(R + r * cv) * cu, (R + r * cv) * su, r * sv], axis=1)
    nrm = np.stack([cv * cu, cv * su, sv], axis=1)
    nrm = _unit(nrm + g.normal(0, sigma, (n, 3)))
    return p.astype(np.float32), nrm.astype(np.float32)


def nonuniform_scan(n: int = 5_000_000, kappa: float = 3.0, radial_noise: float = 0.01, seed: int = 2):
    """Config 3: unit sphere with density ~ exp(kappa*cos(theta)) (about 20:1 front/back at
    kappa=3) plus 1% radial noise, analytic normals."""
    g = _rng(seed)
    # von Mises-Fisher around +z: cos(theta) = 1 + log(u + (1-u) e^{-2k}) / k
    u = g.uniform(0, 1, n)
    ct = 1.0 + np.log(u + (1.0 - u) * np.exp(-2.0 * kappa)) / kappa
    st = np.sqrt(np.clip(1.0 - ct * ct, 0, 1))
    ph = g.uniform(0, 2 * np.pi, n)
    d = np.stack([st * np.cos(ph), st * np.sin(ph), ct], axis=1)
    rad = 1.0 + g.normal(0, radial_noise, n)
    p = d * rad[:, None]
    return p.astype(np.float32), d.astype(np.float32)


def multi_object(n: int = 20_000_000, seed: int = 3):
    """Config 4: 64 spheres / tori of radius 0.03-0.12 at jittered 4x4x4 lattice sites."""
    g = _rng(seed)
    sites = np.stack(np.meshgrid(*[np.arange(4)] * 3, indexing="ij"), axis=-1).reshape(-1, 3)
    centres = (sites + 0.5) / 4.0 + g.uniform(-0.03, 0.03, (64, 3))
    radii = g.uniform(0.03, 0.12, 64)
    radii = np.minimum(radii, 0.1)  # keep neighbours apart on the 0.25 lattice
    is_torus = g.uniform(0, 1, 64) < 0.5
    area = np.where(is_torus, 4 * np.pi**2 * radii * (0.35 * radii), 4 * np.pi * radii**2)
    counts = np.floor(n * area / area.sum()).astype(np.int64)
    counts[0] += n - counts.sum()
    ps, ns = [], []
    for k in range(64):
        m = int(counts[k])
        if is_torus[k]:
            p, nr = torus(m, R=1.0, r=0.35, sigma=0.0, seed=1000 + seed * 64 + k)
            p = p * np.float32(radii[k] / 1.35)
        else:
            p, nr = sphere(m, seed=1000 + seed * 64 + k)
            p = p * np.float32(radii[k])
        ps.append(p + centres[k].astype(np.float32))
        ns.append(nr)
    return np.concatenate(ps).astype(np.float32), np.concatenate(ns).astype(np.float32)


def dense_surface(n: int = 100_000_000, seed: int = 4, chunk: int = 4_000_000):
    """Config 5: radius-1 sphere displaced by a few octaves of smooth trigonometric noise;
    normals from the analytic gradient of the implicit r - (1 + h(dir))."""
    g = _rng(seed)
    ps, ns = [], []
    amps = [0.08 / (2**o) for o in range(5)]
    freqs = [2.0 * (2**o) for o in range(5)]
    phases = g.uniform(0, 2 * np.pi, (5, 3))
    done = 0
    while done < n:
        m = min(chunk, n - done)
        d = _unit(g.standard_normal((m, 3)))
        h = np.zeros(m)
        gh = np.zeros((m, 3))
        for a, f, ph in zip(amps, freqs, phases):
            arg = f * d + ph
            h += a * np.sin(arg).sum(axis=1) / 3.0
            gh += a * f * np.cos(arg) / 3.0
        rad = 1.0 + h
        p = d * rad[:, None]
        # gradient of F(x)=|x|-1-h(x/|x|): d - (I - d d^T) gh / |x|
        t = gh - d * (gh * d).sum(axis=1, keepdims=True)
        nr = _unit(d - t / rad[:, None])
        ps.append(p.astype(np.float32))
        ns.append(nr.astype(np.float32))
        done += m
    return np.concatenate(ps), np.concatenate(ns)


CONFIGS = {
    "sphere100k_d8": dict(gen=sphere, n=100_000, depth=8),
    "torus1m_d9": dict(gen=torus, n=1_000_000, depth=9),
    "scan5m_d10": dict(gen=nonuniform_scan, n=5_000_000, depth=10),
    "multi20m_d11": dict(gen=multi_object, n=20_000_000, depth=11),
    "dense100m_d12": dict(gen=dense_surface, n=100_000_000, depth=12),
}


def make(name: str, n: int | None = None):
    """Return (points, normals, depth) for a named config (optionally with a smaller N)."""
    c = CONFIGS[name]
    p, nr = c["gen"](n if n is not None else c["n"])
    return p, nr, c["depth"]
