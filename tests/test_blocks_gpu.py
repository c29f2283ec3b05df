"""Building blocks of the octree stage on the GPU: the single-pass look-back scan and the stable LSD radix sort
(prb_debug_scan / prb_debug_sort), against numpy on ragged sizes around every tile boundary."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pr():
    from poissonrecon_gpu_b200 import PoissonRecon
    p = PoissonRecon(8)
    yield p
    p.close()


@pytest.mark.parametrize("n", [0, 1, 7, 255, 256, 2047, 2048, 2049, 4096, 65535, 65536 + 3, 33 * 2048 + 5, 1_000_003, 5_000_000])
def test_exclusive_scan(pr, n):
    g = np.random.default_rng(n)
    a = g.integers(0, 9, n).astype(np.int32)
    out, tot = pr.debug_scan(a)
    ref = np.concatenate([[0], np.cumsum(a, dtype=np.int64)[:-1]]) if n else np.zeros(0, np.int64)
    assert tot == int(a.sum())
    assert np.array_equal(out.astype(np.int64), ref)


@pytest.mark.parametrize("n,bits", [(1, 6), (31, 9), (4096, 15), (4097, 24), (60_000, 21), (80_000, 21), (100_000, 24), (1_000_000, 27), (3_000_001, 30), (500_000, 33), (400_000, 36)])
def test_radix_sort_is_stable_and_sorted(pr, n, bits):
    g = np.random.default_rng(n + bits)
    # clustered keys (long runs of equal keys, like samples in one leaf) + uniform ones
    keys = np.where(g.random(n) < 0.5, g.integers(0, 1 << bits, n, dtype=np.uint64), g.integers(0, 64, n, dtype=np.uint64) << np.uint64(max(bits - 6, 0)))
    keys = keys.astype(np.uint64) & np.uint64((1 << bits) - 1)
    ok, oi = pr.debug_sort(keys, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(oi.astype(np.int64), order)
    assert np.array_equal(ok, keys[order])
