"""Structural invariants of the CPU oracle on the shared seeded cases (CPU, seconds).  The same
cases are compared array by array against the CUDA path in tests/test_parity_gpu.py."""
import numpy as np
import pytest

from tests.cases import EDGE_CASES, make_case
from tests.invariants import check_mesh, check_octree


@pytest.mark.parametrize("name", EDGE_CASES + ["sphere3k_d5", "sphere20k_d6"])
def test_oracle_invariants(name, oracle_cls):
    p, n, D = make_case(name)
    o = oracle_cls()
    o.run(p, n, D, 4)
    arrays = {k: o.get(k, "<i4") for k in ("base", "count", "pidx", "pnum", "parent", "didx", "dnum", "neighs", "p2n", "children")}
    arrays["key"] = o.get("key", "<i8")
    arrays["sorted_key"] = o.get("sorted_key", "<i8")
    check_octree(arrays, p.shape[0], D)
    check_mesh(o.get("mesh_v", "<f4").reshape(-1, 3), o.get("mesh_t", "<i4").reshape(-1, 3), o.get("passes", "<i4").reshape(-1, 3),
               finite=(name != "one_point_d5"))   # a single sample has a zero-size bounding box: scale = 0, positions are NaN in the reference too
