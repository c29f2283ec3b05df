#!/usr/bin/env python3
"""Product (libprb.so, CUDA) vs CPU oracle on one synthetic config, stage by stage.

Test infrastructure: the oracle is only ever the checker.  Two comparisons are made:
  free    both pipelines run end to end on their own intermediates;
  forced  each product stage is fed the oracle's input for that stage (prb_set_array +
          prb_run_stage), which isolates every kernel's own error.
Writes gpurun_out/parity_<tag>.json and prints a summary.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from poissonrecon_gpu_b200 import PoissonRecon, synth  # noqa: E402
from ref_compare import Orc, fcmp, icmp  # noqa: E402


def compare(pr, o, D, forced=False):
    out = {}
    base = o.get("base", "<i4")
    cnt = o.get("count", "<i4")
    out["base"] = icmp(pr.get("base", "<i4"), base)
    out["count"] = icmp(pr.get("count", "<i4"), cnt)
    out["center_scale"] = fcmp(pr.get("center_scale", "<f4"), o.get("center_scale", "<f4"))
    for name, dt in (("points", "<f4"), ("normals", "<f4")):
        out[name] = fcmp(pr.get(name, dt), o.get(name, dt))
    for name in ("sorted_idx", "p2n", "pidx", "pnum", "parent", "didx", "dnum", "neighs"):
        out[name] = icmp(pr.get(name, "<i4"), o.get(name, "<i4"))
    out["key"] = icmp(pr.get("key", "<u8").astype(np.int64), o.get("key", "<i8"))
    out["sorted_key"] = icmp(pr.get("sorted_key", "<u8").astype(np.int64), o.get("sorted_key", "<i8"))
    M = int(base[D + 1])
    lt = int(base[D])
    pc = pr.get("children", "<i4").reshape(-1, 8)
    oc = o.get("children", "<i4").reshape(-1, 8)
    out["children_lt_D"] = icmp(pc[:lt], oc[:lt])
    out["M"] = M
    return out, base, cnt


def per_depth(a, b, base, cnt, D):
    return [fcmp(a[int(base[d]):int(base[d]) + int(cnt[d])], b[int(base[d]):int(base[d]) + int(cnt[d])]).get("rel_l2") for d in range(D + 1)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="sphere100k_d8")
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--depth", type=int, default=None)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    pts, nrm, depth = synth.make(a.config, a.n)
    if a.depth:
        depth = a.depth
    D = depth
    tag = a.config + (f"_n{a.n}" if a.n else "") + (f"_d{a.depth}" if a.depth else "")
    rep = {"config": tag, "n": int(pts.shape[0]), "depth": D}
    o = Orc()
    t0 = time.time()
    o.run(pts, nrm, D, 4)
    rep["oracle_wall_s"] = time.time() - t0
    pr = PoissonRecon(D)
    t0 = time.time()
    pr.set_points(pts, nrm)
    pr.run()
    rep["product_wall_s_first"] = time.time() - t0
    st = pr.stats()
    rep["stats"] = st
    free, base, cnt = compare(pr, o, D)
    for name in ("vectorfield", "divergence", "x", "pointvalue"):
        free[name] = fcmp(pr.get(name, "<f4"), o.get(name, "<f4"))
    free["div_rel_l2_per_depth"] = per_depth(pr.get("divergence", "<f4"), o.get("divergence", "<f4"), base, cnt, D)
    free["x_rel_l2_per_depth"] = per_depth(pr.get("x", "<f4"), o.get("x", "<f4"), base, cnt, D)
    free["iso"] = [float(pr.get("iso", "<f4")[0]), float(o.get("iso", "<f4")[0])]
    free["cg_iters"] = [pr.get("cg_iters", "<i4").tolist(), o.get("cg_iters", "<i4").tolist()]
    free["lap_stencil"] = fcmp(pr.get("lap_stencil", "<f4").reshape(-1, 27)[:, [13, 14, 17, 26]].ravel(), o.get("lap_stencil", "<f4"))
    kinds = {0: "main", 1: "coarse", 2: "finer"}
    free["passes"] = [pr.get("passes", "<i4").reshape(-1, 3).tolist(), o.get("passes", "<i4").reshape(-1, 3).tolist()]
    free["subdivide"] = icmp(pr.get("subdivide", "<i4"), o.get("subdivide", "<i4"))
    pv, pt = pr.mesh()
    ov, ot = o.get("mesh_v", "<f4"), o.get("mesh_t", "<i4")
    free["mesh_counts"] = [[int(pv.shape[0]), int(pt.shape[0])], [int(ov.size // 3), int(ot.size // 3)]]
    if pv.size == ov.size:
        free["mesh_v"] = fcmp(pv, ov)
    if pt.size == ot.size:
        free["mesh_t"] = icmp(pt, ot)
    rep["free"] = free

    # ---- teacher forced
    f = {}
    pr.set("vectorfield", o.get("vectorfield", "<f4"))
    pr.run_stage("divergence")
    f["divergence"] = fcmp(pr.get("divergence", "<f4"), o.get("divergence", "<f4"))
    f["div_n_diff_per_depth"] = [int((pr.get("divergence", "<f4")[int(base[d]):int(base[d + 1])] != o.get("divergence", "<f4")[int(base[d]):int(base[d + 1])]).sum()) for d in range(D + 1)]
    pr.set("divergence", o.get("divergence", "<f4"))
    pr.run_stage("solve")
    f["x_rel_l2_per_depth"] = per_depth(pr.get("x", "<f4"), o.get("x", "<f4"), base, cnt, D)
    f["cg_iters"] = [pr.get("cg_iters", "<i4").tolist(), o.get("cg_iters", "<i4").tolist()]
    pr.set("x", o.get("x", "<f4"))
    pr.run_stage("iso")
    f["pointvalue"] = fcmp(pr.get("pointvalue", "<f4"), o.get("pointvalue", "<f4"))
    f["iso"] = [float(pr.get("iso", "<f4")[0]), float(o.get("iso", "<f4")[0])]
    pr.set("iso", o.get("iso", "<f4"))
    pr.run_stage("extract")
    # corner values: product stores them per (cell, corner) slot at the owner
    vs = pr.get("vvalue_slots", "<f4")
    ow, kd = o.get("vertex_owner", "<i4"), o.get("vertex_kind", "<i4")
    f["vvalue"] = fcmp(vs[8 * ow.astype(np.int64) + kd], o.get("vvalue", "<f4"))
    f["passes"] = [pr.get("passes", "<i4").reshape(-1, 3).tolist(), o.get("passes", "<i4").reshape(-1, 3).tolist()]
    f["subdivide"] = icmp(pr.get("subdivide", "<i4"), o.get("subdivide", "<i4"))
    pv, pt = pr.mesh()
    f["mesh_counts"] = [[int(pv.shape[0]), int(pt.shape[0])], [int(ov.size // 3), int(ot.size // 3)]]
    if pv.size == ov.size:
        f["mesh_v"] = fcmp(pv, ov)
    if pt.size == ot.size:
        f["mesh_t"] = icmp(pt, ot)
    rep["forced"] = f
    # timing: a few warm runs
    times = []
    for _ in range(3):
        pr.set_points(pts, nrm)
        pr.run()
        times.append(pr.stats())
    rep["warm_stats"] = times[-1]
    out = a.out or os.path.join(ROOT, "gpurun_out", f"parity_{tag}.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(rep, open(out, "w"), indent=1)

    def short(d):
        return {k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk in ("n_diff", "rel_l2", "max_abs", "shape_mismatch")}) for k, v in d.items()}
    print("FREE", json.dumps(short(free))[:4000])
    print("FORCED", json.dumps(short(f))[:4000])
    ws = times[-1]
    print("WARM ms:", {k: round(v, 3) for k, v in ws.items() if k.startswith("ms_")}, "launches", ws["kernel_launches"], "cg_iters", ws["cg_iters"][:D + 1])


if __name__ == "__main__":
    main()
