#!/usr/bin/env python3
"""Runs the reference CUDA binary (oracle/_ref/ref_poisson_d<D>, built by oracle/build_ref.py)
on the GPU box with array dumps, runs the CPU oracle on the same input, and compares stage by
stage.  This is the oracle-PINNING step: it is test infrastructure, it never
runs as part of the product.

Outputs (under --out, default gpurun_out/ref_compare_<config>/):
  report.json     per-array comparison oracle-vs-reference (+ reference run-to-run variance)
  golden.json     digests of the reference's own outputs (sha256 of integer arrays, counts,
                  float norms, stage timings) -> committed as tests/golden/ref_<config>.json
  ref_stdout_*.txt  the reference's stage timers
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poissonrecon_gpu_b200 import plyio, synth  # noqa: E402

OCTNODE = np.dtype([("key", "<i4"), ("pidx", "<i4"), ("pnum", "<i4"), ("parent", "<i4"), ("children", "<i4", (8,)),
                    ("neighs", "<i4", (27,)), ("didx", "<i4"), ("dnum", "<i4"), ("vertices", "<i4", (8,)),
                    ("edges", "<i4", (12,)), ("faces", "<i4", (6,)), ("hasTriangle", "<i4"), ("hasIntersection", "<i4")])
VERTEXNODE = np.dtype([("pos", "<f4", (3,)), ("owner", "<i4"), ("kind", "<i4"), ("depth", "<i4"), ("nodes", "<i4", (8,))])
EDGENODE = np.dtype([("kind", "<i4"), ("owner", "<i4"), ("nodes", "<i4", (4,))])
assert OCTNODE.itemsize == 276 and VERTEXNODE.itemsize == 56 and EDGENODE.itemsize == 24


class Orc:
    def __init__(self):
        self.lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "build", "liborc.so"))
        self.lib.orc_create.restype = ctypes.c_void_p
        self.lib.orc_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        self.lib.orc_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_longlong]
        self.lib.orc_get.restype = ctypes.c_longlong
        self.lib.orc_destroy.argtypes = [ctypes.c_void_p]
        self.h = self.lib.orc_create()

    def run(self, p, n, depth, stages=4):
        p = np.ascontiguousarray(p, np.float32)
        n = np.ascontiguousarray(n, np.float32)
        return self.lib.orc_run(self.h, p.ctypes.data, n.ctypes.data, p.shape[0], depth, stages)

    def set(self, name, arr):
        arr = np.ascontiguousarray(arr)
        self.lib.orc_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_longlong]
        r = self.lib.orc_set(self.h, name.encode(), arr.ctypes.data, arr.nbytes)
        assert r == 0, (name, r)

    def stage(self, name):
        self.lib.orc_stage.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        r = self.lib.orc_stage(self.h, name.encode())
        assert r == 0, (name, r)

    def get(self, name, dtype):
        nb = self.lib.orc_get(self.h, name.encode(), None, 0)
        if nb < 0:
            raise KeyError(name)
        a = np.empty(nb // np.dtype(dtype).itemsize, dtype)
        if nb:
            self.lib.orc_get(self.h, name.encode(), a.ctypes.data, nb)
        return a


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def fcmp(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    if a.shape != b.shape:
        return {"shape_mismatch": [list(a.shape), list(b.shape)]}
    if a.size == 0:
        return {"n": 0}
    d = a - b
    nb = float(np.linalg.norm(b))
    return {"n": int(a.size), "max_abs": float(np.abs(d).max()), "rel_l2": float(np.linalg.norm(d) / nb) if nb > 0 else float(np.linalg.norm(d)),
            "n_diff": int((a != b).sum()), "ref_norm": nb}


def icmp(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    if a.shape != b.shape:
        return {"shape_mismatch": [list(a.shape), list(b.shape)]}
    return {"n": int(a.size), "n_diff": int((a != b).sum())}


def load_ref(d):
    r = {}

    def f(name, dt):
        p = os.path.join(d, name + ".bin")
        return np.fromfile(p, dt) if os.path.exists(p) else None
    r["nodes"] = f("nodearray", OCTNODE)
    r["nodes_after"] = f("nodearray_after_mc", OCTNODE)
    r["base"] = f("base", "<i4")
    r["count"] = f("count", "<i4")
    r["points"] = f("points", "<f4")
    r["normals"] = f("normals", "<f4")
    r["p2n"] = f("p2n", "<i4")
    r["center_scale"] = f("center_scale", "<f4")
    r["vectorfield"] = f("vectorfield", "<f4")
    r["divergence"] = f("divergence", "<f4")
    r["x"] = f("x", "<f4")
    r["iso"] = f("iso", "<f4")
    r["pointvalue"] = f("pointvalue", "<f4")
    r["vvalue"] = f("vvalue", "<f4")
    r["vertexarray"] = f("vertexarray", VERTEXNODE)
    r["edgearray"] = f("edgearray", EDGENODE)
    r["vef_sizes"] = f("vef_sizes", "<i4")
    r["subdividenode"] = f("subdividenode", OCTNODE)
    r["mesh_v"] = f("mesh_v", "<f4")
    r["mesh_t"] = f("mesh_t", "<i4")
    r["passes"] = []
    pp = os.path.join(d, "passes.txt")
    if os.path.exists(pp):
        for line in open(pp):
            k, nv, nt = line.split()
            r["passes"].append([k, int(nv), int(nt)])
    r["cg"] = []
    cp = os.path.join(d, "cg.txt")
    if os.path.exists(cp):
        for line in open(cp):
            t = line.split()
            r["cg"].append([int(t[0]), int(t[1]), float(t[2]), float(t[3])])
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="sphere100k_d8")
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--depth", type=int, default=None)
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--out", default=None)
    ap.add_argument("--skip-oracle", action="store_true")
    a = ap.parse_args()
    pts, nrm, depth = synth.make(a.config, a.n)
    if a.depth:
        depth = a.depth
    tag = a.config if a.n is None else f"{a.config}_n{a.n}"
    if a.depth:
        tag += f"_d{a.depth}"
    out = a.out or os.path.join(ROOT, "gpurun_out", "ref_compare_" + tag)
    os.makedirs(out, exist_ok=True)
    work = f"/tmp/refcmp_{tag}"
    os.makedirs(work, exist_ok=True)
    inp = os.path.join(work, "in.ply")
    plyio.write_points_ply(inp, pts, nrm, binary=True)
    refbin = os.path.join(ROOT, "oracle", "_ref", f"ref_poisson_d{depth}")
    report = {"config": tag, "n": int(pts.shape[0]), "depth": depth, "ref_runs": []}
    refs = []
    for k in range(a.runs):
        dd = os.path.join(work, f"dump{k}")
        if os.path.isdir(dd):
            for f in os.listdir(dd):
                os.remove(os.path.join(dd, f))
        os.makedirs(dd, exist_ok=True)
        env = dict(os.environ, REF_DUMP_DIR=dd)
        t0 = time.time()
        pr = subprocess.run([refbin, inp, os.path.join(work, f"out{k}.ply")], env=env, capture_output=True, text=True, timeout=3000)
        wall = time.time() - t0
        open(os.path.join(out, f"ref_stdout_{k}.txt"), "w").write(pr.stdout + "\n--- stderr ---\n" + pr.stderr)
        report["ref_runs"].append({"returncode": pr.returncode, "wall_s": wall})
        if pr.returncode != 0:
            print("reference failed", pr.returncode, pr.stderr[-2000:])
        refs.append(load_ref(dd))
    R = refs[0]
    # plain run without dumps for timing
    t0 = time.time()
    pr = subprocess.run([refbin, inp, os.path.join(work, "out_t.ply")], capture_output=True, text=True, timeout=3000)
    report["ref_plain_wall_s"] = time.time() - t0
    open(os.path.join(out, "ref_stdout_plain.txt"), "w").write(pr.stdout + "\n--- stderr ---\n" + pr.stderr)

    nd = R["nodes"]
    base = R["base"]
    M = nd.shape[0]
    D = depth
    golden = {"config": tag, "n": int(pts.shape[0]), "depth": D, "M": int(M), "base": base.tolist(), "count": R["count"].tolist(),
              "center_scale": R["center_scale"].tolist(), "sha": {}, "float": {}}
    for fld in ("key", "pidx", "pnum", "parent", "neighs", "didx", "dnum"):
        golden["sha"][fld] = sha(nd[fld])
    lt = slice(0, int(base[D]))   # children are only defined below depth D (Q3)
    golden["sha"]["children_lt_D"] = sha(nd["children"][lt])
    golden["sha"]["p2n"] = sha(R["p2n"])
    golden["sha"]["points"] = sha(R["points"])
    golden["sha"]["normals"] = sha(R["normals"])
    for name in ("vectorfield", "divergence", "x", "vvalue", "pointvalue"):
        if R[name] is not None:
            v = R[name].astype(np.float64)
            golden["float"][name] = {"n": int(v.size), "l2": float(np.linalg.norm(v)), "sum": float(v.sum()), "sha": sha(R[name])}
    golden["x_l2_per_depth"] = [float(np.linalg.norm(R["x"][int(base[d]):int(base[d]) + int(R["count"][d])].astype(np.float64))) for d in range(D + 1)]
    golden["div_l2_per_depth"] = [float(np.linalg.norm(R["divergence"][int(base[d]):int(base[d]) + int(R["count"][d])].astype(np.float64))) for d in range(D + 1)]
    golden["iso"] = float(R["iso"][0]) if R["iso"] is not None else None
    golden["vef_sizes"] = R["vef_sizes"].tolist() if R["vef_sizes"] is not None else None
    golden["passes"] = R["passes"]
    golden["cg"] = R["cg"]
    golden["subdivide_num"] = int(R["subdividenode"].shape[0]) if R["subdividenode"] is not None else None
    golden["subdivide_ids_sha"] = sha(R["subdividenode"]["neighs"][:, 13]) if R["subdividenode"] is not None else None
    golden["mesh"] = {"nv": int(R["mesh_v"].size // 3) if R["mesh_v"] is not None else 0, "nt": int(R["mesh_t"].size // 3) if R["mesh_t"] is not None else 0}
    if R["mesh_v"] is not None:
        golden["mesh"]["v_sha"] = sha(R["mesh_v"])
        golden["mesh"]["t_sha"] = sha(R["mesh_t"])
        mv = R["mesh_v"].reshape(-1, 3).astype(np.float64)
        golden["mesh"]["v_sum"] = mv.sum(axis=0).tolist()
    # small float fixtures: coarse-depth x / divergence values (depth <= 3)
    nb3 = int(base[4]) if D >= 4 else M
    golden["x_coarse"] = R["x"][:nb3].astype(np.float64).tolist()
    golden["div_coarse"] = R["divergence"][:nb3].astype(np.float64).tolist()
    json.dump(golden, open(os.path.join(out, "golden.json"), "w"))

    # run-to-run variance of the reference itself
    if len(refs) > 1:
        S = refs[1]
        var = {}
        for fld in ("key", "pidx", "pnum", "parent", "neighs", "didx", "dnum"):
            var[fld] = icmp(S["nodes"][fld], nd[fld])
        for name in ("vectorfield", "divergence", "x", "vvalue", "iso", "mesh_v"):
            if R[name] is not None and S[name] is not None:
                var[name] = fcmp(S[name], R[name])
        var["passes_equal"] = (S["passes"] == R["passes"])
        var["passes_run1"] = S["passes"]
        var["cg_run1"] = S["cg"]
        report["ref_run_to_run"] = var

    if not a.skip_oracle:
        o = Orc()
        t0 = time.time()
        o.run(pts, nrm, depth, 4)
        report["oracle_wall_s"] = time.time() - t0
        cmpd = {}
        cmpd["base"] = icmp(o.get("base", "<i4")[:D + 1], base)
        cmpd["count"] = icmp(o.get("count", "<i4"), R["count"])
        cmpd["center_scale"] = fcmp(o.get("center_scale", "<f4"), R["center_scale"])
        cmpd["points"] = fcmp(o.get("points", "<f4"), R["points"])
        cmpd["normals"] = fcmp(o.get("normals", "<f4"), R["normals"])
        cmpd["p2n"] = icmp(o.get("p2n", "<i4"), R["p2n"])
        okey = o.get("key", "<i8")
        if okey.shape[0] == M:
            cmpd["key"] = icmp(okey, nd["key"].astype(np.int64))
            for fld in ("pidx", "pnum", "parent", "didx", "dnum"):
                cmpd[fld] = icmp(o.get(fld, "<i4"), nd[fld])
                if cmpd[fld]["n_diff"]:
                    bad = np.nonzero(o.get(fld, "<i4") != nd[fld])[0]
                    cmpd[fld]["first_bad"] = [[int(i), int(o.get(fld, "<i4")[i]), int(nd[fld][i])] for i in bad[:10]]
            on = o.get("neighs", "<i4").reshape(-1, 27)
            cmpd["neighs"] = icmp(on, nd["neighs"])
            if cmpd["neighs"]["n_diff"]:
                bad = np.nonzero((on != nd["neighs"]).any(axis=1))[0]
                cmpd["neighs"]["bad_rows"] = int(bad.size)
                cmpd["neighs"]["first_bad_rows"] = [int(i) for i in bad[:10]]
            oc = o.get("children", "<i4").reshape(-1, 8)
            cmpd["children_lt_D"] = icmp(oc[lt], nd["children"][lt])
            if cmpd["children_lt_D"]["n_diff"]:
                bad = np.nonzero((oc[lt] != nd["children"][lt]).any(axis=1))[0]
                cmpd["children_lt_D"]["first_bad_rows"] = [[int(i), oc[i].tolist(), nd["children"][i].tolist()] for i in bad[:6]]
        else:
            cmpd["key"] = {"shape_mismatch": [int(okey.shape[0]), int(M)]}
        for name in ("vectorfield", "divergence", "x", "pointvalue"):
            cmpd[name] = fcmp(o.get(name, "<f4"), R[name])
        ox, od = o.get("x", "<f4"), o.get("divergence", "<f4")
        if ox.shape[0] == M:
            cmpd["x_rel_l2_per_depth"] = []
            cmpd["div_rel_l2_per_depth"] = []
            for d in range(D + 1):
                s = slice(int(base[d]), int(base[d]) + int(R["count"][d]))
                cmpd["x_rel_l2_per_depth"].append(fcmp(ox[s], R["x"][s]).get("rel_l2"))
                cmpd["div_rel_l2_per_depth"].append(fcmp(od[s], R["divergence"][s]).get("rel_l2"))
        cmpd["iso"] = fcmp(o.get("iso", "<f4"), R["iso"])
        cmpd["iso_values"] = [float(o.get("iso", "<f4")[0]), float(R["iso"][0])]
        cmpd["cg_iters_oracle"] = o.get("cg_iters", "<i4").tolist()
        cmpd["cg_iters_ref"] = [c[1] for c in R["cg"]]
        va = R["vertexarray"]
        cmpd["vertex_owner"] = icmp(o.get("vertex_owner", "<i4"), va["owner"])
        cmpd["vertex_kind"] = icmp(o.get("vertex_kind", "<i4"), va["kind"])
        cmpd["vertex_pos"] = fcmp(o.get("vertex_pos", "<f4"), va["pos"])
        cmpd["vvalue"] = fcmp(o.get("vvalue", "<f4"), R["vvalue"])
        ov = o.get("vvalue", "<f4")
        if ov.shape == R["vvalue"].shape:
            cmpd["vvalue_sign_flips"] = int(((ov < 0) != (R["vvalue"] < 0)).sum())
            # with the reference's own x the sign pattern must be identical: evaluate flips by depth
            cmpd["vvalue_maxabs_ref"] = float(np.abs(R["vvalue"]).max())
        ea = R["edgearray"]
        cmpd["edge_owner"] = icmp(o.get("edge_owner", "<i4"), ea["owner"])
        cmpd["edge_kind"] = icmp(o.get("edge_kind", "<i4"), ea["kind"])
        if R["nodes_after"] is not None:
            # back pointers written by the maintain* kernels (1-based ids)
            pass
        cmpd["subdivide"] = icmp(o.get("subdivide", "<i4"), R["subdividenode"]["neighs"][:, 13])
        op = o.get("passes", "<i4").reshape(-1, 3).tolist()
        kinds = {0: "main", 1: "coarse", 2: "finer"}
        cmpd["passes_oracle"] = [[kinds[p[0]], p[1], p[2]] for p in op]
        cmpd["passes_ref"] = R["passes"]
        omv, omt = o.get("mesh_v", "<f4"), o.get("mesh_t", "<i4")
        cmpd["mesh_counts"] = {"oracle": [int(omv.size // 3), int(omt.size // 3)], "ref": [golden["mesh"]["nv"], golden["mesh"]["nt"]]}
        if R["mesh_v"] is not None and omv.shape == R["mesh_v"].shape:
            cmpd["mesh_v"] = fcmp(omv, R["mesh_v"])
        if R["mesh_t"] is not None and omt.shape == R["mesh_t"].shape:
            cmpd["mesh_t"] = icmp(omt, R["mesh_t"])
        report["oracle_vs_ref"] = cmpd
        # ---- teacher-forced, stage by stage: each oracle stage fed with the reference's own input
        if okey.shape[0] == M and cmpd["neighs"]["n_diff"] == 0:
            forced = {}
            o.set("vectorfield", R["vectorfield"])
            o.stage("divergence")
            forced["divergence_given_ref_V"] = fcmp(o.get("divergence", "<f4"), R["divergence"])
            fd = o.get("divergence", "<f4")
            forced["divergence_given_ref_V_per_depth"] = [fcmp(fd[int(base[d]):int(base[d]) + int(R["count"][d])], R["divergence"][int(base[d]):int(base[d]) + int(R["count"][d])]) for d in range(D + 1)]
            o.set("divergence", R["divergence"])
            o.stage("solve")
            fx = o.get("x", "<f4")
            forced["x_given_ref_div_per_depth"] = [fcmp(fx[int(base[d]):int(base[d]) + int(R["count"][d])], R["x"][int(base[d]):int(base[d]) + int(R["count"][d])]) for d in range(D + 1)]
            forced["cg_iters_given_ref_div"] = o.get("cg_iters", "<i4").tolist()
            o.set("x", R["x"])
            o.stage("iso")
            forced["pointvalue_given_ref_x"] = fcmp(o.get("pointvalue", "<f4"), R["pointvalue"])
            forced["iso_given_ref_x"] = [float(o.get("iso", "<f4")[0]), float(R["iso"][0])]
            o.set("iso", R["iso"])
            o.stage("mc")
            fv = o.get("vvalue", "<f4")
            forced["vvalue_given_ref_x_iso"] = fcmp(fv, R["vvalue"])
            if fv.shape == R["vvalue"].shape:
                forced["vvalue_sign_flips"] = int(((fv < 0) != (R["vvalue"] < 0)).sum())
                dep = va["depth"]
                forced["vvalue_n_diff_per_depth"] = [int(((fv != R["vvalue"]) & (dep == d)).sum()) for d in range(D + 1)]
            forced["subdivide"] = icmp(o.get("subdivide", "<i4"), R["subdividenode"]["neighs"][:, 13])
            op = o.get("passes", "<i4").reshape(-1, 3).tolist()
            forced["passes_oracle"] = [[kinds[p[0]], p[1], p[2]] for p in op]
            forced["passes_ref"] = R["passes"]
            omv, omt = o.get("mesh_v", "<f4"), o.get("mesh_t", "<i4")
            forced["mesh_counts"] = {"oracle": [int(omv.size // 3), int(omt.size // 3)], "ref": [golden["mesh"]["nv"], golden["mesh"]["nt"]]}
            if R["mesh_v"] is not None and omv.shape == R["mesh_v"].shape:
                forced["mesh_v"] = fcmp(omv, R["mesh_v"])
            if R["mesh_t"] is not None and omt.shape == R["mesh_t"].shape:
                forced["mesh_t"] = icmp(omt, R["mesh_t"])
            report["oracle_forced_vs_ref"] = forced
    json.dump(report, open(os.path.join(out, "report.json"), "w"), indent=1)
    print(json.dumps(report, indent=1)[:6000])


if __name__ == "__main__":
    main()
