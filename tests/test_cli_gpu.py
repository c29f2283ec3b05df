"""The drop-in command line (csrc/main_cli.cpp, replacing the reference's main(), main.cu:3247-4573):
`poisson_recon --in points.ply --out mesh.ply --depth D` produces the same mesh as the library path."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "poissonrecon_gpu_b200", "poisson_recon")


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["ply", "bnpts"])
def test_cli_matches_library(tmp_path, fmt):
    from poissonrecon_gpu_b200 import PoissonRecon, plyio, synth
    assert os.path.exists(EXE), "poisson_recon is not built (python -c 'import __graft_entry__ as g; g.build()')"
    p, n = synth.sphere(20_000, seed=11)
    depth = 6
    inp = str(tmp_path / f"in.{fmt}")
    out = str(tmp_path / "out.ply")
    if fmt == "ply":
        plyio.write_points_ply(inp, p, n)
    else:
        plyio.write_bnpts(inp, p, n)
    r = subprocess.run([EXE, "--in", inp, "--out", out, "--depth", str(depth), "--json"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "Total points number:20000" in r.stdout          # the reference's own progress line (main.cu:573)
    v, t = plyio.read_mesh_ply(out)
    pr = PoissonRecon(depth)
    pr.set_points(p, n)
    pr.run()
    lv, lt = pr.mesh()
    st = pr.stats()
    pr.close()
    assert v.shape == lv.shape and t.shape == lt.shape
    assert np.array_equal(t, lt)
    # the file holds p * scale + center written with %g (6 significant digits, plyfile.cu:2769-2837)
    world = lv * np.float32(st["scale"]) + np.array(st["center"], np.float32)
    assert np.allclose(v, world, rtol=2e-5, atol=1e-6)


@pytest.mark.gpu
def test_cli_errors(tmp_path):
    r = subprocess.run([EXE, "--in", str(tmp_path / "missing.ply"), "--out", str(tmp_path / "o.ply")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert r.returncode != 0
    r = subprocess.run([EXE, "--out", str(tmp_path / "o.ply")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    assert r.returncode == 2 and "usage" in r.stdout


@pytest.mark.gpu
def test_cli_dump_and_weld(tmp_path):
    """--dump writes the parity arrays of prb_get_array as raw files; --weld merges the seam vertices the passes duplicate."""
    from poissonrecon_gpu_b200 import PoissonRecon, plyio, synth
    p, n = synth.sphere(8_000, seed=9)          # sparse deep tree: several refinement passes with seams
    depth = 8
    inp, out, outw, dump = str(tmp_path / "in.bnpts"), str(tmp_path / "out.ply"), str(tmp_path / "outw.ply"), str(tmp_path / "dump")
    plyio.write_bnpts(inp, p, n)
    r = subprocess.run([EXE, "--in", inp, "--out", out, "--depth", str(depth), "--binary", "--dump", dump], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    pr = PoissonRecon(depth)
    pr.set_points(p, n)
    pr.run()
    for name, dt in (("key", "<u8"), ("neighs", "<i4"), ("x", "<f4"), ("divergence", "<f4"), ("passes", "<i4"), ("iso", "<f4")):
        assert np.array_equal(np.fromfile(os.path.join(dump, name + ".bin"), dt), pr.get(name, dt)), name
    v, t = plyio.read_mesh_ply(out)
    r = subprocess.run([EXE, "--in", inp, "--out", outw, "--depth", str(depth), "--binary", "--weld"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    vw, tw = plyio.read_mesh_ply(outw)
    assert tw.shape == t.shape and vw.shape[0] < v.shape[0]
    assert np.array_equal(vw[tw], v[t])                               # same triangles, position by position
    # (the weld compares positions in the unit cube, before the float transform to file coordinates, so two welded vertices may
    # still round to the same file position: what must hold is that no position is lost)
    assert np.unique(vw, axis=0).shape[0] == np.unique(v, axis=0).shape[0]
    pr.close()


@pytest.mark.gpu
def test_cli_two_gpus_matches_one(tmp_path):
    """--gpus 2 (forked ranks, CUDA-IPC arenas, distributed mesh assembled by rank 0) writes the same file as --gpus 1."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    from poissonrecon_gpu_b200 import plyio, synth
    p, n, depth = synth.make("torus1m_d9", 300_000)
    inp = str(tmp_path / "in.bnpts")
    plyio.write_bnpts(inp, p, n)
    outs = []
    for g in (1, 2):
        out = str(tmp_path / f"out{g}.ply")
        r = subprocess.run([EXE, "--in", inp, "--out", out, "--depth", str(depth), "--binary", "--gpus", str(g)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:]
        outs.append(open(out, "rb").read())
    assert outs[0] == outs[1]
