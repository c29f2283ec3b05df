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
