import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # make is incremental: a no-op when the in-tree libraries are current
    for d in (os.path.join(ROOT, "poissonrecon_gpu_b200", "csrc"), os.path.join(ROOT, "oracle")):
        r = subprocess.run(["make"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise pytest.UsageError(f"build failed in {d}:\n{r.stdout[-3000:]}")


@pytest.fixture(scope="session")
def oracle_cls():
    from tests.oracle_binding import Oracle
    return Oracle


@pytest.fixture(scope="session")
def sphere100k():
    from poissonrecon_gpu_b200 import synth
    return synth.make("sphere100k_d8")


@pytest.fixture(scope="session")
def sphere100k_oracle(sphere100k, oracle_cls):
    """The full oracle run on config 1 (about 8 s): shared by the CPU pinning tests and the GPU parity tests."""
    p, n, d = sphere100k
    o = oracle_cls()
    o.run(p, n, d, 4)
    return o
