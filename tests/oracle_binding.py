"""TEST INFRASTRUCTURE ONLY: ctypes binding of the CPU oracle (oracle/build/liborc.so).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
-- never by the product package.  The oracle is a sequential CPU restatement of the
reference pipeline (oracle/poisson_oracle.cpp), pinned against the reference's own CUDA
binary run on a B200 (tests/golden/ref_sphere100k_d8*.json)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "build", "liborc.so")


def ensure_built():
    if not os.path.exists(LIB):
        subprocess.run(["make"], cwd=os.path.join(ROOT, "oracle"), check=True, stdout=subprocess.DEVNULL)
    return LIB


class Oracle:
    def __init__(self):
        self.lib = ctypes.CDLL(ensure_built())
        L = self.lib
        L.orc_create.restype = ctypes.c_void_p
        L.orc_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.orc_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_longlong]
        L.orc_get.restype = ctypes.c_longlong
        L.orc_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_longlong]
        L.orc_stage.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        self.h = L.orc_create()

    def __del__(self):
        try:
            if self.h:
                self.lib.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def run(self, p, n, depth, stages=4):
        """stages: 1 octree, 2 +splat/divergence, 3 +solve/iso, 4 everything."""
        p = np.ascontiguousarray(p, np.float32)
        n = np.ascontiguousarray(n, np.float32)
        r = self.lib.orc_run(self.h, p.ctypes.data, n.ctypes.data, p.shape[0], depth, stages)
        assert r == 0, f"oracle run failed: {r}"

    def set(self, name, arr):
        arr = np.ascontiguousarray(arr)
        r = self.lib.orc_set(self.h, name.encode(), arr.ctypes.data, arr.nbytes)
        assert r == 0, (name, r)

    def stage(self, name):
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
