"""ctypes binding of include/prb.h.

Mirrors the reference's stage structure: ``set_points`` ->
``build_octree`` (pipelineBuildNodeArray, main.cu:511) -> ``splat`` (computeVectorField +
divergence, main.cu:3355-3462) -> ``solve`` (LaplacianIteration, main.cu:1223 + iso value,
main.cu:3480) -> ``extract`` (marching cubes + refinement, main.cu:3504-4564); ``run`` is the
whole ``main()``.  Errors surface as :class:`PrbError` carrying ``prb_last_error()``; nothing
here computes on the CPU.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class PrbError(RuntimeError):
    pass


class PrbStats(ctypes.Structure):
    _fields_ = [
        ("n_points", ctypes.c_int64), ("depth", ctypes.c_int32), ("n_nodes", ctypes.c_int32),
        ("nodes_per_depth", ctypes.c_int32 * 16), ("cg_iters", ctypes.c_int32 * 16),
        ("n_subdivide", ctypes.c_int32), ("n_passes", ctypes.c_int32),
        ("n_vertices", ctypes.c_int64), ("n_triangles", ctypes.c_int64),
        ("iso_value", ctypes.c_float), ("center", ctypes.c_float * 3), ("scale", ctypes.c_float),
        ("ms_h2d", ctypes.c_float), ("ms_octree", ctypes.c_float), ("ms_splat", ctypes.c_float),
        ("ms_divergence", ctypes.c_float), ("ms_solve", ctypes.c_float), ("ms_iso", ctypes.c_float),
        ("ms_extract", ctypes.c_float), ("ms_total", ctypes.c_float),
        ("cg_row_iters", ctypes.c_int64), ("kernel_launches", ctypes.c_int32),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


def shard_plan(count: int, world: int):
    """Contiguous split of `count` units over `world` ranks (prb_mg_plan; host only, no GPU)."""
    lib = load_library()
    out = (ctypes.c_int64 * (world + 1))()
    rc = lib.prb_mg_plan(int(count), int(world), out)
    if rc != 0:
        raise PrbError(f"prb_mg_plan: {lib.prb_last_error().decode()}")
    return list(out)


def deal_passes(depth_max: int, depths, counts, world: int):
    """Rank that runs each refinement pass (prb_mg_deal_passes; host only, no GPU)."""
    lib = load_library()
    d = np.ascontiguousarray(depths, np.int32)
    c = np.ascontiguousarray(counts, np.int32)
    o = np.zeros(d.size, np.int32)
    rc = lib.prb_mg_deal_passes(int(depth_max), int(d.size), d.ctypes.data, c.ctypes.data, int(world), o.ctypes.data)
    if rc != 0:
        raise PrbError(f"prb_mg_deal_passes: {lib.prb_last_error().decode()}")
    return o.tolist()


def assemble_mesh(parts, nv: int, nt: int):
    """The whole mesh from every rank's (layout, vertices, triangles): pieces written at their global offsets."""
    V = np.zeros((nv, 3), np.float32)
    T = np.zeros((nt, 3), np.int32)
    for lay_r, v_r, t_r in parts:
        av = at = 0
        for _, vb, pv, tb, pt in np.asarray(lay_r).reshape(-1, 5).tolist():
            V[vb:vb + pv] = v_r[av:av + pv]
            T[tb:tb + pt] = t_r[at:at + pt]
            av += pv; at += pt
    return V, T


def lib_path() -> str:
    return os.path.join(_HERE, "libprb.so")


_lib = None


def load_library():
    """Load libprb.so (built in-tree by csrc/Makefile).  Fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise PrbError(f"{p} not found: build it with `make -C poissonrecon_gpu_b200/csrc` (no CPU fallback exists)")
    lib = ctypes.CDLL(p)
    vp, ci, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.prb_create.argtypes = [ci, ci, ctypes.POINTER(vp)]
    lib.prb_destroy.argtypes = [vp]
    lib.prb_destroy.restype = None
    lib.prb_last_error.restype = ctypes.c_char_p
    lib.prb_set_points.argtypes = [vp, vp, vp, cll]
    lib.prb_set_points_sharded.argtypes = [vp, vp, vp, cll]
    for f in ("prb_build_octree", "prb_splat", "prb_solve", "prb_extract", "prb_run"):
        getattr(lib, f).argtypes = [vp]
    lib.prb_get_mesh.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(cll), ctypes.POINTER(vp), ctypes.POINTER(cll)]
    lib.prb_get_mesh_device.argtypes = lib.prb_get_mesh.argtypes
    lib.prb_get_stats.argtypes = [vp, ctypes.POINTER(PrbStats)]
    lib.prb_get_array.argtypes = [vp, ctypes.c_char_p, vp, cll]
    lib.prb_get_array.restype = cll
    lib.prb_set_array.argtypes = [vp, ctypes.c_char_p, vp, cll]
    lib.prb_set_option.argtypes = [vp, ctypes.c_char_p, ctypes.c_double]
    lib.prb_run_stage.argtypes = [vp, ctypes.c_char_p]
    lib.prb_get_stream.argtypes = [vp, ctypes.POINTER(vp)]
    lib.prb_mg_init.argtypes = [vp, ci, ci, cll, vp]
    lib.prb_mg_set_peer.argtypes = [vp, ci, vp]
    lib.prb_mg_barrier.argtypes = [vp]
    lib.prb_mg_plan.argtypes = [cll, ci, ctypes.POINTER(cll)]
    lib.prb_mg_deal_passes.argtypes = [ci, ci, vp, vp, ci, vp]
    lib.prb_host_tables.argtypes = [ci, ctypes.c_char_p, vp, cll]
    lib.prb_host_tables.restype = cll
    lib.prb_debug_scan.argtypes = [vp, vp, cll, vp, ctypes.POINTER(cll)]
    lib.prb_debug_sort.argtypes = [vp, vp, cll, ci, vp, vp]
    _lib = lib
    return lib


EXPORTS = ["prb_create", "prb_destroy", "prb_last_error", "prb_set_points", "prb_set_points_sharded", "prb_build_octree", "prb_splat", "prb_solve",
           "prb_extract", "prb_run", "prb_get_mesh", "prb_get_mesh_device", "prb_get_stats", "prb_get_array", "prb_set_array",
           "prb_set_option", "prb_run_stage", "prb_get_stream", "prb_host_tables", "prb_mg_init", "prb_mg_set_peer", "prb_mg_barrier",
           "prb_mg_plan", "prb_mg_deal_passes", "prb_debug_scan", "prb_debug_sort"]


class PoissonRecon:
    """One reconstruction context on one GPU (``prb_context``)."""

    def __init__(self, depth: int, device: int = 0):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        self._check(self.lib.prb_create(device, depth, ctypes.byref(self.h)))
        self.depth = depth

    def _check(self, rc):
        if rc != 0:
            raise PrbError(f"prb error {rc}: {self.lib.prb_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.prb_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs: numpy float32 [n,3] (host) or raw device pointers (ints) with n
    def set_points(self, xyz, normals, n: int | None = None):
        if isinstance(xyz, np.ndarray):
            xyz = np.ascontiguousarray(xyz, np.float32)
            normals = np.ascontiguousarray(normals, np.float32)
            self._keep = (xyz, normals)
            n = xyz.shape[0]
            px, pn = xyz.ctypes.data, normals.ctypes.data
        else:
            px, pn = int(xyz), int(normals)
        self._check(self.lib.prb_set_points(self.h, px, pn, n))

    def set_points_sharded(self, xyz_slice, normals_slice, n_total: int):
        """Multi-GPU: this rank's slice [n_total*rank/world, n_total*(rank+1)/world) of the cloud (numpy arrays or raw pointers)."""
        if isinstance(xyz_slice, np.ndarray):
            xyz_slice = np.ascontiguousarray(xyz_slice, np.float32)
            normals_slice = np.ascontiguousarray(normals_slice, np.float32)
            self._keep = (xyz_slice, normals_slice)
            px, pn = xyz_slice.ctypes.data, normals_slice.ctypes.data
        else:
            px, pn = int(xyz_slice), int(normals_slice)
        self._check(self.lib.prb_set_points_sharded(self.h, px, pn, int(n_total)))

    def build_octree(self):
        self._check(self.lib.prb_build_octree(self.h))

    def splat(self):
        self._check(self.lib.prb_splat(self.h))

    def solve(self):
        self._check(self.lib.prb_solve(self.h))

    def extract(self):
        self._check(self.lib.prb_extract(self.h))

    def run(self):
        self._check(self.lib.prb_run(self.h))

    def run_stage(self, name: str):
        self._check(self.lib.prb_run_stage(self.h, name.encode()))

    def set_option(self, key: str, value: float):
        self._check(self.lib.prb_set_option(self.h, key.encode(), float(value)))

    # ---- multi-GPU (one process per GPU; see include/prb.h)
    def mg_init(self, rank: int, world: int, arena_bytes: int) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self._check(self.lib.prb_mg_init(self.h, rank, world, int(arena_bytes), buf))
        return bytes(buf.raw)

    def mg_set_peer(self, peer_rank: int, handle: bytes):
        buf = ctypes.create_string_buffer(handle, 64)
        self._check(self.lib.prb_mg_set_peer(self.h, peer_rank, buf))

    def mg_barrier(self):
        self._check(self.lib.prb_mg_barrier(self.h))

    def mg_setup(self, arena_bytes: int, group=None):
        """Allocate this rank's arena, exchange the IPC handles over torch.distributed (the plumbing
        backend: NCCL or gloo) and open every peer.  All ranks of the group must call it."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        mine = self.mg_init(rank, world, arena_bytes)
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        for r, hd in enumerate(handles):
            if r != rank:
                self.mg_set_peer(r, hd)
        dist.barrier(group)
        self.mg_barrier()
        return rank, world

    def stream(self) -> int:
        """The context's cudaStream_t as an integer (wrap with torch.cuda.ExternalStream)."""
        p = ctypes.c_void_p()
        self._check(self.lib.prb_get_stream(self.h, ctypes.byref(p)))
        return int(p.value or 0)

    def stats(self) -> dict:
        s = PrbStats()
        self._check(self.lib.prb_get_stats(self.h, ctypes.byref(s)))
        return s.as_dict()

    def mesh(self):
        """(vertices float32 [nv,3] in the unit cube, triangles int32 [nt,3]) copied to host."""
        pv, pt = ctypes.c_void_p(), ctypes.c_void_p()
        nv, nt = ctypes.c_int64(), ctypes.c_int64()
        self._check(self.lib.prb_get_mesh(self.h, ctypes.byref(pv), ctypes.byref(nv), ctypes.byref(pt), ctypes.byref(nt)))
        v = np.ctypeslib.as_array(ctypes.cast(pv, ctypes.POINTER(ctypes.c_float)), (nv.value, 3)).copy() if nv.value else np.zeros((0, 3), np.float32)
        t = np.ctypeslib.as_array(ctypes.cast(pt, ctypes.POINTER(ctypes.c_int32)), (nt.value, 3)).copy() if nt.value else np.zeros((0, 3), np.int32)
        return v, t

    def mesh_host_view(self):
        """Like mesh() but returns views of the context's pinned host buffers (valid until the next
        prb_set_points / prb_extract): the device -> host copy without a second host copy."""
        pv, pt = ctypes.c_void_p(), ctypes.c_void_p()
        nv, nt = ctypes.c_int64(), ctypes.c_int64()
        self._check(self.lib.prb_get_mesh(self.h, ctypes.byref(pv), ctypes.byref(nv), ctypes.byref(pt), ctypes.byref(nt)))
        v = np.ctypeslib.as_array(ctypes.cast(pv, ctypes.POINTER(ctypes.c_float)), (nv.value, 3)) if nv.value else np.zeros((0, 3), np.float32)
        t = np.ctypeslib.as_array(ctypes.cast(pt, ctypes.POINTER(ctypes.c_int32)), (nt.value, 3)) if nt.value else np.zeros((0, 3), np.int32)
        return v, t

    def mesh_layout(self):
        """int64 [pieces, 5]: (pass, first global vertex, vertices, first global triangle, triangles) of the pieces this context holds."""
        return self.get("mesh_layout", "<i8").reshape(-1, 5)

    def mesh_global(self, group=None):
        """The whole mesh on every rank of a torch.distributed group (1 GPU: same as mesh()): pieces gathered and written at their global offsets."""
        v, t = self.mesh()
        lay = self.mesh_layout()
        st = self.stats()
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return v, t
        parts = [None] * dist.get_world_size(group)
        dist.all_gather_object(parts, (lay, v, t), group=group)
        return assemble_mesh(parts, st["n_vertices"], st["n_triangles"])

    def mesh_device_size(self):
        nv, nt = ctypes.c_int64(), ctypes.c_int64()
        pv, pt = ctypes.c_void_p(), ctypes.c_void_p()
        self._check(self.lib.prb_get_mesh_device(self.h, ctypes.byref(pv), ctypes.byref(nv), ctypes.byref(pt), ctypes.byref(nt)))
        return nv.value, nt.value

    # ---- unit-test hooks
    def debug_scan(self, a):
        a = np.ascontiguousarray(a, np.int32)
        out = np.empty_like(a)
        tot = ctypes.c_int64()
        self._check(self.lib.prb_debug_scan(self.h, a.ctypes.data, a.size, out.ctypes.data, ctypes.byref(tot)))
        return out, tot.value

    def debug_sort(self, keys, key_bits: int):
        keys = np.ascontiguousarray(keys, np.uint64)
        ok, oi = np.empty_like(keys), np.empty(keys.size, np.int32)
        self._check(self.lib.prb_debug_sort(self.h, keys.ctypes.data, keys.size, key_bits, ok.ctypes.data, oi.ctypes.data))
        return ok, oi

    def get(self, name: str, dtype):
        nb = self.lib.prb_get_array(self.h, name.encode(), None, 0)
        if nb < 0:
            raise PrbError(f"prb_get_array({name}): {self.lib.prb_last_error().decode()}")
        a = np.empty(nb // np.dtype(dtype).itemsize, dtype)
        if nb:
            r = self.lib.prb_get_array(self.h, name.encode(), a.ctypes.data, nb)
            if r < 0:
                raise PrbError(f"prb_get_array({name}): {self.lib.prb_last_error().decode()}")
        return a

    def set(self, name: str, arr):
        arr = np.ascontiguousarray(arr)
        self._check(self.lib.prb_set_array(self.h, name.encode(), arr.ctypes.data, arr.nbytes))
