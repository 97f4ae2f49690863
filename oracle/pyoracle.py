"""TEST INFRASTRUCTURE - ctypes binding of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.

Record layouts mirror the reference's `Triangle` (include/rad.h:25-39, 236 B) and `TreeNode`
(include/rad.h:41-45, 60 B) so raw dumps written by oracle/ref_harness load without conversion.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

TRI_DTYPE = np.dtype([
    ("p", "<f4", (3, 3)), ("Frgb", "<i4", (3,)), ("Brgb", "<i4", (3,)),
    ("accB", "<f4", (3, 3)), ("deltaaccB", "<f4", (3, 3)), ("deltaB", "<f4", (3,)),
    ("se", "<f4", (3, 4)), ("c", "<f4", (3,)), ("n", "<f4", (3,)),
    ("area", "<f4"), ("parent", "<i4"), ("neighbor", "<i4", (3,)),
])
NODE_DTYPE = np.dtype([("g0", "<f4", (3,)), ("g1", "<f4", (3,)), ("sub", "<i4", (8,)), ("list", "<i4")])
assert TRI_DTYPE.itemsize == 236 and NODE_DTYPE.itemsize == 60

HDR_DTYPE = np.dtype([
    ("magic", "S8"), ("n0", "<i4"), ("trinum", "<i4"), ("nnodes", "<i4"), ("nlist", "<i4"),
    ("gridnum", "<i4"), ("mingridlen", "<f4"), ("G0", "<f4", (3,)), ("G1", "<f4", (3,)),
    ("converge", "<f4"), ("parallel_src", "<i4"), ("light", "<f4"), ("tri_limit", "<i4"),
    ("area_limit", "<f4"), ("sample_area", "<f4"), ("delta_ff", "<f4"),
    ("tri_record_bytes", "<i4"), ("node_record_bytes", "<i4"), ("reserved", "<i4", (8,)),
])


class Stats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in (
        "rays", "visible", "calff", "leaf_visits", "descent_steps", "tri_tests", "tri_accepts",
        "precision_exceeded", "receivers", "splits")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class _Scene(C.Structure):
    _fields_ = [
        ("tri", C.c_void_p), ("n0", C.c_int), ("trinum", C.c_int), ("capacity", C.c_int),
        ("nodes", C.c_void_p), ("nnodes", C.c_int), ("list", C.c_void_p), ("nlist", C.c_int),
        ("G0", C.c_float * 3), ("G1", C.c_float * 3), ("mingridlen", C.c_float), ("gridnum", C.c_int),
        ("converge", C.c_float), ("parallel_src", C.c_int), ("tri_limit", C.c_int),
        ("area_limit", C.c_float), ("sample_area", C.c_float), ("delta_ff", C.c_float),
        ("max_tri", C.c_int), ("split_enabled", C.c_int), ("threads", C.c_int), ("stats", Stats),
    ]


_lib = None


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "radiosity_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        P = C.POINTER
        _lib.orc_select.argtypes = [P(_Scene), P(C.c_int), C.c_int]
        _lib.orc_select.restype = C.c_int
        _lib.orc_shot.argtypes = [P(_Scene), P(C.c_int), C.c_int, P(C.c_int)]
        _lib.orc_shot.restype = C.c_int
        _lib.orc_shot_with.argtypes = [P(_Scene), P(C.c_int), C.c_int]
        _lib.orc_shot_with.restype = None
        _lib.orc_ray_hitted_batch.argtypes = [P(_Scene), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib.orc_ray_hitted_batch.restype = None
        _lib.orc_vertex_ff.argtypes = [P(_Scene), C.c_int, C.c_int, P(C.c_float)]
        _lib.orc_vertex_ff.restype = C.c_float
        _lib.orc_snapshot_color.argtypes = [C.c_void_p, P(C.c_int)]
        _lib.orc_snapshot_color.restype = None
    return _lib


class Scene:
    """A frozen (or growing, when split_enabled) scene: TriStore slots + octree + options."""

    def __init__(self, tri, n0, nodes, tlist, G0, G1, mingridlen, gridnum, converge=1200.0,
                 parallel_src=1, tri_limit=1000000, area_limit=5.0, sample_area=-1.0,
                 delta_ff=0.000005, max_tri=1000000, split_enabled=False, threads=None, capacity=None):
        trinum = len(tri)
        capacity = capacity or trinum
        self.tri = np.zeros(capacity, dtype=TRI_DTYPE)
        self.tri[:trinum] = tri
        self.nodes = np.ascontiguousarray(nodes, dtype=NODE_DTYPE)
        self.list = np.ascontiguousarray(tlist, dtype=np.int32)
        s = self.s = _Scene()
        s.tri = self.tri.ctypes.data
        s.n0, s.trinum, s.capacity = int(n0), trinum, capacity
        s.nodes, s.nnodes = self.nodes.ctypes.data, len(self.nodes)
        s.list, s.nlist = self.list.ctypes.data, len(self.list)
        for i in range(3):
            s.G0[i], s.G1[i] = float(G0[i]), float(G1[i])
        s.mingridlen, s.gridnum = float(mingridlen), int(gridnum)
        s.converge, s.parallel_src, s.tri_limit = float(converge), int(parallel_src), int(tri_limit)
        s.area_limit, s.sample_area, s.delta_ff = float(area_limit), float(sample_area), float(delta_ff)
        s.max_tri, s.split_enabled = int(max_tri), int(bool(split_enabled))
        s.threads = int(threads or os.cpu_count() or 1)

    # ---- construction ----------------------------------------------------------------
    @classmethod
    def from_dump(cls, path, **over):
        """Load a RADSCN1 dump (oracle/ref_harness --dump-scene, or the product's own dump)."""
        with open(path, "rb") as f:
            h = np.frombuffer(f.read(HDR_DTYPE.itemsize), dtype=HDR_DTYPE)[0]
            assert h["magic"].startswith(b"RADSCN1"), h["magic"]
            assert h["tri_record_bytes"] == 236 and h["node_record_bytes"] == 60
            tri = np.frombuffer(f.read(236 * int(h["trinum"])), dtype=TRI_DTYPE)
            nodes = np.frombuffer(f.read(60 * int(h["nnodes"])), dtype=NODE_DTYPE)
            tl = np.frombuffer(f.read(4 * int(h["nlist"])), dtype=np.int32)
        kw = dict(converge=h["converge"], parallel_src=h["parallel_src"], tri_limit=h["tri_limit"],
                  area_limit=h["area_limit"], sample_area=h["sample_area"], delta_ff=h["delta_ff"])
        kw.update(over)
        return cls(tri, h["n0"], nodes, tl, h["G0"], h["G1"], h["mingridlen"], h["gridnum"], **kw)

    # ---- views -----------------------------------------------------------------------
    @property
    def trinum(self):
        return int(self.s.trinum)

    @property
    def slots(self):
        return self.tri[: self.trinum]

    @property
    def stats(self):
        return self.s.stats.as_dict()

    def reset_stats(self):
        C.memset(C.byref(self.s.stats), 0, C.sizeof(Stats))

    # ---- the path --------------------------------------------------------------------
    def select(self, cap=64):
        out = (C.c_int * cap)()
        n = lib().orc_select(C.byref(self.s), out, cap)
        return list(out[:n])

    def shot(self):
        """One parallelRadiosity() pass; returns the shooter list ([] when converged)."""
        cap = max(1, int(self.s.parallel_src))
        out = (C.c_int * cap)()
        n = C.c_int(0)
        lib().orc_shot(C.byref(self.s), out, cap, C.byref(n))
        return list(out[: n.value])

    def shot_with(self, shooters):
        arr = (C.c_int * len(shooters))(*shooters)
        lib().orc_shot_with(C.byref(self.s), arr, len(shooters))

    def run(self, max_shots=None):
        seq = []
        while max_shots is None or len(seq) < max_shots:
            sh = self.shot()
            if not sh:
                break
            seq.append(sh)
        return seq

    def ray_hitted(self, p, v, frm):
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 3)
        frm = np.ascontiguousarray(frm, dtype=np.int32).reshape(-1)
        out = np.empty(len(p), dtype=np.int32)
        lib().orc_ray_hitted_batch(C.byref(self.s), p.ctypes.data, v.ctypes.data, frm.ctypes.data,
                                   len(p), out.ctypes.data)
        return out

    def vertex_ff(self, shooter, receiver, p):
        pp = (C.c_float * 3)(*[float(x) for x in p])
        return float(lib().orc_vertex_ff(C.byref(self.s), int(shooter), int(receiver), pp))

    def snapshot_colors(self):
        out = np.zeros((self.trinum, 3), dtype=np.int32)
        rgb = (C.c_int * 3)()
        base = self.tri.ctypes.data
        for i in range(self.trinum):
            lib().orc_snapshot_color(base + 236 * i, rgb)
            out[i] = rgb[:]
        return out


def ref_binary(name="ref_harness"):
    """Path of a prebuilt unmodified-reference binary under oracle/_ref, or None."""
    p = os.path.join(HERE, "_ref", name)
    return p if os.path.exists(p) and os.access(p, os.X_OK) else None
