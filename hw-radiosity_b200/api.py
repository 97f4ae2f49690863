"""ctypes binding of libradk.so (include/radk.h).

Mirrors the reference's seam for the hot path: `Model` is InitModelFromFile + InitRadiosity +
PrePartitionTriangles (rad.cpp:244-252, 259-356; shade.cpp:556-598), `Context.step` is the
`parallelRadiosity()` loop of DoRadiosity (rad.cpp:499-514), `Context.ray_hitted` is RayHitted
(raycast.h:5).  All computation happens inside the CUDA library; importing this module without
the built library raises, and Context() without a GPU raises RadkError (no CPU fallback).
"""
import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))

TRI_DTYPE = np.dtype([
    ("p", "<f4", (3, 3)), ("Frgb", "<i4", (3,)), ("Brgb", "<i4", (3,)),
    ("accB", "<f4", (3, 3)), ("deltaaccB", "<f4", (3, 3)), ("deltaB", "<f4", (3,)),
    ("se", "<f4", (3, 4)), ("c", "<f4", (3,)), ("n", "<f4", (3,)),
    ("area", "<f4"), ("parent", "<i4"), ("neighbor", "<i4", (3,)),
])
NODE_DTYPE = np.dtype([("g0", "<f4", (3,)), ("g1", "<f4", (3,)), ("sub", "<i4", (8,)), ("list", "<i4")])


class RadkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"radk error {code}: {msg}")
        self.code = code


class SceneStruct(C.Structure):
    _fields_ = [("tri", C.c_void_p), ("n0", C.c_int), ("trinum", C.c_int), ("nodes", C.c_void_p),
                ("nnodes", C.c_int), ("list", C.c_void_p), ("nlist", C.c_int), ("G0", C.c_float * 3),
                ("G1", C.c_float * 3), ("min_grid_len", C.c_float), ("grid_num", C.c_int)]


class Opts(C.Structure):
    _fields_ = [("converge", C.c_float), ("parallel_src", C.c_int), ("tri_limit", C.c_int),
                ("area_limit", C.c_float), ("sample_area", C.c_float), ("delta_ff", C.c_float),
                ("max_tri", C.c_int), ("split_enabled", C.c_int), ("device", C.c_int), ("rank", C.c_int),
                ("world", C.c_int), ("collect_stats", C.c_int), ("dedup", C.c_int), ("per_shot_kernels", C.c_int),
                ("ff_rows", C.c_int), ("reserved", C.c_int * 4)]


class ModelOpts(C.Structure):
    _fields_ = [("light_scale", C.c_float), ("tri_limit", C.c_int), ("max_tri", C.c_int),
                ("max_tree_nodes", C.c_int), ("prepartition", C.c_int), ("sample_area", C.c_float),
                ("reserved", C.c_int * 6)]


STATS_FIELDS = ["shots", "rays", "visible", "calff", "traced_rays", "leaf_visits", "descent_steps", "tri_tests",
                "tri_accepts", "precision_exceeded", "receivers", "splits", "kernel_launches", "ray_guard_trips",
                "ff_rows_filled", "rounds", "shots_scheduled", "rounds_reapplied"]


class Stats(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in STATS_FIELDS]


# every symbol include/radk.h declares (tests check the library exports all of them)
EXPORTS = [
    "radk_default_opts", "radk_last_error", "radk_version", "radk_default_model_opts", "radk_model_load",
    "radk_model_parse", "radk_model_free", "radk_model_scene", "radk_model_sample_area", "radk_model_min_area",
    "radk_model_max_area", "radk_model_triangles", "radk_model_set_triangles", "radk_model_dump",
    "radk_model_write", "radk_model_write_colors", "radk_model_write_wait", "radk_model_format", "radk_ctx_create", "radk_ctx_destroy", "radk_ctx_reset",
    "radk_ctx_rewind", "radk_ctx_step", "radk_ctx_select", "radk_ctx_shoot", "radk_ctx_shooter_log", "radk_ctx_trinum",
    "radk_ctx_download", "radk_ctx_download_triangles", "radk_ctx_snapshot_colors", "radk_ctx_stats",
    "radk_ctx_device_ms", "radk_ctx_round_timeline", "radk_ctx_profile", "radk_ctx_sizes", "radk_ray_hitted", "radk_ray_sees", "radk_vertex_ff",
    "radk_ctx_peer_handle", "radk_ctx_peer_connect", "radk_ctx_peer_connect_local", "radk_ctx_owned_slots",
]

_lib = None


def library_path():
    # RADK_LIB: another build of the same library (A/B of compile-time variants, tools/ff_variant.py); never a fallback
    return os.environ.get("RADK_LIB") or os.path.join(PKG_DIR, "libradk.so")


def build_library(force=False):
    """nvcc-compile libradk.so for sm_100a in-tree (cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", PKG_DIR, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", PKG_DIR, "all"], stdout=subprocess.DEVNULL)
    return library_path()


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `make -C hw-radiosity_b200` "
                          "(__graft_entry__.build()); there is no Python or CPU fallback")
    L = C.CDLL(path)
    vp, ip = C.c_void_p, C.POINTER(C.c_int)
    L.radk_last_error.restype = C.c_char_p
    L.radk_default_opts.argtypes = [C.POINTER(Opts)]
    L.radk_default_model_opts.argtypes = [C.POINTER(ModelOpts)]
    L.radk_model_load.argtypes = [C.c_char_p, C.POINTER(ModelOpts), C.POINTER(vp)]
    L.radk_model_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(ModelOpts), C.POINTER(vp)]
    L.radk_model_free.argtypes = [vp]
    L.radk_model_free.restype = None
    L.radk_model_scene.argtypes = [vp, C.POINTER(SceneStruct)]
    for f in ("radk_model_sample_area", "radk_model_min_area", "radk_model_max_area"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = C.c_float
    L.radk_model_triangles.argtypes = [vp, ip]
    L.radk_model_triangles.restype = vp
    L.radk_model_set_triangles.argtypes = [vp, vp, C.c_int]
    L.radk_model_dump.argtypes = [vp, C.POINTER(Opts), C.c_char_p]
    L.radk_model_write.argtypes = [vp, C.c_char_p, C.c_int]
    L.radk_model_write_colors.argtypes = [vp, C.c_char_p, C.c_int, vp, C.c_int]
    L.radk_model_write_wait.argtypes = [vp]
    L.radk_model_format.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.radk_ctx_create.argtypes = [C.POINTER(SceneStruct), C.POINTER(Opts), C.POINTER(vp)]
    L.radk_ctx_destroy.argtypes = [vp]
    L.radk_ctx_destroy.restype = None
    L.radk_ctx_reset.argtypes = [vp, C.POINTER(SceneStruct)]
    L.radk_ctx_rewind.argtypes = [vp]
    L.radk_ctx_profile.argtypes = [vp, C.c_longlong, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.radk_ctx_sizes.argtypes = [vp, ip, ip, ip, ip]
    L.radk_ctx_step.argtypes = [vp, C.c_longlong, C.POINTER(C.c_longlong), ip]
    L.radk_ctx_select.argtypes = [vp, ip, C.c_int, ip]
    L.radk_ctx_shoot.argtypes = [vp, ip, C.c_int]
    L.radk_ctx_shooter_log.argtypes = [vp, ip, C.c_int, C.POINTER(C.c_longlong)]
    L.radk_ctx_trinum.argtypes = [vp]
    L.radk_ctx_download.argtypes = [vp, vp, vp, vp, C.c_int]
    L.radk_ctx_download_triangles.argtypes = [vp, vp, C.c_int, ip]
    L.radk_ctx_snapshot_colors.argtypes = [vp, vp, C.c_int]
    L.radk_ctx_stats.argtypes = [vp, C.POINTER(Stats)]
    L.radk_ctx_device_ms.argtypes = [vp]
    L.radk_ctx_device_ms.restype = C.c_double
    L.radk_ctx_round_timeline.argtypes = [vp, C.POINTER(C.c_double)]
    L.radk_ray_hitted.argtypes = [vp, vp, vp, vp, C.c_int, vp]
    L.radk_ray_sees.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp]
    L.radk_ctx_peer_handle.argtypes = [vp, vp]
    L.radk_ctx_peer_connect.argtypes = [vp, vp, C.c_int]
    L.radk_ctx_peer_connect_local.argtypes = [vp, C.POINTER(vp), C.c_int]
    L.radk_ctx_owned_slots.argtypes = [vp, vp, C.c_int, ip]
    L.radk_vertex_ff.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise RadkError(rc, load_library().radk_last_error().decode(errors="replace"))


def default_opts(**kw):
    o = Opts()
    load_library().radk_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Model:
    """Host-side scene: .tri file -> frozen mesh + octree (the reference's init phase)."""

    def __init__(self, tri_path=None, text=None, light=1.0, tri_limit=1000000, max_tri=None, prepartition=True,
                 sample_area=-1.0, max_tree_nodes=None):
        L = load_library()
        mo = ModelOpts()
        L.radk_default_model_opts(C.byref(mo))
        mo.light_scale, mo.tri_limit, mo.prepartition, mo.sample_area = light, tri_limit, int(prepartition), sample_area   # prepartition: 0/1/2
        if max_tri:
            mo.max_tri = max_tri
        if max_tree_nodes:
            mo.max_tree_nodes = max_tree_nodes
        self._h = C.c_void_p()
        if tri_path is not None:
            _check(L.radk_model_load(os.fsencode(tri_path), C.byref(mo), C.byref(self._h)))
        else:
            data = text if isinstance(text, bytes) else text.encode()
            _check(L.radk_model_parse(data, len(data), C.byref(mo), C.byref(self._h)))
        self.scene = SceneStruct()
        _check(L.radk_model_scene(self._h, C.byref(self.scene)))

    def close(self):
        if getattr(self, "_h", None):
            load_library().radk_model_free(self._h)
            self._h = None

    __del__ = close

    def _refresh(self):
        _check(load_library().radk_model_scene(self._h, C.byref(self.scene)))

    @property
    def trinum(self):
        return self.scene.trinum

    @property
    def n0(self):
        return self.scene.n0

    @property
    def sample_area(self):
        return load_library().radk_model_sample_area(self._h)

    @property
    def triangles(self):
        """Live (mutable) view of the model's TriStore as a structured array."""
        n = C.c_int()
        ptr = load_library().radk_model_triangles(self._h, C.byref(n))
        buf = (C.c_char * (n.value * TRI_DTYPE.itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=TRI_DTYPE)

    @property
    def nodes(self):
        buf = (C.c_char * (self.scene.nnodes * NODE_DTYPE.itemsize)).from_address(self.scene.nodes)
        return np.frombuffer(buf, dtype=NODE_DTYPE)

    @property
    def tri_list(self):
        buf = (C.c_char * (self.scene.nlist * 4)).from_address(self.scene.list)
        return np.frombuffer(buf, dtype=np.int32)

    def set_triangles(self, tri):
        tri = np.ascontiguousarray(tri, dtype=TRI_DTYPE)
        _check(load_library().radk_model_set_triangles(self._h, tri.ctypes.data, len(tri)))
        self._refresh()

    def dump(self, path, opts=None):
        _check(load_library().radk_model_dump(self._h, C.byref(opts) if opts is not None else None, os.fsencode(path)))

    def write(self, prefix, loop=0):
        _check(load_library().radk_model_write(self._h, os.fsencode(prefix), loop))

    def write_colors(self, prefix, rgb, loop=0, wait=True):
        """Background snapshot from per-slot colours (Context.snapshot_colors()), cached geometry text."""
        rgb = np.ascontiguousarray(rgb, dtype=np.int32)
        _check(load_library().radk_model_write_colors(self._h, os.fsencode(prefix), loop, rgb.ctypes.data, len(rgb)))
        if wait:
            _check(load_library().radk_model_write_wait(self._h))

    def format(self):
        n = C.c_size_t()
        L = load_library()
        _check(L.radk_model_format(self._h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _check(L.radk_model_format(self._h, buf, n.value, C.byref(n)))
        return buf.raw[: n.value]


class Context:
    """Device-resident shooting loop over one shard of the receivers of a Model."""

    def __init__(self, model, opts=None, **kw):
        L = load_library()
        self.model = model
        self.opts = opts if opts is not None else default_opts(**kw)
        self._h = C.c_void_p()
        _check(L.radk_ctx_create(C.byref(model.scene), C.byref(self.opts), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            load_library().radk_ctx_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, model=None):
        _check(load_library().radk_ctx_reset(self._h, C.byref((model or self.model).scene)))

    def rewind(self):
        _check(load_library().radk_ctx_rewind(self._h))

    def sizes(self):
        v = [C.c_int() for _ in range(4)]
        _check(load_library().radk_ctx_sizes(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("receivers", "unique_vertices", "occluders", "nodes"), [x.value for x in v]))

    def profile(self, max_shots=-1):
        """Per-kernel-class device time (CUDA events on the context's stream) over a run."""
        ms, n, shots = (C.c_double * 3)(), (C.c_longlong * 3)(), C.c_longlong()
        _check(load_library().radk_ctx_profile(self._h, max_shots, ms, n, C.byref(shots)))
        names = ("select", "form_factor", "update")
        return {"shots": shots.value, "ms": dict(zip(names, list(ms))), "launches": dict(zip(names, list(n)))}

    def step(self, max_shots=-1):
        done, conv = C.c_longlong(), C.c_int()
        _check(load_library().radk_ctx_step(self._h, max_shots, C.byref(done), C.byref(conv)))
        return done.value, bool(conv.value)

    def select(self, cap=64):
        out, n = (C.c_int * cap)(), C.c_int()
        _check(load_library().radk_ctx_select(self._h, out, cap, C.byref(n)))
        return list(out[: n.value])

    def shooter_log(self, cap=1 << 20):
        n = C.c_longlong()
        L = load_library()
        _check(L.radk_ctx_shooter_log(self._h, None, 0, C.byref(n)))
        k = int(min(n.value, cap))
        out = (C.c_int * max(k, 1))()
        _check(L.radk_ctx_shooter_log(self._h, out, k, C.byref(n)))
        return list(out[:k])

    @property
    def trinum(self):
        return load_library().radk_ctx_trinum(self._h)

    def download(self, out=None):
        """Per-slot (accB[n,3,3], deltaB[n,3], parent[n]); pass a previous result as `out` to reuse its buffers."""
        n = self.trinum
        if out is not None and len(out[0]) == n:
            accB, deltaB, parent = out
        else:
            accB = np.zeros((n, 3, 3), dtype=np.float32)
            deltaB = np.zeros((n, 3), dtype=np.float32)
            parent = np.full(n, -1, dtype=np.int32)
        _check(load_library().radk_ctx_download(self._h, accB.ctypes.data, deltaB.ctypes.data, parent.ctypes.data, n))
        return accB, deltaB, parent

    def download_triangles(self):
        """Adaptive-split contexts: the whole (grown) TriStore as 236-byte records."""
        n = C.c_int()
        L = load_library()
        _check(L.radk_ctx_download_triangles(self._h, None, 0, C.byref(n)))
        tri = np.zeros(n.value, dtype=TRI_DTYPE)
        _check(L.radk_ctx_download_triangles(self._h, tri.ctypes.data, n.value, C.byref(n)))
        return tri

    def download_into(self, model):
        """Merge the device radiosity back into the model's slots (for the .fin writer)."""
        accB, deltaB, parent = self.download()
        tri = model.triangles
        live = parent >= 0
        tri["accB"][live] = accB[live]
        tri["deltaB"][live] = deltaB[live]

    def snapshot_colors(self):
        n = self.trinum
        rgb = np.zeros((n, 3), dtype=np.int32)
        _check(load_library().radk_ctx_snapshot_colors(self._h, rgb.ctypes.data, n))
        return rgb

    def stats(self):
        st = Stats()
        _check(load_library().radk_ctx_stats(self._h, C.byref(st)))
        return {k: int(getattr(st, k)) for k in STATS_FIELDS}

    @property
    def device_ms(self):
        return load_library().radk_ctx_device_ms(self._h)

    def round_timeline(self):
        """Mean start / end times (us after the round's start) of the kernels of a round, for contexts created with
        RADK_DEBUG_PREDICT=1 in the environment; None otherwise."""
        out = (C.c_double * 14)()
        _check(load_library().radk_ctx_round_timeline(self._h, out))
        if out[12] <= 0:
            return None
        names = ("predict", "predict_end", "merge", "merge_end", "candidate_rays", "ff_batch", "chain", "chain_end", "apply", "commit",
                 "commit_end", "gap_to_next_round", "rounds", "top64_selfcheck_mismatches")
        return {n: round(out[i], 1) for i, n in enumerate(names) if out[i] >= 0}

    # ---- multi-GPU introductions -----------------------------------------------------------
    def peer_handle(self):
        buf = C.create_string_buffer(64)
        _check(load_library().radk_ctx_peer_handle(self._h, buf))
        return buf.raw

    def peer_connect(self, handles):
        blob = b"".join(handles)
        _check(load_library().radk_ctx_peer_connect(self._h, blob, len(handles)))

    @staticmethod
    def connect_local(contexts):
        arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
        for c in contexts:
            _check(load_library().radk_ctx_peer_connect_local(c._h, arr, len(contexts)))

    def owned_slots(self):
        n = C.c_int()
        L = load_library()
        _check(L.radk_ctx_owned_slots(self._h, None, 0, C.byref(n)))
        out = np.empty(n.value, dtype=np.int32)
        _check(L.radk_ctx_owned_slots(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def shoot(self, slots):
        arr = (C.c_int * len(slots))(*[int(x) for x in slots])
        _check(load_library().radk_ctx_shoot(self._h, arr, len(slots)))

    def vertex_ff(self, shooter, receivers, vertices):
        r = np.ascontiguousarray(receivers, dtype=np.int32)
        v = np.ascontiguousarray(vertices, dtype=np.int32)
        out = np.empty(len(r), dtype=np.float32)
        _check(load_library().radk_vertex_ff(self._h, int(shooter), r.ctypes.data, v.ctypes.data, len(r), out.ctypes.data))
        return out

    def ray_sees(self, p, v, frm, logsrc):
        """1 where RayHitted(p, v, from, logsrc) == logsrc, by the any-hit walk of the batched form-factor kernel."""
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 3)
        frm = np.ascontiguousarray(frm, dtype=np.int32).reshape(-1)
        out = np.empty(len(p), dtype=np.int32)
        _check(load_library().radk_ray_sees(self._h, p.ctypes.data, v.ctypes.data, frm.ctypes.data, int(logsrc), len(p),
                                            out.ctypes.data))
        return out

    def ray_hitted(self, p, v, frm):
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(-1, 3)
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 3)
        frm = np.ascontiguousarray(frm, dtype=np.int32).reshape(-1)
        out = np.empty(len(p), dtype=np.int32)
        _check(load_library().radk_ray_hitted(self._h, p.ctypes.data, v.ctypes.data, frm.ctypes.data, len(p),
                                              out.ctypes.data))
        return out
