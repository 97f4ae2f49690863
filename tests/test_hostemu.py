"""CPU-tier check of the code the kernels run: rt_core.cuh (octree walk, triangle test, form factor,
shooter setup) and scene_flatten.hpp (cell table, receiver sort, vertex sharing) compiled with g++
by tests/hostemu and compared bit-for-bit with the oracle.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, nan_prone_rays, oracle_scene


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(ROOT, "tests", "hostemu", "hostemu.cpp")
    so = os.path.join(ROOT, "tests", "hostemu", "libhostemu.so")
    deps = [src] + [os.path.join(ROOT, "hw-radiosity_b200", "csrc", f) for f in ("rt_core.cuh", "shot_core.cuh", "scene_flatten.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "hw-radiosity_b200", "csrc"),
                               src, "-o", so])
    lib = C.CDLL(so)
    lib.emu_run.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
    lib.emu_rays.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p]
    lib.emu_sees.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
    lib.emu_set_mode.argtypes = [C.c_int]
    lib.emu_run_chain.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
    return lib


def run_emu(emu, m, shots, rank=0, world=1, converge=1200.0):
    n = m.trinum
    accB = np.zeros((n, 3, 3), np.float32)
    dB = np.zeros((n, 3), np.float32)
    sh = np.zeros(max(shots, 1), np.int32)
    ns = C.c_int()
    cnt = (C.c_longlong * 8)()
    assert emu.emu_run(C.addressof(m.scene), converge, rank, world, shots, accB.ctypes.data, dB.ctypes.data,
                       sh.ctypes.data, C.byref(ns), cnt) == 0
    return accB, dB, list(sh[: ns.value]), list(cnt)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name,shots", [("box", 6), ("church", 6), ("blocks", 8), ("sbox", 40), ("model3s", 40)])
def test_device_core_matches_oracle(hr, po, emu, models_dir, name, shots, mode):
    """mode 0: cal_ff/ray_walk (per-shot kernels); mode 1: ff_geom + ray_sees (bounded any-hit leaf visit of ff_batch_kernel)."""
    emu.emu_set_mode(mode)
    m = hr.Model(str(models_dir / f"{name}.tri"))
    sc = oracle_scene(po, m)
    seq = sc.run(shots)
    accB, dB, sh, cnt = run_emu(emu, m, shots)
    live = sc.slots["parent"] >= 0
    assert sh == [s[0] for s in seq]
    assert np.array_equal(accB[live].view(np.uint32), sc.slots["accB"][live].view(np.uint32))
    assert np.array_equal(dB[live].view(np.uint32), sc.slots["deltaB"][live].view(np.uint32))
    st = sc.stats
    assert cnt[:3] == [st["rays"], st["visible"], st["receivers"]]


def test_ray_walk_matches_oracle_on_random_rays(hr, po, emu, models_dir):
    m = hr.Model(str(models_dir / "church.tri"))
    sc = oracle_scene(po, m)
    rng = np.random.default_rng(7)
    tri = m.triangles
    leaves = np.nonzero(tri["parent"] >= 0)[0]
    a, b = rng.choice(leaves, 20000), rng.choice(leaves, 20000)
    p = tri["p"][a, rng.integers(0, 3, len(a))]
    v = tri["c"][b] - p
    frm = tri["parent"][a]
    want = sc.ray_hitted(p, v, frm)
    got = np.empty(len(p), np.int32)
    pp, vv, ff = np.ascontiguousarray(p), np.ascontiguousarray(v), np.ascontiguousarray(frm)
    assert emu.emu_rays(C.addressof(m.scene), pp.ctypes.data, vv.ctypes.data, ff.ctypes.data, len(p), got.ctypes.data) == 0
    assert np.array_equal(got, want)
    assert (want >= 0).mean() > 0.5


@pytest.mark.parametrize("name", ["church", "blocks", "room"])
def test_bounded_any_hit_walk_equals_reference_walk(hr, po, emu, models_dir, name):
    """ray_sees (one bounded any-hit scan per leaf, the walk of ff_batch_kernel) must say "the ray returns logsrc"
    exactly when the reference's RayHitted returns logsrc - also for in-plane rays, where the reference accepts a NaN t
    (0/0) and its nearest-hit scan becomes order dependent (the case leaf_visit hands to the sequential scan)."""
    m = hr.Model(str(models_dir / f"{name}.tri"))
    sc = oracle_scene(po, m)
    rng = np.random.default_rng(11)
    tri = m.triangles
    n0 = m.n0
    leaves = np.nonzero(tri["parent"] >= 0)[0]
    n = 30000
    a, b = rng.choice(leaves, n), rng.choice(leaves, n)
    p = tri["p"][a, rng.integers(0, 3, n)].copy()
    tgt = tri["c"][b].copy()
    logsrc = tri["parent"][b].astype(np.int32).copy()
    frm = tri["parent"][a].astype(np.int32).copy()
    # a third of the rays run inside the plane of their own original toward a sibling on the same original
    # (vertex -> centre of another child of the same parent): a1 == a2 == 0 for every coplanar occluder
    by_parent = {}
    for i in leaves[:: max(1, len(leaves) // 20000)]:
        by_parent.setdefault(int(tri["parent"][i]), []).append(int(i))
    k = 0
    for par, kids in by_parent.items():
        if len(kids) < 2 or k >= n // 3:
            continue
        for x, y in zip(kids[:-1], kids[1:]):
            if k >= n // 3:
                break
            p[k] = tri["p"][x, k % 3]
            tgt[k] = tri["c"][y]
            logsrc[k] = par
            frm[k] = -1 if k % 2 else int(rng.integers(0, n0))      # banned index: none / some other original
            k += 1
    assert k > 1000
    # another part aims at an original but returns to it from an unrelated start (logsrc listed in many leaves)
    v = (tgt - p).astype(np.float32)
    want = sc.ray_hitted(p, v, frm)
    got = np.empty(n, np.int32)
    pp, vv, ff, ll = (np.ascontiguousarray(x) for x in (p, v, frm, logsrc))
    assert emu.emu_sees(C.addressof(m.scene), pp.ctypes.data, vv.ctypes.data, ff.ctypes.data, ll.ctypes.data, n, got.ctypes.data) == 0
    assert np.array_equal(got.astype(bool), want == logsrc)
    assert (want == logsrc).sum() > 100 and (want != logsrc).sum() > 100


def test_sharded_receivers_union_equals_single(hr, emu, models_dir):
    """Each rank's shard, shot with the same global shooter sequence, reproduces its part of the
    single-GPU result (the shooter here is found per shard, so compare only one shot)."""
    m = hr.Model(str(models_dir / "sbox.tri"))
    accB1, dB1, sh1, _ = run_emu(emu, m, 1)
    # world=1 rank=0 again through the sharding path with world 2: shard 0 holds the light? use ranks' union of one shot
    # when the (single) global shooter lives in shard r, that shard's result equals the global one on its receivers.
    hit = 0
    for rank in range(2):
        accB, dB, sh, cnt = run_emu(emu, m, 1, rank=rank, world=2)
        if sh == sh1:
            touched = np.nonzero((accB != 0).any(axis=(1, 2)) | (dB != 0).any(axis=1))[0]
            assert np.array_equal(accB[touched].view(np.uint32), accB1[touched].view(np.uint32))
            hit += 1
    assert hit >= 1


@pytest.mark.parametrize("name,shots,world,rows", [("sbox", -1, 1, 32), ("sbox", 60, 2, 3), ("model3s", -1, 3, 32),
                                                   ("blocks", -1, 2, 32), ("box", 120, 1, 32), ("church", 70, 2, 16)])
def test_chained_shot_loop_matches_oracle(hr, po, emu, models_dir, name, shots, world, rows):
    """The chained loop of fast_kernels.cuh (schedule on the candidate records, one apply pass per round, commit of
    the valid prefix) emulated for `world` ranks with the code of shot_core.cuh: same shooters, same bits."""
    m = hr.Model(str(models_dir / f"{name}.tri"))
    sc = oracle_scene(po, m)
    seq = sc.run(None if shots < 0 else shots)
    n = m.trinum
    accB, dB = np.zeros((n, 3, 3), np.float32), np.zeros((n, 3), np.float32)
    sh = np.zeros(max(len(seq), 1), np.int32)
    ns = C.c_int()
    cnt = (C.c_longlong * 8)()
    assert emu.emu_run_chain(C.addressof(m.scene), 1200.0, world, rows, len(seq) if shots < 0 else shots, accB.ctypes.data,
                             dB.ctypes.data, sh.ctypes.data, C.byref(ns), cnt) == 0
    live = sc.slots["parent"] >= 0
    assert list(sh[: ns.value]) == [s[0] for s in seq]
    assert np.array_equal(accB[live].view(np.uint32), sc.slots["accB"][live].view(np.uint32))
    assert np.array_equal(dB[live].view(np.uint32), sc.slots["deltaB"][live].view(np.uint32))
    assert cnt[1] <= max(1, ns.value)              # rounds


@pytest.mark.parametrize("name,P,shots,world,rows", [("sbox", 4, -1, 1, 32), ("sbox", 4, -1, 2, 8), ("sbox", 10, 40, 3, 64),
                                                     ("blocks", 6, -1, 2, 32), ("model3s", 3, -1, 1, 32)])
def test_chained_loop_with_parallel_sources_matches_oracle(hr, po, emu, models_dir, name, P, shots, world, rows):
    """-parallel_src P on the chained loop: an iteration = min((|C|+4)/5, P) co-shooters (rad.cpp:387-451); the shooter
    count is only scheduled ahead while the lower bound on |C| makes it certain.  Same iterations, same bits."""
    m = hr.Model(str(models_dir / f"{name}.tri"))
    sc = oracle_scene(po, m, parallel_src=P)
    seq = sc.run(None if shots < 0 else shots)
    n = m.trinum
    accB, dB = np.zeros((n, 3, 3), np.float32), np.zeros((n, 3), np.float32)
    sh = np.zeros(max(len(seq), 1), np.int32)
    ns = C.c_int()
    cnt = (C.c_longlong * 8)()
    emu.emu_set_parallel_src(P)
    try:
        assert emu.emu_run_chain(C.addressof(m.scene), 1200.0, world, rows, len(seq) if shots < 0 else shots, accB.ctypes.data,
                                 dB.ctypes.data, sh.ctypes.data, C.byref(ns), cnt) == 0
    finally:
        emu.emu_set_parallel_src(1)
    live = sc.slots["parent"] >= 0
    assert list(sh[: ns.value]) == [s[0] for s in seq]
    assert np.array_equal(accB[live].view(np.uint32), sc.slots["accB"][live].view(np.uint32))
    assert np.array_equal(dB[live].view(np.uint32), sc.slots["deltaB"][live].view(np.uint32))


@pytest.mark.parametrize("name", ["church", "room", "box"])
def test_nan_prone_rays_match_oracle(hr, po, emu, models_dir, name):
    """leaf_visit's NaN rule (an entry with t = 0/0 ends the walk in its leaf; logsrc survives it only with a larger
    index) against the oracle's sequential TriListHitted on rays built to run along edges and inside occluder planes."""
    m = hr.Model(str(models_dir / f"{name}.tri"))
    sc = oracle_scene(po, m)
    p, v, frm, logsrc = nan_prone_rays(m.triangles, m.n0, 60000)
    want = sc.ray_hitted(p, v, frm)
    got = np.empty(len(p), np.int32)
    assert emu.emu_sees(C.addressof(m.scene), p.ctypes.data, v.ctypes.data, frm.ctypes.data, logsrc.ctypes.data, len(p), got.ctypes.data) == 0
    assert np.array_equal(got.astype(bool), want == logsrc)
    assert (want == logsrc).sum() > 1000 and (want != logsrc).sum() > 1000


def test_div3_equals_division(emu):
    """div3 (x * RN(1/3) + one exact residual step, shot_core.cuh) against x / 3.0f: every 509th float bit pattern
    (8.4 M values over the whole range: zeros, denormals, infinities included) plus a dense run around 1.0."""
    emu.emu_div3_check.restype = C.c_longlong
    emu.emu_div3_check.argtypes = [C.c_uint, C.c_uint, C.c_longlong]
    assert emu.emu_div3_check(0, 509, (1 << 32) // 509) == 0
    assert emu.emu_div3_check(0x3f000000, 1, 1 << 22) == 0
    assert emu.emu_div3_check(0x80000000, 1, 1 << 16) == 0 and emu.emu_div3_check(0x7f7ffff0, 1, 32) == 0
