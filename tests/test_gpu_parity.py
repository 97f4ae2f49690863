"""GPU parity tests proper: the CUDA path, called through the C ABI (libradk.so), against the CPU
oracle on the same inputs and against the reference's golden vectors.

Bar (BASELINE.json north_star): visibility bits identical (parity mode: zero mismatches expected),
per-vertex radiosity within 1e-4 relative - here required BIT-EXACT, which is attainable because the
kernels are compiled -fmad=false in the reference's operation order (SURVEY.md Fact 5)."""
import numpy as np
import pytest

from conftest import oracle_scene, state_digest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4  # north_star tolerance; the asserts below are stricter (bit equality)


def gpu_state(ctx):
    accB, deltaB, parent = ctx.download()
    return accB, deltaB, parent


def assert_state_equal(ctx, sc):
    accB, deltaB, parent = gpu_state(ctx)
    live = sc.slots["parent"] >= 0
    assert np.array_equal(parent[live], sc.slots["parent"][live])
    ref_acc, ref_dB = sc.slots["accB"][live], sc.slots["deltaB"][live]
    # tolerance check first (gives a readable number if something is off), then bit equality
    denom = np.maximum(np.abs(ref_acc), 1e-30)
    rel = np.max(np.abs(accB[live] - ref_acc) / denom)
    assert rel <= REL_TOL, f"max relative accB error {rel}"
    assert np.array_equal(accB[live].view(np.uint32), ref_acc.view(np.uint32))
    assert np.array_equal(deltaB[live].view(np.uint32), ref_dB.view(np.uint32))


@pytest.mark.parametrize("name,shots", [("box", 20), ("church", 20), ("room", 20), ("blocks", 20), ("hall", 5)])
def test_first_shots_match_oracle_and_golden(hr, po, golden, models_dir, name, shots):
    g = golden["models"][name]["prefix"]
    m = hr.Model(str(models_dir / f"{name}.tri"))
    ctx = hr.Context(m, collect_stats=1)
    done, conv = ctx.step(shots)
    assert done == g["shots"]
    assert ctx.shooter_log() == [s[0] for s in g["shooters"]]
    st = ctx.stats()
    assert (st["rays"], st["visible"]) == (g["rays"], g["visible"])      # visibility: zero mismatches
    assert st["precision_exceeded"] == 0 and st["ray_guard_trips"] == 0
    sc = oracle_scene(po, m)
    sc.run(shots)
    assert_state_equal(ctx, sc)
    accB, deltaB, parent = gpu_state(ctx)
    tri = m.triangles.copy()
    live = parent >= 0
    tri["accB"][live], tri["deltaB"][live] = accB[live], deltaB[live]
    d = state_digest(tri)
    assert d["accB_sha"] == g["accB_sha"] and d["deltaB_sha"] == g["deltaB_sha"]


@pytest.mark.parametrize("name", ["church", "blocks", "model1", "model2s", "model1s", "sbox", "model3s"])
def test_convergence_matches_golden_and_fin_bytes(hr, golden, models_dir, name):
    import hashlib
    g = golden["models"][name]["full"]
    m = hr.Model(str(models_dir / f"{name}.tri"))
    ctx = hr.Context(m, collect_stats=1)
    done, conv = ctx.step(-1)
    assert conv and done == g["shots"]
    assert ctx.shooter_log() == [s[0] for s in g["shooters"]]
    st = ctx.stats()
    assert (st["rays"], st["visible"]) == (g["rays"], g["visible"])
    ctx.download_into(m)
    d = state_digest(m.triangles)
    assert d["accB_sha"] == g["accB_sha"] and d["deltaB_sha"] == g["deltaB_sha"]
    assert hashlib.sha256(m.format()).hexdigest() == g["fin_sha"]        # byte-identical .fin


@pytest.mark.parametrize("name", ["church", "blocks", "model1", "sbox", "model3s"])
def test_fast_path_convergence_matches_golden(hr, golden, models_dir, name):
    """The batched path (predict / form-factor cache / persistent shot loop) must shoot the same
    patches in the same order and end in the same bits as the reference."""
    g = golden["models"][name]["full"]
    m = hr.Model(str(models_dir / f"{name}.tri"))
    ctx = hr.Context(m)                      # collect_stats=0 -> fast path
    done, conv = ctx.step(-1)
    assert conv and done == g["shots"]
    assert ctx.shooter_log() == [s[0] for s in g["shooters"]]
    ctx.download_into(m)
    d = state_digest(m.triangles)
    assert d["accB_sha"] == g["accB_sha"] and d["deltaB_sha"] == g["deltaB_sha"]


def test_room_120_shots_match_oracle(hr, po, models_dir):
    """BASELINE config 2 (room.tri -converge 1200 -delta_ff 0.000005 = defaults): N >= 100 shots of the default
    (batched) path against the oracle, shooter sequence and every accB / deltaB bit."""
    m = hr.Model(str(models_dir / "room.tri"))
    ctx = hr.Context(m)
    assert ctx.step(120)[0] == 120
    sc = oracle_scene(po, m)
    seq = sc.run(120)
    assert ctx.shooter_log() == [s[0] for s in seq]
    assert_state_equal(ctx, sc)


def _golden_full():
    import json
    import os
    from conftest import GOLDEN_DIR
    with open(os.path.join(GOLDEN_DIR, "golden_full.json")) as f:
        return json.load(f)["models"]


@pytest.mark.parametrize("name", ["box", "hall", "room"])
def test_long_runs_converge_to_the_reference(hr, models_dir, name):
    """BASELINE configs 1 and 2 (and hall) to convergence: thousands of shots (box 6 294 with the -ffp-contract=off
    reference, hall 9 009, room ~16.7 k; /root/reference/src/rad.cpp:491-523 is the loop) against goldens generated by
    the UNMODIFIED reference (tests/golden/make_golden_full.py): shooter sequence, final accB / deltaB bits, .fin bytes."""
    import hashlib
    gold = _golden_full()
    if name not in gold:
        pytest.skip(f"tests/golden/golden_full.json has no entry for {name} (generate it with make_golden_full.py)")
    g = gold[name]
    m = hr.Model(str(models_dir / f"{name}.tri"))
    ctx = hr.Context(m)                      # the default (batched) path
    done, conv = ctx.step(-1)
    assert conv and done == g["shots"]
    assert ctx.shooter_log() == g["shooters"]
    ctx.download_into(m)
    d = state_digest(m.triangles)
    assert d["accB_sha"] == g["accB_sha"] and d["deltaB_sha"] == g["deltaB_sha"]
    assert hashlib.sha256(m.format()).hexdigest() == g["fin_sha"]        # byte-identical .fin


@pytest.mark.parametrize("name", ["box"])
def test_long_run_ray_counts_match_the_reference(hr, models_dir, name):
    """The statistics path (per-shot kernels) over a whole long run: the reference-equivalent ray and visible-ray
    totals of 6 294 shots equal the unmodified reference's RayHitted hook counts (zero visibility mismatches)."""
    g = _golden_full()[name]
    m = hr.Model(str(models_dir / f"{name}.tri"))
    ctx = hr.Context(m, collect_stats=1)
    done, conv = ctx.step(-1)
    assert conv and done == g["shots"]
    st = ctx.stats()
    assert (st["rays"], st["visible"]) == (g["rays"], g["visible"])
    assert ctx.shooter_log() == g["shooters"]


@pytest.mark.parametrize("name,shots", [("box", 300), ("room", 150), ("hall", 60)])
def test_fast_path_equals_per_shot_kernels(hr, models_dir, name, shots):
    m = hr.Model(str(models_dir / f"{name}.tri"))
    a, b = hr.Context(m), hr.Context(m, per_shot_kernels=1)
    assert a.step(shots)[0] == shots and b.step(shots)[0] == shots
    assert a.shooter_log() == b.shooter_log()
    for x, y in zip(a.download(), b.download()):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_small_cache_and_budget_edges(hr, models_dir):
    m = hr.Model(str(models_dir / "model3s.tri"))
    ref = hr.Context(m, per_shot_kernels=1)
    ref.step(101)
    for rows in (2, 5):
        c = hr.Context(m, ff_rows=rows)
        for k in (1, 2, 33, 65):
            assert c.step(k) == (k, False)
        assert c.shooter_log() == ref.shooter_log()
        for x, y in zip(c.download(), ref.download()):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


@pytest.mark.parametrize("name,rows", [("church", 128), ("church", 64), ("model3s", 96), ("sbox", 128)])
def test_wide_candidate_sets_match_golden(hr, golden, models_dir, name, rows):
    """Up to 128 candidates per round (four per lane of the chain warp, bitonic top-K, CTA-wide cache rebuild): the
    schedule gets longer and is cut by violations more often, the committed shots stay the reference's."""
    g = golden["models"][name]["full"]
    m = hr.Model(str(models_dir / f"{name}.tri"))
    ctx = hr.Context(m, ff_rows=rows)
    done, conv = ctx.step(-1)
    assert conv and done == g["shots"]
    assert ctx.shooter_log() == [s[0] for s in g["shooters"]]
    ctx.download_into(m)
    d = state_digest(m.triangles)
    assert d["accB_sha"] == g["accB_sha"] and d["deltaB_sha"] == g["deltaB_sha"]


def test_stats_off_gives_identical_state(hr, models_dir):
    m = hr.Model(str(models_dir / "sbox.tri"))
    a, b = hr.Context(m, collect_stats=1), hr.Context(m, collect_stats=0)
    a.step(50), b.step(50)
    for x, y in zip(a.download(), b.download()):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    assert a.shooter_log() == b.shooter_log()


def test_step_in_pieces_equals_one_call(hr, models_dir):
    m = hr.Model(str(models_dir / "model3s.tri"))
    a, b = hr.Context(m), hr.Context(m)
    a.step(37)
    for k in (1, 5, 31):
        assert b.step(k)[0] == k
    for x, y in zip(a.download(), b.download()):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_per_receiver_path_equals_shared_vertex_path(hr, models_dir):
    m = hr.Model(str(models_dir / "sbox.tri"))
    a, b = hr.Context(m, collect_stats=1), hr.Context(m, collect_stats=1, dedup=-1)
    a.step(25), b.step(25)
    for x, y in zip(a.download(), b.download()):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    sa, sb = a.stats(), b.stats()
    assert (sa["rays"], sa["visible"], sa["calff"], sa["receivers"]) == (sb["rays"], sb["visible"], sb["calff"], sb["receivers"])
    assert sa["traced_rays"] < sb["traced_rays"]


def test_multi_source_matches_golden(hr, golden, models_dir):
    for vname in ["sbox_src4", "room_go"]:
        g = golden["variants"][vname]
        kw = {"tri_limit": 40000} if vname == "room_go" else {}
        m = hr.Model(str(models_dir / f"{g['model']}.tri"), **kw)
        P = int(g["argv"][g["argv"].index("-parallel_src") + 1])
        ctx = hr.Context(m, parallel_src=P, collect_stats=1)
        done, conv = ctx.step(-1)
        assert conv and done == g["shots"], vname
        st = ctx.stats()
        assert (st["rays"], st["visible"]) == (g["rays"], g["visible"]), vname
        ctx.download_into(m)
        d = state_digest(m.triangles)
        assert d["accB_sha"] == g["accB_sha"] and d["deltaB_sha"] == g["deltaB_sha"], vname


def test_multi_source_batched_path_matches_golden(hr, golden, models_dir):
    """-parallel_src on the default (batched) path: exact top-P selection pass, iterations of co-shooters in the chained
    schedule.  Same iterations, same bits as the unmodified reference (the reference's own go.sh configuration: room_go)."""
    for vname, rows in [("sbox_src4", 0), ("sbox_src4", 8), ("room_go", 0)]:
        g = golden["variants"][vname]
        kw = {"tri_limit": 40000} if vname == "room_go" else {}
        m = hr.Model(str(models_dir / f"{g['model']}.tri"), **kw)
        P = int(g["argv"][g["argv"].index("-parallel_src") + 1])
        ctx = hr.Context(m, opts=hr.default_opts(parallel_src=P, ff_rows=rows))
        done, conv = ctx.step(-1)
        assert conv and done == g["shots"], (vname, done, g["shots"])
        assert ctx.stats()["rounds"] > 0                # the batched path ran
        ctx.download_into(m)
        d = state_digest(m.triangles)
        assert d["accB_sha"] == g["accB_sha"] and d["deltaB_sha"] == g["deltaB_sha"], vname
        assert ctx.shooter_log() == [x[0] for x in g["shooters"]]   # first shooter of every iteration


def test_multi_source_batched_path_in_pieces(hr, po, models_dir):
    """Stepping with a budget (iterations, like the reference's loop counter and -write_cycle) and 50 co-shooters."""
    m = hr.Model(str(models_dir / "model3s.tri"))
    for P, pieces in [(6, [1, 3, 7, -1]), (50, [2, -1])]:
        sc = oracle_scene(po, m, parallel_src=P)
        seq = sc.run(None)
        ctx = hr.Context(m, parallel_src=P)
        total = 0
        for k in pieces:
            done, conv = ctx.step(k)
            total += done
            assert done == k or k < 0 or conv
        assert conv and total == len(seq)
        assert ctx.shooter_log() == [x[0] for x in seq]
        acc, dB, par = ctx.download()
        live = sc.slots["parent"] >= 0
        assert np.array_equal(acc[live].view(np.uint32), sc.slots["accB"][live].view(np.uint32))
        assert np.array_equal(dB[live].view(np.uint32), sc.slots["deltaB"][live].view(np.uint32))


def test_light_and_converge_options(hr, golden, models_dir):
    g = golden["variants"]["church_light05"]
    m = hr.Model(str(models_dir / "church.tri"), light=0.5)
    ctx = hr.Context(m, converge=600.0, collect_stats=1)
    ctx.step(g["shots"])
    assert ctx.shooter_log() == [s[0] for s in g["shooters"]]
    ctx.download_into(m)
    d = state_digest(m.triangles)
    assert d["accB_sha"] == g["accB_sha"] and d["deltaB_sha"] == g["deltaB_sha"]


def test_ray_batch_matches_oracle(hr, po, models_dir):
    """RayHitted parity on explicit rays: vertex -> other triangle's centre, banned = own parent."""
    for name in ["church", "room"]:
        m = hr.Model(str(models_dir / f"{name}.tri"))
        sc = oracle_scene(po, m)
        ctx = hr.Context(m)
        rng = np.random.default_rng(11)
        tri = m.triangles
        leaves = np.nonzero(tri["parent"] >= 0)[0]
        a, b = rng.choice(leaves, 200000), rng.choice(leaves, 200000)
        p = tri["p"][a, rng.integers(0, 3, len(a))]
        v = tri["c"][b] - p
        frm = tri["parent"][a]
        want = sc.ray_hitted(p, v, frm)
        got = ctx.ray_hitted(p, v, frm)
        assert int((got != want).sum()) == 0
        # rays leaving the scene and degenerate directions
        p2 = np.array([[0, 0, 0], [1e9, 0, 0], [0, 0, 0]], np.float32) + np.float32(m.scene.G0[0])
        v2 = np.array([[0, 0, 0], [1, 0, 0], [1e-7, 1e-7, 1]], np.float32)
        assert np.array_equal(ctx.ray_hitted(p2, v2, [-1, -1, -1])[1:], sc.ray_hitted(p2, v2, [-1, -1, -1])[1:])


def test_ray_sees_matches_oracle(hr, po, models_dir):
    """The walk the timed path runs (ray_sees: bounded any-hit leaf visits + per-shooter leaf mask), ray by ray through
    radk_ray_sees, against the oracle's RayHitted(...) == logsrc: rays from receiver vertices toward points of the
    target original (mostly visible or blocked late), toward other triangles (mostly not returning logsrc), rays lying
    IN the plane of an occluder through their origin (a1 = a2 = 0: t is NaN, raycast.cpp:283-313), zero directions."""
    for name in ["church", "room", "blocks"]:
        m = hr.Model(str(models_dir / f"{name}.tri"))
        sc = oracle_scene(po, m)
        ctx = hr.Context(m)
        rng = np.random.default_rng(5)
        tri = m.triangles
        leaves = np.nonzero(tri["parent"] >= 0)[0]
        mismatches = visible = 0
        for logsrc in rng.choice(np.unique(tri["parent"][leaves]), 6, replace=False):
            kids = leaves[tri["parent"][leaves] == logsrc]
            n = 30000
            a = rng.choice(leaves, n)
            p = tri["p"][a, rng.integers(0, 3, n)]
            frm = tri["parent"][a].copy()
            target = tri["c"][rng.choice(kids, n)]                       # a point on the target original
            other = tri["c"][rng.choice(leaves, n)]
            v = np.where((np.arange(n) % 3 == 0)[:, None], other - p, target - p).astype(np.float32)
            # in-plane rays: origin = a vertex of some input triangle, direction = one of its edges (t = 0/0 against it)
            k = 2000
            o = rng.integers(0, m.n0, k)
            p[:k] = tri["p"][o, 0]
            v[:k] = tri["p"][o, 1] - tri["p"][o, 0]
            frm[:k] = -1
            v[k:k + 10] = 0.0
            want = (sc.ray_hitted(p, v, frm) == logsrc).astype(np.int32)
            got = ctx.ray_sees(p, v, frm, int(logsrc))
            mismatches += int((got != want).sum())
            visible += int(want[k + 10:].sum())
        assert mismatches == 0
        assert visible > 0                                               # the sample does contain visible rays


def test_nan_prone_rays_match_oracle_on_device(hr, po, models_dir):
    """The NaN rule of leaf_visit on the device (radk_ray_sees): rays along edges / inside occluder planes, each toward
    its own original, grouped by original for the probe."""
    from conftest import nan_prone_rays
    for name in ["church", "box"]:
        m = hr.Model(str(models_dir / f"{name}.tri"))
        sc = oracle_scene(po, m)
        ctx = hr.Context(m)
        p, v, frm, logsrc = nan_prone_rays(m.triangles, m.n0, 40000, seed=9)
        want = sc.ray_hitted(p, v, frm)
        ids, counts = np.unique(logsrc, return_counts=True)
        checked = mismatches = 0
        for src in ids[np.argsort(-counts)][:120]:      # the originals with the most rays
            sel = logsrc == src
            got = ctx.ray_sees(p[sel], v[sel], frm[sel], int(src))
            mismatches += int((got.astype(bool) != (want[sel] == src)).sum())
            checked += int(sel.sum())
        assert mismatches == 0 and checked > 1000, (name, mismatches, checked)


def test_select_matches_oracle(hr, po, models_dir):
    m = hr.Model(str(models_dir / "church.tri"))
    sc = oracle_scene(po, m, parallel_src=6)
    ctx = hr.Context(m, parallel_src=6)
    assert ctx.select() == sc.select()


def test_snapshot_colors(hr, po, models_dir):
    m = hr.Model(str(models_dir / "model3s.tri"))
    ctx = hr.Context(m)
    ctx.step(100)
    sc = oracle_scene(po, m)
    sc.run(100)
    rgb = ctx.snapshot_colors()
    want = sc.snapshot_colors()
    live = sc.slots["parent"] >= 0
    assert np.array_equal(rgb[live], want[live]) and (rgb[~live] == -1).all()


def test_reset_reruns_identically(hr, models_dir):
    m = hr.Model(str(models_dir / "model1s.tri"))
    ctx = hr.Context(m)
    d1, c1 = ctx.step(-1)
    s1 = [x.copy() for x in ctx.download()]
    ctx.reset()
    d2, c2 = ctx.step(-1)
    assert (d1, c1) == (d2, c2)
    for x, y in zip(s1, ctx.download()):
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_shoot_with_chosen_sources_matches_oracle(hr, po, models_dir):
    m = hr.Model(str(models_dir / "sbox.tri"))
    sc = oracle_scene(po, m)
    ctx = hr.Context(m, per_shot_kernels=1)
    leaves = np.nonzero(m.triangles["parent"] >= 0)[0]
    lights = [int(i) for i in leaves if (m.triangles["deltaB"][i] > 0).any()]
    for srcs in ([lights[0]], lights[1:4]):
        sc.shot_with(srcs)
        ctx.shoot(srcs)
    assert_state_equal(ctx, sc)
    with pytest.raises(hr.RadkError):
        ctx.shoot([int(np.nonzero(m.triangles["parent"] < 0)[0][0])])   # logical slots cannot shoot


def test_vertex_form_factors_match_oracle(hr, po, models_dir):
    m = hr.Model(str(models_dir / "model3s.tri"))
    sc = oracle_scene(po, m)
    ctx = hr.Context(m)
    tri = m.triangles
    leaves = np.nonzero(tri["parent"] >= 0)[0]
    shooter = int(sc.select()[0])
    rng = np.random.default_rng(3)
    recv = rng.choice(leaves, 3000)
    vert = rng.integers(0, 3, len(recv))
    got = ctx.vertex_ff(shooter, recv, vert)
    want = np.array([sc.vertex_ff(shooter, int(r), tri["p"][r, v]) for r, v in zip(recv, vert)], dtype=np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert (want > 0).any()
