"""Edge cases of the loop on the GPU against the oracle: nothing to shoot, tiny scenes, an emitter too
small to be split (AdaptCalFF's `maxlength < 0.1` exit, shade.cpp:154-155), thresholds at and below zero,
more parallel sources than candidates, a saturating light scale, non-power-of-two grids."""
import numpy as np
import pytest

from conftest import oracle_scene

pytestmark = pytest.mark.gpu

ROOM = """Triangle
200 200 200 0 0 0
0 0 0
40 0 0
40 40 0
Triangle
200 200 200 0 0 0
0 0 0
40 40 0
0 40 0
Triangle
150 100 50 0 0 0
0 0 30
40 40 30
40 0 30
Triangle
150 100 50 0 0 0
0 0 30
0 40 30
40 40 30
"""


def light(size, z=15.0, cx=20.0, cy=20.0, rgb="255 255 255"):
    a, b = cx - size / 2, cx + size / 2
    c, d = cy - size / 2, cy + size / 2
    return (f"Triangle\n{rgb} {rgb}\n{a} {c} {z}\n{b} {c} {z}\n{b} {d} {z}\n"
            f"Triangle\n{rgb} {rgb}\n{a} {c} {z}\n{b} {d} {z}\n{a} {d} {z}\n")


def compare(hr, po, text, shots=-1, model_kw=None, **ctx_kw):
    m = hr.Model(text=text.encode(), **(model_kw or {}))
    okw = {k: v for k, v in ctx_kw.items() if k in ("converge", "parallel_src")}
    sc = oracle_scene(po, m, **okw)
    seq = sc.run(None if shots < 0 else shots)
    live = sc.slots["parent"] >= 0
    for extra in ({}, {"collect_stats": 1}):
        ctx = hr.Context(m, **ctx_kw, **extra)
        done, conv = ctx.step(shots)
        assert done == len(seq)
        assert ctx.shooter_log() == [s[0] for s in seq]
        acc, dB, _ = ctx.download()
        assert np.array_equal(acc[live].view(np.uint32), sc.slots["accB"][live].view(np.uint32))
        assert np.array_equal(dB[live].view(np.uint32), sc.slots["deltaB"][live].view(np.uint32))
        if extra:
            st = ctx.stats()
            assert (st["rays"], st["visible"]) == (sc.stats["rays"], sc.stats["visible"])
    return len(seq), m


def test_no_emitter_converges_immediately(hr, po):
    n, m = compare(hr, po, ROOM)
    assert n == 0
    ctx = hr.Context(m)
    assert ctx.step(-1) == (0, True) and ctx.step(5) == (0, True)


def test_small_scene_to_convergence(hr, po):
    n, _ = compare(hr, po, ROOM + light(8.0))
    assert n > 3


def test_emitter_too_small_to_split(hr, po):
    # longest edge^2 = 2 * 0.2^2 = 0.08 < 0.1: the centre sample alone is used (single ray per vertex)
    n, _ = compare(hr, po, ROOM + light(0.2), shots=6)
    assert n >= 2


def test_zero_and_negative_threshold(hr, po):
    compare(hr, po, ROOM + light(8.0), shots=40, converge=0.0)
    compare(hr, po, ROOM + light(8.0), shots=40, converge=-5.0)      # spent shooters (delta 0) stay candidates


def test_more_parallel_sources_than_candidates(hr, po):
    compare(hr, po, ROOM + light(8.0), shots=12, parallel_src=50)


def test_dim_and_saturating_lights(hr, po):
    compare(hr, po, ROOM + light(8.0), shots=10, model_kw={"light": 0.25})
    compare(hr, po, ROOM + light(8.0), shots=10, model_kw={"light": 3.0})


def check_synthetic(hr, po, leaves, expect_pow2):
    from hw_radiosity_b200 import synth
    m = hr.Model(text=synth.room_scene(leaves, seed=1), tri_limit=3000000, max_tri=3000000)
    assert (m.scene.grid_num in (1, 2, 4, 8, 16, 32, 64)) == expect_pow2
    sc = oracle_scene(po, m, converge=1.0)
    seq = sc.run(2)
    ctx = hr.Context(m, converge=1.0, collect_stats=1)
    assert ctx.step(2)[0] == len(seq) and ctx.shooter_log() == [s[0] for s in seq]
    live = sc.slots["parent"] >= 0
    acc, dB, _ = ctx.download()
    assert np.array_equal(acc[live].view(np.uint32), sc.slots["accB"][live].view(np.uint32))
    st = ctx.stats()
    assert (st["rays"], st["visible"]) == (sc.stats["rays"], sc.stats["visible"])
    fast = hr.Context(m, converge=1.0)
    fast.step(2)
    acc2, _, _ = fast.download()
    assert np.array_equal(acc2[live].view(np.uint32), acc[live].view(np.uint32))
    return st


def test_synthetic_million_leaves_fifty_shots(hr, po):
    """The scaling scene family at 1 M leaves, 50 shots: default (batched) path on one GPU and on two in-process ranks
    against the oracle port - shooter sequence and every accB / deltaB bit."""
    import threading
    from hw_radiosity_b200 import synth
    m = hr.Model(text=synth.room_scene(1000000, seed=1), tri_limit=3000000, max_tri=3000000)
    sc = oracle_scene(po, m, converge=1.0)
    seq = sc.run(50)
    live = sc.slots["parent"] >= 0
    fast = hr.Context(m, converge=1.0)
    assert fast.step(50)[0] == len(seq) == 50 and fast.shooter_log() == [s[0] for s in seq]
    acc, dB, _ = fast.download()
    assert np.array_equal(acc[live].view(np.uint32), sc.slots["accB"][live].view(np.uint32))
    assert np.array_equal(dB[live].view(np.uint32), sc.slots["deltaB"][live].view(np.uint32))
    fast.close()
    ctxs = [hr.Context(m, opts=hr.default_opts(rank=r, world=2, converge=1.0)) for r in range(2)]
    hr.Context.connect_local(ctxs)
    th = [threading.Thread(target=c.step, args=(50,)) for c in ctxs]
    [t.start() for t in th]
    [t.join(timeout=600) for t in th]
    assert not any(t.is_alive() for t in th), "a rank hung"
    for c in ctxs:
        own = c.owned_slots()
        a2, d2, _ = c.download()
        assert c.shooter_log() == [s[0] for s in seq]
        assert np.array_equal(a2[own].view(np.uint32), sc.slots["accB"][own].view(np.uint32))
        assert np.array_equal(d2[own].view(np.uint32), sc.slots["deltaB"][own].view(np.uint32))


def test_synthetic_room(hr, po):
    st = check_synthetic(hr, po, 100000, True)
    assert st["ray_guard_trips"] == 0


def test_non_power_of_two_grid_does_not_hang(hr, po):
    """GridNum = 63: the reference's RayHitted can bounce between two leaves forever on such a scene; the
    device cuts the walk (kMaxWalkSteps) and the oracle applies the same cut, so they still agree."""
    check_synthetic(hr, po, 1000000, False)


def test_round_timeline_and_top64_selection_self_check(hr, models_dir, monkeypatch):
    """Diagnostics build of the round (RADK_DEBUG_PREDICT): the top-64 selection of predict_kernel (registers + shuffles)
    agrees with the full sort in every round, and the timeline shows the chain running BESIDE the form factors (it ends
    before the apply pass starts, and starts long before the form factors can have finished)."""
    monkeypatch.setenv("RADK_DEBUG_PREDICT", "1")
    m = hr.Model(str(models_dir / "box.tri"))
    ctx = hr.Context(m)
    monkeypatch.delenv("RADK_DEBUG_PREDICT")
    done, conv = ctx.step(1500)
    assert done == 1500
    t = ctx.round_timeline()
    assert t is not None and t["rounds"] >= 10
    assert t["top64_selfcheck_mismatches"] == 0
    assert t["predict_end"] < t["ff_batch"] and t["candidate_rays"] <= t["chain"] < t["chain_end"] <= t["apply"] < t["commit"]
    plain = hr.Context(m)
    assert plain.round_timeline() is None
