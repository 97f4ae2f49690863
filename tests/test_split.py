"""Adaptive splitting inside the loop (north_star subsystem 4, BASELINE config 3).

Oracle = the reference built split-enabled ("variant A": include/config.h without DONT_PARTITION /
PRE_PARTITION, see oracle/ref_harness/variantA), single thread.  Goldens: tests/golden/golden.json
"split_variants".  With -sample_area so large that no random interior sample is drawn the split
decisions are deterministic: the GPU pool must then equal the reference's pool exactly - same slots,
same geometry, same radiosity bits.  With sampling on, the reference consumes glibc rand() in visiting
order; the GPU uses a counter-based hash, so that case is compared statistically."""
import numpy as np
import pytest

from conftest import oracle_scene, sha


def split_args(argv):
    o = {"area_limit": 5.0, "sample_area": None, "tri_limit": 1000000}
    for k, v in zip(argv[::2], argv[1::2]):
        if k == "-adapt_area":
            o["area_limit"] = float(v)
        elif k == "-sample_area":
            o["sample_area"] = float(v)
        elif k == "-triangle":
            o["tri_limit"] = int(v)
    return o


def load(hr, models_dir, g):
    o = split_args(g["argv"])
    m = hr.Model(str(models_dir / f"{g['model']}.tri"), prepartition=False, tri_limit=o["tri_limit"],
                 sample_area=o["sample_area"] if o["sample_area"] else -1.0)
    return m, o


@pytest.mark.parametrize("vname", ["blocks_split", "blocks_split_nosample", "sbox_split_nosample", "room_split_nosample"])
def test_oracle_split_variant_pinned(hr, po, golden, models_dir, vname):
    g = golden["split_variants"][vname]
    m, o = load(hr, models_dir, g)
    sc = oracle_scene(po, m, split_enabled=True, capacity=o["tri_limit"], threads=1, area_limit=o["area_limit"], tri_limit=o["tri_limit"])
    seq = sc.run()
    assert (sc.trinum, len(seq)) == (g["trinum"], g["shots"])
    assert seq == g["shooters"]
    live = sc.slots["parent"] >= 0
    assert sha(sc.slots["parent"]) == g["parent_sha"] and sha(sc.slots["accB"][live]) == g["accB_sha"]


@pytest.mark.gpu
@pytest.mark.parametrize("vname", ["blocks_split_nosample", "sbox_split_nosample", "room_split_nosample"])
def test_gpu_split_exact_when_sampling_is_off(hr, golden, models_dir, vname):
    g = golden["split_variants"][vname]
    m, o = load(hr, models_dir, g)
    ctx = hr.Context(m, split_enabled=1, area_limit=o["area_limit"], sample_area=m.sample_area, tri_limit=o["tri_limit"])
    done, conv = ctx.step(-1)
    tri = ctx.download_triangles()
    assert conv and (len(tri), done) == (g["trinum"], g["shots"])
    assert ctx.shooter_log() == [s[0] for s in g["shooters"]]
    live = tri["parent"] >= 0
    assert int(live.sum()) == g["leaves"]
    assert sha(tri["parent"]) == g["parent_sha"]              # same slots became logical, children in the same places
    assert sha(tri["p"][live]) == g["p_sha"]                  # same geometry
    assert sha(tri["accB"][live]) == g["accB_sha"] and sha(tri["deltaB"][live]) == g["deltaB_sha"]
    st = ctx.stats()
    assert st["splits"] > 0 and st["rays"] == g["rays"] and st["visible"] == g["visible"]


@pytest.mark.gpu
def test_gpu_split_with_sampling_is_statistically_equivalent(hr, golden, models_dir):
    g = golden["split_variants"]["blocks_split"]               # BASELINE.json config 3: blocks.tri -adapt_area 5
    m, o = load(hr, models_dir, g)
    ctx = hr.Context(m, split_enabled=1, area_limit=o["area_limit"], sample_area=m.sample_area, tri_limit=o["tri_limit"])
    done, conv = ctx.step(-1)
    tri = ctx.download_triangles()
    live = tri["parent"] >= 0
    assert conv and abs(done - g["shots"]) <= 1
    assert abs(int(live.sum()) - g["leaves"]) <= 0.03 * g["leaves"]            # leaf count within 3 %
    acc = float(tri["accB"][live].astype(np.float64).sum())
    assert abs(acc - g["accB_sum"]) <= 0.03 * abs(g["accB_sum"])               # accumulated energy within 3 %
    # the leaves still tile the input surfaces: total leaf area == total input area
    m0 = hr.Model(str(models_dir / "blocks.tri"), prepartition=False)
    assert abs(float(tri["area"][live].sum()) - float(m0.triangles["area"].sum())) <= 1e-3 * float(m0.triangles["area"].sum())


@pytest.mark.gpu
def test_gpu_split_respects_the_triangle_limit(hr, models_dir):
    m = hr.Model(str(models_dir / "sbox.tri"), prepartition=False, tri_limit=3000, sample_area=1e9)
    ctx = hr.Context(m, split_enabled=1, area_limit=2.0, sample_area=1e9, tri_limit=3000)
    ctx.step(30)
    assert m.trinum < ctx.trinum <= 3000 - 4 + 3
