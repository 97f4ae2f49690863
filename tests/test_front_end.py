"""Host front end (tri reader, octree build, emitter/neighbour init, PrePartition, .fin writer)
against golden vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from conftest import sha

MODELS = ["box", "room", "hall", "church", "blocks", "model1", "model2s", "model1s", "sbox", "model3s"]


@pytest.mark.parametrize("name", MODELS)
def test_frozen_mesh_and_octree_match_reference(hr, golden, models_dir, name):
    g = golden["models"][name]
    m = hr.Model(str(models_dir / f"{name}.tri"))
    assert (m.n0, m.trinum) == (g["n0"], g["trinum"])
    tri = m.triangles
    assert int((tri["parent"] >= 0).sum()) == g["leaves"]
    assert (m.scene.nnodes, m.scene.nlist, m.scene.grid_num) == (g["nodes"], g["listlen"], g["gridnum"])
    assert int(np.float32(m.scene.min_grid_len).view(np.uint32)) == g["mingridlen_bits"]
    assert [int(np.float32(x).view(np.uint32)) for x in m.scene.G0] == g["G0_bits"]
    assert [int(np.float32(x).view(np.uint32)) for x in m.scene.G1] == g["G1_bits"]
    assert int(np.float32(m.sample_area).view(np.uint32)) == g["sample_area_bits"]
    # whole 236-byte records, bit for bit (geometry, zeroed side planes of re-split originals, neighbours, ...)
    assert sha(tri) == g["tristore_sha"]
    assert sha(m.nodes) == g["nodes_sha"]
    assert sha(m.tri_list) == g["list_sha"]


def test_triangle_limit_option(hr, golden, models_dir):
    v = golden["variants"]["room_go"]
    m = hr.Model(str(models_dir / "room.tri"), tri_limit=40000)
    assert m.trinum == v["scene_trinum"]
    assert sha(m.triangles) == v["tristore_sha"]


def test_light_scale_option(hr, golden, models_dir):
    v = golden["variants"]["church_light05"]
    m = hr.Model(str(models_dir / "church.tri"), light=0.5)
    assert sha(m.triangles) == v["tristore_sha"]


def test_parse_from_memory_equals_load(hr, models_dir):
    text = (models_dir / "blocks.tri").read_bytes()
    a = hr.Model(str(models_dir / "blocks.tri"))
    b = hr.Model(text=text)
    assert sha(a.triangles) == sha(b.triangles)


def test_missing_file_is_an_io_error(hr):
    with pytest.raises(hr.RadkError) as e:
        hr.Model("/nonexistent/none.tri")
    assert e.value.code == 2 and "File open error" in str(e.value)


def test_malformed_input(hr):
    with pytest.raises(hr.RadkError):
        hr.Model(text=b"Triangle \n1 2 3 0 0 0\n0 0 0\n1 0 0\n")


def test_tree_node_limit(hr, models_dir):
    # the reference exits when the octree outgrows MaxTreeNode (raycast.cpp:175-178)
    with pytest.raises(hr.RadkError) as e:
        hr.Model(str(models_dir / "church.tri"), max_tree_nodes=100)
    assert e.value.code == 3


def test_float_text_matches_reference_writer(hr, po, golden, models_dir):
    """.fin text: run the oracle loop to convergence on small models, write with the product writer,
    compare the file hash with the reference's own PrintOut output."""
    from conftest import oracle_scene
    for name in ["model2s", "model1s", "blocks"]:
        g = golden["models"][name]["full"]
        m = hr.Model(str(models_dir / f"{name}.tri"))
        sc = oracle_scene(po, m)
        sc.run()
        m.set_triangles(sc.slots)
        assert hashlib.sha256(m.format()).hexdigest() == g["fin_sha"], name


def test_writer_number_formatting(hr):
    # one triangle whose coordinates exercise the digit generator: zero, negative zero, tiny,
    # integers, 8-digit truncation, large values
    tri = b"Triangle\n10 20 30 0 0 0\n0 -0.0 1e-12\n123456.789 -0.000123456789 3.14159274\n16777216 1e9 -2.5\n"
    m = hr.Model(text=tri, prepartition=False)
    lines = m.format().decode().splitlines()
    assert lines[0] == "Triangle " and lines[1] == "0 0 0 0 0 0"
    assert lines[2].split()[:3] == ["0.0", "-0.0", "0.0"]
    assert lines[3].split()[:3] == ["123456.7890625", "-0.00012345", "3.14159274"]
    assert lines[4].split()[:3] == ["16777216.0", "1000000000.0", "-2.5"]


def test_colour_snapshot_writer_equals_plain_writer(hr, po, models_dir, tmp_path):
    """radk_model_write_colors (cached geometry text + colour lines, background thread) writes the same bytes as
    the PrintOut restatement, for several snapshots in a row."""
    from conftest import oracle_scene
    m = hr.Model(str(models_dir / "model3s.tri"))
    sc = oracle_scene(po, m)
    for loop in (10, 20, 0):
        sc.run(10)
        m.set_triangles(sc.slots)
        want = m.format()
        rgb = sc.snapshot_colors()
        rgb[sc.slots["parent"] < 0] = -1
        m.write_colors(str(tmp_path / "snap"), rgb, loop=loop)
        name = "snap.fin" if loop == 0 else f"snap.{loop}"
        assert (tmp_path / name).read_bytes() == want


@pytest.mark.parametrize("name", ["box", "blocks", "model1"])
def test_poly_input_equals_sibling_tri(hr, golden, models_dir, name):
    """.poly front end (the reference ships .poly siblings it cannot read itself): fanning the N-gons gives
    exactly the sibling .tri model, so the frozen mesh equals the reference's for that model."""
    m = hr.Model(str(models_dir / f"{name}.poly"))
    g = golden["models"][name]
    assert (m.n0, m.trinum) == (g["n0"], g["trinum"])
    assert sha(m.triangles) == g["tristore_sha"]


def test_malformed_poly(hr):
    with pytest.raises(hr.RadkError):
        hr.Model(text=b"101 101 101 0 0 0 4\n0 0 0\n1 0 0\n")


@pytest.mark.parametrize("mode,trinum,digest", [
    (1, 49774, "4760a300890462d1d0f928fb56a0d6fe24b59281f2f3ccf6b010897f9f8a2383"),
    (2, 195201, "294cf25d6526d117f8f60c78d2fe6f7d5c69cc9cb4b5635239c12a9baf2cc375"),
])
def test_prepartition_of_synthetic_scene_is_pinned(hr, tmp_path, mode, trinum, digest):
    """The seeded synthetic room (bench config 5) through PrePartition as the reference runs it (1) and through the
    exhaustive variant the scaling scenes use (2): pins the work-list order (largest area first, larger slot on ties)
    - these digests were produced with the reference's std::set work list and must survive any faster structure."""
    from hw_radiosity_b200 import synth
    p = tmp_path / "s.tri"
    p.write_bytes(synth.room_scene(60000, seed=3))
    m = hr.Model(str(p), prepartition=mode, tri_limit=4000000, max_tri=4000000)
    assert m.trinum == trinum
    assert hashlib.sha256(np.ascontiguousarray(m.triangles).tobytes()).hexdigest() == digest
