import gzip
import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# tests/test_gpu_multi.py runs several ranks on ONE device, each rank's kernels waiting for the others' mail: every
# stream needs its own hardware queue (default 8 per device) or a waiting kernel can sit in front of the sender
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _cuda_usable():
    """One probe per session: can the product's library create anything on a CUDA device?"""
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a plain `pytest` on a machine without a GPU must not look red: gpu-marked tests are skipped there
    # (with -m gpu on the B200 box they run; the product itself still fails loudly without a device)
    if _cuda_usable():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (no CPU fallback in the product)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def models_dir(tmp_path_factory):
    """The reference's model/*.tri inputs, unpacked from tests/golden/models/*.tri.gz."""
    d = tmp_path_factory.mktemp("models")
    src = os.path.join(GOLDEN_DIR, "models")
    for fn in os.listdir(src):
        if fn.endswith(".tri.gz") or fn.endswith(".poly.gz"):
            with gzip.open(os.path.join(src, fn)) as f, open(d / fn[:-3], "wb") as g:
                g.write(f.read())
    return d


@pytest.fixture(scope="session")
def hr():
    import hw_radiosity_b200 as mod
    mod.load_library()
    return mod


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


def oracle_scene(po, model, **kw):
    """A CPU-oracle scene over a copy of the product front end's frozen mesh + octree."""
    s = model.scene
    return po.Scene(model.triangles.copy(), model.n0, model.nodes.copy(), model.tri_list.copy(), s.G0, s.G1,
                    s.min_grid_len, s.grid_num, sample_area=model.sample_area, **kw)


def state_digest(tri):
    live = tri["parent"] >= 0
    return {"trinum": len(tri), "leaves": int(live.sum()), "accB_sha": sha(tri["accB"][live]),
            "deltaB_sha": sha(tri["deltaB"][live])}


def nan_prone_rays(tri, n0, n, seed=5):
    """Rays that stress the NaN rule of leaf_visit (rt_core.cuh): from receiver vertices toward vertices, edge midpoints and
    centres of leaves that are slot neighbours (siblings, children of the neighbouring original) or random - many run
    along an edge or inside the plane of an occluder through their origin (a1 == a2 == 0, t = 0/0).  Each ray asks for
    its own target's original.  Returns p, v, banned original (or -1), logsrc."""
    rng = np.random.default_rng(seed)
    leaves = np.nonzero(tri["parent"] >= 0)[0]
    a = rng.choice(leaves, n)
    near = np.clip(a + rng.integers(-40, 41, n), 0, len(tri) - 1)
    near = np.where(tri["parent"][near] >= 0, near, a)
    b = np.where(rng.random(n) < 0.7, near, rng.choice(leaves, n))
    p = tri["p"][a, rng.integers(0, 3, n)].copy()
    kind = rng.integers(0, 3, n)
    vsel = tri["p"][b, rng.integers(0, 3, n)]
    mid = ((tri["p"][b, 0] + tri["p"][b, 1]) / 2).astype(np.float32)
    tgt = np.where(kind[:, None] == 0, tri["c"][b], np.where(kind[:, None] == 1, vsel, mid)).astype(np.float32)
    logsrc = tri["parent"][b].astype(np.int32).copy()
    frm = np.where(rng.random(n) < 0.5, tri["parent"][a], -1).astype(np.int32)
    v = (tgt - p).astype(np.float32)
    keep = np.abs(v).sum(axis=1) > 0
    return tuple(np.ascontiguousarray(x[keep]) for x in (p, v, frm, logsrc))
