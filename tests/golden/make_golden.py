#!/usr/bin/env python
"""Regenerates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref/ref_harness,
the -ffp-contract=off build of /root/reference compiled by oracle/ref_harness/Makefile).

Run in the build container (needs /root/reference): python tests/golden/make_golden.py
The model inputs under tests/golden/models/*.tri.gz are gzip copies of the reference's model/*.tri.
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

# model -> (shots for the prefix check, run to convergence too?)
MODELS = {
    "box": (20, False), "room": (20, False), "hall": (5, False),
    "church": (20, True), "blocks": (20, True), "model1": (20, True), "model2s": (20, True),
    "model1s": (20, True), "sbox": (20, True), "model3s": (20, True),
}
VARIANTS = {  # extra option sets: name -> (model, reference argv, shots)
    "room_go": ("room", ["-triangle", "40000", "-parallel_src", "10"], -1),
    "church_light05": ("church", ["-light", "0.5", "-converge", "600"], 10),
    "sbox_src4": ("sbox", ["-parallel_src", "4"], -1),
}


# split-enabled reference ("variant A", oracle/ref_harness/variantA/config.h), single thread (the
# reference races otherwise).  *_nosample: -sample_area so large that no rand() sample is drawn, which
# makes the split decisions deterministic and exactly reproducible on the GPU.
SPLIT_VARIANTS = {
    "blocks_split": ("blocks", ["-adapt_area", "5"]),
    "blocks_split_nosample": ("blocks", ["-adapt_area", "5", "-sample_area", "1000000000"]),
    "sbox_split_nosample": ("sbox", ["-adapt_area", "2", "-sample_area", "1000000000"]),
    "room_split_nosample": ("room", ["-adapt_area", "20", "-sample_area", "1000000000", "-triangle", "60000"]),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def state_digest(sc):
    live = sc.slots["parent"] >= 0
    return {"trinum": sc.trinum, "leaves": int(live.sum()),
            "accB_sha": sha(sc.slots["accB"][live]), "deltaB_sha": sha(sc.slots["deltaB"][live]),
            "accB_sum": float(sc.slots["accB"][live].astype(np.float64).sum())}


def run_harness(tri, extra, shots, tmp, tag, write_fin=False, binary="ref_harness", jobs=None):
    harness = po.ref_binary(binary)
    assert harness, "build oracle/_ref first (make -C oracle)"
    args = [harness, "-jobs", str(jobs or os.cpu_count()), "-o", os.path.join(tmp, tag), tri] + extra + [
        "--dump-scene", os.path.join(tmp, tag + ".scene"), "--dump-final", os.path.join(tmp, tag + ".final"),
        "--dump-shooters", os.path.join(tmp, tag + ".shooters"), "--json"]
    if shots >= 0:
        args += ["--shots", str(shots)]
    if write_fin:
        args += ["--write-fin"]
    out = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout.decode(errors="replace")
    info = json.loads([l for l in out.splitlines() if l.startswith("HARNESS_JSON")][0][len("HARNESS_JSON "):])
    shooters = [[int(x) for x in l.split()[:2]] for l in open(os.path.join(tmp, tag + ".shooters"))]
    per_shot = {}
    for k, slot in shooters:
        per_shot.setdefault(k, []).append(slot)
    seq = [per_shot[k] for k in sorted(per_shot)][: info["shots"]]
    fin = po.Scene.from_dump(os.path.join(tmp, tag + ".final"))
    res = {"shots": info["shots"], "rays": info["rays"], "visible": info["visible"], "shooters": seq}
    res.update(state_digest(fin))
    if write_fin:
        res["fin_sha"] = hashlib.sha256(open(os.path.join(tmp, tag + ".fin"), "rb").read()).hexdigest()
    return info, res


def main():
    gold = {"generator": "oracle/_ref/ref_harness (-O2 -ffp-contract=off build of the unmodified reference)", "models": {}, "variants": {}}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (nshots, full) in MODELS.items():
            tri = os.path.join(tmp, name + ".tri")
            with open(tri, "wb") as f:
                f.write(gzip.open(os.path.join(HERE, "models", name + ".tri.gz")).read())
            info, prefix = run_harness(tri, [], nshots, tmp, name + "_p")
            sc = po.Scene.from_dump(os.path.join(tmp, name + "_p.scene"))
            g = {
                "n0": info["n0"], "trinum": info["trinum"], "leaves": info["leaves"], "nodes": info["nodes"],
                "listlen": info["listlen"], "gridnum": info["gridnum"],
                "mingridlen_bits": int(np.float32(sc.s.mingridlen).view(np.uint32)),
                "G0_bits": [int(np.float32(x).view(np.uint32)) for x in sc.s.G0],
                "G1_bits": [int(np.float32(x).view(np.uint32)) for x in sc.s.G1],
                "sample_area_bits": int(np.float32(sc.s.sample_area).view(np.uint32)),
                "tristore_sha": sha(sc.slots), "nodes_sha": sha(sc.nodes), "list_sha": sha(sc.list),
                "prefix": prefix,
            }
            if full:
                _, g["full"] = run_harness(tri, [], -1, tmp, name + "_f", write_fin=True)
            gold["models"][name] = g
            print(name, g["trinum"], g["leaves"], prefix["shots"], g.get("full", {}).get("shots"), flush=True)
        for vname, (model, extra, shots) in VARIANTS.items():
            tri = os.path.join(tmp, model + ".tri")
            info, res = run_harness(tri, extra, shots, tmp, vname, write_fin=True)
            sc = po.Scene.from_dump(os.path.join(tmp, vname + ".scene"))
            res.update({"model": model, "argv": extra, "tristore_sha": sha(sc.slots), "scene_trinum": info["trinum"]})
            gold["variants"][vname] = res
            print(vname, info["trinum"], res["shots"], flush=True)
        gold["split_variants"] = {}
        for vname, (model, extra) in SPLIT_VARIANTS.items():
            tri = os.path.join(tmp, model + ".tri")
            info, res = run_harness(tri, extra, -1, tmp, vname, binary="ref_harness_split", jobs=1)
            fin = po.Scene.from_dump(os.path.join(tmp, vname + ".final"))
            live = fin.slots["parent"] >= 0
            res.update({"model": model, "argv": extra, "parent_sha": sha(fin.slots["parent"]), "p_sha": sha(fin.slots["p"][live]),
                        "deltaB_sum": float(fin.slots["deltaB"][live].astype(np.float64).sum())})
            gold["split_variants"][vname] = res
            print(vname, res["trinum"], res["leaves"], res["shots"], flush=True)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
