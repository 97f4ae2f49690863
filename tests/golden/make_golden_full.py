#!/usr/bin/env python
"""To-convergence goldens of the long-running BASELINE configs (box 4 838 shots, room 16 721 shots, hall),
from the UNMODIFIED reference (oracle/_ref/ref_harness, -O2 -ffp-contract=off; /root/reference/src/rad.cpp:491-523
is the loop being run).  Kept apart from golden.json because the shooter sequences are long: the result goes to
tests/golden/golden_full.json (compact JSON, one entry per model, written after every model so a long run can be
interrupted).

  python tests/golden/make_golden_full.py [box hall room] [--jobs N]

CPU cost at 8 threads: box ~1.5 min, hall ~10 min, room ~25 min.
"""
import gzip
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

OUT = os.path.join(HERE, "golden_full.json")


def main():
    argv = sys.argv[1:]
    jobs = None
    if "--jobs" in argv:
        i = argv.index("--jobs")
        jobs = int(argv[i + 1])
        del argv[i:i + 2]
    names = argv or ["box", "hall", "room"]
    gold = json.load(open(OUT)) if os.path.exists(OUT) else {
        "generator": "oracle/_ref/ref_harness (-O2 -ffp-contract=off build of the unmodified reference), to convergence",
        "models": {}}
    with tempfile.TemporaryDirectory() as tmp:
        for name in names:
            tri = os.path.join(tmp, name + ".tri")
            with open(tri, "wb") as f:
                f.write(gzip.open(os.path.join(HERE, "models", name + ".tri.gz")).read())
            info, res = mg.run_harness(tri, [], -1, tmp, name + "_f", write_fin=True, jobs=jobs)
            res["shooters"] = [s[0] for s in res["shooters"]]      # single-source runs: first (only) shooter of each shot
            res["loop_s"] = info.get("loop_s")
            gold["models"][name] = res
            print(name, res["shots"], res["rays"], info.get("loop_s"), flush=True)
            with open(OUT, "w") as f:
                json.dump(gold, f, separators=(",", ":"), sort_keys=True)
                f.write("\n")


if __name__ == "__main__":
    main()
