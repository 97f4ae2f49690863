#!/bin/bash
# Evidence for profiles/: launch list + one ncu --set full capture of the dominant kernel + plain bench lines.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 585 -c 200 --csv --log-file gpurun_out/launches_church.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_run.log 2>&1
KREGEX=ff_batch OUT=prof_ff_final bash tools/ncu_ff.sh
ncu --set full --clock-control none --import-source on -k regex:shots_kernel -s 3 -c 1 -f -o gpurun_out/prof_shots_final python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_shots_run.log 2>&1
for wl in church box room synth1m; do
  python bench.py --workload $wl > gpurun_out/bench_${wl}.json 2> gpurun_out/bench_${wl}.err
  tail -c 300 gpurun_out/bench_${wl}.err
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_church_reference_arm.json 2> gpurun_out/bench_ref.err

python - <<'PY'
import json
for n in ("church", "box", "room", "synth1m"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, round(d["ms_per_step"], 3), d["roofline"]["kernel_ms_per_step"], "frac", round(d["roofline"]["frac"], 3), "e2e", d["e2e"])
    except Exception as e:
        print(n, "FAILED", e)
PY
