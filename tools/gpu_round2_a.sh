#!/bin/bash
# round-2 GPU check: parity tests, bench line, per-kernel profile print
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_a.log
tail -15 gpurun_out/r02_pytest_a.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_a.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['seconds_per_step'], 'launches', d['gpu_launches'])
print(d['roofline']['kernel_ms_per_step'], d['roofline']['kernel_launches_per_step'])
print(d['shot_roofline'])
PY
