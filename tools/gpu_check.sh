#!/bin/bash
# parity tests + the three model bench lines (run under gpurun)
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log | head -2
CONFIGS="${CONFIGS:-cur:}" bash tools/gpu_env_ab.sh
