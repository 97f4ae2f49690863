#!/bin/bash
mkdir -p gpurun_out
RADK_NO_GRAPH=1 python tools/run_model.py church 3
RADK_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_church_a.csv python tools/run_model.py church 2 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_launches_church_a.csv 0.5
