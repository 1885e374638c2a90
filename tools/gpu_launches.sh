#!/bin/bash
# launch list (per-kernel durations) of two canonical training steps, graphs and side-stream overlap off
TAG=${1:-l}
mkdir -p gpurun_out
FSB200_GRAPHS=0 FSB200_NO_OVERLAP=1 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches.csv python tools/one_step.py 2 64 > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv | head -${2:-45}
