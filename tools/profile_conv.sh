#!/bin/bash
# launch list of two steps + ncu --set full of the first four conv_tc launches of the second step
set -u
TAG=${1:-r}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches.csv python tools/one_step.py 2 64 > gpurun_out/${TAG}_launches.log 2>&1
$NCU --set full --import-source on -k regex:conv_tc_kernel -s 38 -c 4 -f -o gpurun_out/${TAG}_conv_tc python tools/one_step.py 2 64 > gpurun_out/${TAG}_conv_tc.log 2>&1
