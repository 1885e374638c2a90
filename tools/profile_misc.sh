#!/bin/bash
# launch list of two steps + ncu --set full of the block-0 direct conv kernels and the feature kernel (second step)
set -u
TAG=${1:-r}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches.csv python tools/one_step.py 2 64 > gpurun_out/${TAG}_launches.log 2>&1
$NCU --set full --import-source on -k regex:'conv0_fwd_kernel|conv0_bwd_kernel|feat_kernel' -s 3 -c 3 -f -o gpurun_out/${TAG}_conv0 python tools/one_step.py 2 64 > gpurun_out/${TAG}_conv0.log 2>&1
