#!/bin/bash
# ncu --set full of kernels matching a regex in the second training step: tools/gpu_ncu.sh <tag> <regex> <skip> <count>
TAG=$1; RE=$2; SKIP=${3:-0}; CNT=${4:-2}
mkdir -p gpurun_out
FSB200_GRAPHS=0 FSB200_NO_OVERLAP=1 ncu --clock-control none --set full --import-source on -k regex:"$RE" -s $SKIP -c $CNT -f -o gpurun_out/${TAG} python tools/one_step.py 2 64 > gpurun_out/${TAG}.log 2>&1
tail -3 gpurun_out/${TAG}.log
