#!/bin/bash
# per-layer kernel time of the conv GEMM under the FSB200_TC_DBG timing switches
for shape in "64 100 100 64 215 3" "64 100 150 64 215 3" "64 150 150 32 107 3" "64 150 225 32 107 3" "64 100 100 64 215 1"; do
  for d in 0 48 7 55 63; do
    t=$(FSB200_TC_DBG=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel --csv python tools/conv_layer_time.py $shape 2>/dev/null | grep conv_tc_kernel | tail -1 | awk -F'","' '{print $NF}' | tr -d '"')
    echo "shape [$shape] DBG=$d  $t"
  done
done
