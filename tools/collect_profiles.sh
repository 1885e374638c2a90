#!/bin/bash
# Copies the judged summaries of a tools/gpu_check.sh run from gpurun_out/ (scratch) into profiles/ (tracked).
# usage: tools/collect_profiles.sh <gpurun tag, e.g. r02p> <profile tag, e.g. r02>
set -u
SRC=${1:-r02p}
DST=${2:-r02}
G=gpurun_out
P=profiles
mkdir -p $P
python tools/launch_summary.py $G/${SRC}_launches.csv > $P/${DST}_launch_summary.txt
{
  echo "# ncu --set full --clock-control none --import-source on, python tools/one_step.py 2 64 (graphs and side-stream overlap off),"
  echo "# second (warm) step; tools/gpu_check.sh.  columns: duration, DRAM read / write per launch, DRAM GB/s, ncu dram % of its"
  echo "# own peak, tensor-pipe active %, registers, achieved occupancy"
  echo "### conv_tc_kernel: block0 conv1 (1x1), block0 conv2 (3x3), block0 conv3 (1x1), block1 entry (3x3, N = 160)"
  python tools/ncu_summary.py $G/${SRC}_conv_tc.ncu-rep
  echo "### wgrad_tc_kernel: block1 entry, block0 conv3 (1x1), block0 conv2 (3x3), block0 conv1 (1x1)"
  python tools/ncu_summary.py $G/${SRC}_wgrad_tc.ncu-rep
  echo "### feature kernel, block-0 entry conv forward / backward (tcgen05)"
  python tools/ncu_summary.py $G/${SRC}_feat_conv0.ncu-rep
  for extra in bn_bwd elt_fwd; do
    if [ -f $G/${SRC}_${extra}.ncu-rep ]; then echo "### ${extra}"; python tools/ncu_summary.py $G/${SRC}_${extra}.ncu-rep; fi
  done
} > $P/${DST}_ncu_summary.txt
# raw per-launch metrics bench.py reads roofline.traffic from
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size
: > $P/${DST}_ncu_raw.csv
first=1
for r in conv_tc wgrad_tc feat_conv0; do
  if [ $first = 1 ]; then ncu -i $G/${SRC}_$r.ncu-rep --page raw --csv --metrics $M 2>/dev/null >> $P/${DST}_ncu_raw.csv; first=0
  else ncu -i $G/${SRC}_$r.ncu-rep --page raw --csv --metrics $M 2>/dev/null | tail -n +3 >> $P/${DST}_ncu_raw.csv; fi
done
for c in 2d 1d mixup sweep torch; do
  [ -s $G/${SRC}_bench_$c.json ] && tail -1 $G/${SRC}_bench_$c.json > $P/${DST}_bench_$c.json
done
cp $G/${SRC}_launches.csv $P/${DST}_launches.csv
# SASS evidence of the tcgen05 / TMEM / TMA path in the shipped library
SO=freesound-classification_b200/fsb200/libfsb200.so
{
  echo "# cuobjdump -sass $SO | grep -c <mnemonic>"
  cuobjdump -sass $SO > /tmp/fsb_sass.txt
  for m in UTCHMMA UTCBAR LDTM UTMALDG UTMASTG UBLKCP SYNCS; do echo "$m $(grep -c $m /tmp/fsb_sass.txt)"; done
  echo "# per kernel"
  awk '/Function :/ {fn=$3} /UTCHMMA|LDTM|UTMALDG|UTMASTG|UBLKCP/ {split($0,a," "); for(i in a) if (a[i] ~ /^(UTCHMMA|LDTM|UTMALDG|UTMASTG|UBLKCP)/) {sub(/\..*/,"",a[i]); c[fn" "a[i]]++}} END {for (k in c) print c[k], k}' /tmp/fsb_sass.txt | sort -k2 | c++filt 2>/dev/null | cut -c1-160
} > $P/${DST}_sass_counts.txt
ls -la $P | grep ${DST}_
