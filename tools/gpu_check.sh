#!/bin/bash
# On the GPU box: full GPU suite, bench lines of every config, launch list + ncu captures; everything lands in gpurun_out/.
mkdir -p gpurun_out
echo "== full GPU suite, default switches"
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_t15_full.log 2>&1; tail -12 gpurun_out/r02_t15_full.log
for g in 0 1; do
  FSB200_GRAPHS=$g python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_b7_g$g.json 2> gpurun_out/r02_b7_g$g.err
  tail -c 200 gpurun_out/r02_b7_g$g.err
done
python bench.py --config mixup_dp --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_b7_mixup.json 2> gpurun_out/r02_b7_mixup.err; tail -c 200 gpurun_out/r02_b7_mixup.err
NCU="ncu --clock-control none"
FSB200_GRAPHS=0 $NCU --set full --import-source on -k regex:'conv0_bwd_kernel|conv0_tc_fwd_kernel' -s 2 -c 2 -f -o gpurun_out/r02b_conv0 python tools/one_step.py 2 64 > gpurun_out/r02b_conv0.log 2>&1
tail -2 gpurun_out/r02b_conv0.log
