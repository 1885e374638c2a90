#!/bin/bash
mkdir -p gpurun_out
echo "== GPU suite (network + fullsize + round2)"
python -m pytest tests/test_gpu_network.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py -m gpu -q --tb=short > gpurun_out/r02_t16_full.log 2>&1; tail -8 gpurun_out/r02_t16_full.log
grep FULLSIZE gpurun_out/r02_t16_full.log | cut -c1-330
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_b8.json 2> gpurun_out/r02_b8.err; tail -c 300 gpurun_out/r02_b8.err
NCU="ncu --clock-control none"
FSB200_GRAPHS=0 $NCU --set full --import-source on -k regex:'conv0_tc_bwd_kernel|conv0_tc_fwd_kernel' -s 2 -c 2 -f -o gpurun_out/r02c_conv0 python tools/one_step.py 2 64 > gpurun_out/r02c_conv0.log 2>&1
tail -2 gpurun_out/r02c_conv0.log
