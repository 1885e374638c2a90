#!/bin/bash
# Round-2 record run: full GPU suite, bench lines of the four configs, launch list, ncu --set full of the dominant kernels.
mkdir -p gpurun_out
TAG=${1:-r02p}
echo "== GPU suite"
python -m pytest tests -m gpu -q --tb=short > gpurun_out/${TAG}_tests.log 2>&1; tail -4 gpurun_out/${TAG}_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_2d.json 2> gpurun_out/${TAG}_bench_2d.err; tail -c 300 gpurun_out/${TAG}_bench_2d.err
python bench.py --config 1d --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_1d.json 2> gpurun_out/${TAG}_bench_1d.err
python bench.py --config mixup_dp --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_mixup.json 2> gpurun_out/${TAG}_bench_mixup.err
python bench.py --config infer_sweep --clips 2000 --no-cpu-baseline > gpurun_out/${TAG}_bench_sweep.json 2> gpurun_out/${TAG}_bench_sweep.err
python bench.py --impl torch_gpu --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_torch.json 2> gpurun_out/${TAG}_bench_torch.err; tail -c 300 gpurun_out/${TAG}_bench_torch.err
NCU="ncu --clock-control none"
FSB200_GRAPHS=0 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches.csv python tools/one_step.py 2 64 > gpurun_out/${TAG}_launches.log 2>&1
FULL="$NCU --set full --import-source on"
export FSB200_GRAPHS=0 FSB200_NO_OVERLAP=1
$FULL -k regex:conv_tc_kernel -s 38 -c 4 -f -o gpurun_out/${TAG}_conv_tc python tools/one_step.py 2 64 > gpurun_out/${TAG}_conv_tc.log 2>&1
$FULL -k regex:wgrad_tc_kernel -s 34 -c 4 -f -o gpurun_out/${TAG}_wgrad_tc python tools/one_step.py 2 64 > gpurun_out/${TAG}_wgrad_tc.log 2>&1
$FULL -k regex:'feat2048_mel_kernel|feat_kernel|conv0_tc_fwd_kernel|conv0_tc_bwd_kernel' -s 3 -c 3 -f -o gpurun_out/${TAG}_feat_conv0 python tools/one_step.py 2 64 > gpurun_out/${TAG}_feat_conv0.log 2>&1
tail -2 gpurun_out/${TAG}_feat_conv0.log
# compact BatchNorm-backward kernels of block 0 (the last launches of the step)
# (gpurun merges at most 64 MiB back: the element-wise captures go without source import)
LIGHT="$NCU --set full"
$LIGHT -k regex:'bn_bwd_c8_reduce_kernel|bn_bwd_c8_apply_kernel|bn_bwd_c8res_reduce_kernel|bn_bwd_c8res_apply_kernel' -s 86 -c 6 -f -o gpurun_out/${TAG}_bn_bwd python tools/one_step.py 2 64 > gpurun_out/${TAG}_bn_bwd.log 2>&1
$LIGHT -k regex:'bn_act_fwd_simple_kernel|bn_act_fwd_kernel' -s 26 -c 2 -f -o gpurun_out/${TAG}_elt_fwd python tools/one_step.py 2 64 > gpurun_out/${TAG}_elt_fwd.log 2>&1
du -sh gpurun_out
