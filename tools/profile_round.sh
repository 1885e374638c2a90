#!/bin/bash
# Run on the GPU box (under gpurun): launch list of two canonical training steps + ncu --set full captures of the
# dominant kernels of the second (warm) step.  Outputs land in gpurun_out/.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches.csv python tools/one_step.py 2 64 > gpurun_out/${TAG}_launches.log 2>&1
FULL="$NCU --set full --import-source on"
# conv_tc_kernel: 38 launches per step (19 fwd + 19 dgrad); second step's first four = block0 conv1/conv2/conv3, block1 entry
$FULL -k regex:conv_tc_kernel -s 38 -c 4 -f -o gpurun_out/${TAG}_conv_tc python tools/one_step.py 2 64 > gpurun_out/${TAG}_conv_tc.log 2>&1
# wgrad_tc_kernel: 19 per step, backward order; index 15 = block1 entry, 16..18 = block0 conv3/conv2/conv1
FSB200_NO_OVERLAP=1 $FULL -k regex:wgrad_tc_kernel -s 34 -c 4 -f -o gpurun_out/${TAG}_wgrad_tc python tools/one_step.py 2 64 > gpurun_out/${TAG}_wgrad_tc.log 2>&1
# element-wise family (block 0 instances = the last launches of the step), feature kernel, block-0 direct conv
FSB200_NO_OVERLAP=1 $FULL -k regex:'bn_act_bwd_apply_kernel|bn_act_bwd_reduce_kernel' -s 98 -c 6 -f -o gpurun_out/${TAG}_bn_bwd python tools/one_step.py 2 64 > gpurun_out/${TAG}_bn_bwd.log 2>&1
$FULL -k regex:'bn_act_fwd_simple_kernel|bn_act_fwd_kernel' -s 24 -c 4 -f -o gpurun_out/${TAG}_elt_fwd python tools/one_step.py 2 64 > gpurun_out/${TAG}_elt_fwd.log 2>&1
$FULL -k regex:'feat_kernel|conv0_fwd_kernel|conv0_bwd_kernel' -s 3 -c 3 -f -o gpurun_out/${TAG}_feat_conv0 python tools/one_step.py 2 64 > gpurun_out/${TAG}_feat_conv0.log 2>&1
ls -la gpurun_out
