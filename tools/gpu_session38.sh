#!/bin/bash
# ncu --set full of the BatchNorm backward kernels of one step: the 8x12x256 layer (first in the backward) and the 60x90x32 layer (last)
O=gpurun_out/s38; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pad_bn_relu_bwd_apply|pad_reduce" -c 2 -o $O/bn_bwd_small -f python tools/profile_step.py 1 > $O/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pad_bn_relu_bwd_apply|pad_reduce" -s 14 -c 2 -o $O/bn_bwd_large -f python tools/profile_step.py 1 > $O/ncu2.log 2>&1
ls -la $O; tail -2 $O/ncu2.log
