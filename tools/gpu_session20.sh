#!/bin/bash
# source-level ncu captures of the non-convolution kernels of one step (looking for bank conflicts / serialised loads)
O=gpurun_out/s20; mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"pad_bn_relu_bwd_apply|pad_reduce|pad_bn_apply_relu|flatwgrad|avgpool|stem_s2d_u8|adam_flat|wgrad_fold|colsum|sumsq" -o $O/batch -f python tools/profile_step.py 1 > $O/ncu.log 2>&1
tail -3 $O/ncu.log; ls -la $O
