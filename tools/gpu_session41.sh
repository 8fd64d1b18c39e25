#!/bin/bash
O=gpurun_out/s41; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pad_bn_apply_relu" -c 2 -o $O/bn_apply -f python tools/profile_step.py 1 > $O/ncu1.log 2>&1
ls -la $O
