#!/bin/bash
# where the convolution kernels' time goes outside the MMA loop: per-CTA entry / loop start / loop end / exit timestamps
O=gpurun_out/s12; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
FC_DEBUG=1 REPS=3 timeout 300 python tools/conv_probe.py 512 fwd > $O/conv_fwd_debug.txt 2>&1
FC_DEBUG=1 REPS=3 timeout 300 python tools/conv_probe.py 512 dgrad > $O/conv_dgrad_debug.txt 2>&1
cat $O/conv_fwd_debug.txt
