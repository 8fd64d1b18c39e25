#!/bin/bash
O=gpurun_out/s30; mkdir -p $O
timeout 900 python -m pytest tests/test_flat_gpu.py -q --timeout 600 -k "fwd_dgrad_wgrad" > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -n 4 $O/pytest.log
CVAD_FC_TUNE3=1 REPS=10 timeout 300 python tools/conv_probe.py 512 wgrad-staged > $O/wgrad_old.txt 2>&1
REPS=10 timeout 300 python tools/conv_probe.py 512 wgrad-staged > $O/wgrad_new.txt 2>&1
grep -o "^L[0-9] [^(]*\|wgrad-staged *[0-9.]* us\|totals.*" $O/wgrad_old.txt | paste -sd' '; grep -o "^L[0-9] [^(]*\|wgrad-staged *[0-9.]* us\|totals.*" $O/wgrad_new.txt | paste -sd' '
