#!/bin/bash
# flat convolution pipeline variants: merged-plane stride-2 data-gradient, third segment stage, double-buffered accumulators at N = 128
O=gpurun_out/s13; mkdir -p $O
timeout 900 python -m pytest tests/test_flat_gpu.py -q -x --timeout 600 > $O/pytest_flat.log 2>&1; echo "pytest rc $?" >> $O/pytest_flat.log
tail -n 5 $O/pytest_flat.log
for v in "base:CVAD_DGRAD_MODE=1 CVAD_FC_TUNE0=4 CVAD_FC_TUNE1=2" "new:" "nosrc3:CVAD_FC_TUNE1=2" "sub4:CVAD_FC_TUNE0=4" "dg1:CVAD_DGRAD_MODE=1"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs FC_DEBUG=1 REPS=5 timeout 300 python tools/conv_probe.py 512 fwd > $O/fwd_$name.txt 2>&1
  env $envs FC_DEBUG=1 REPS=5 timeout 300 python tools/conv_probe.py 512 dgrad > $O/dgrad_$name.txt 2>&1
  echo "== $name"; grep -o "^L[0-9].*GFLOP)\|fwd  *[0-9.]* us\|dgrad  *[0-9.]* us\|totals.*" $O/fwd_$name.txt $O/dgrad_$name.txt | paste -sd' ' | fold -w 400
done
