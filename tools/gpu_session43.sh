#!/bin/bash
# stability of the order-sensitive comparisons: the graph / uint8 / trajectory tests repeated, then the whole suite twice
O=gpurun_out/s43; mkdir -p $O
for i in 1 2 3 4 5 6; do
  timeout 600 python -m pytest tests/test_models_gpu.py -q --timeout 600 -k "graphed or uint8 or trajectory or ma_train or ma0_train" > $O/rep_$i.log 2>&1; echo "rep $i rc $?" ; tail -n 1 $O/rep_$i.log
done
for i in 1 2; do timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/full_$i.log 2>&1; echo "full $i rc $?"; tail -n 1 $O/full_$i.log; done
grep -h "\[graph\] mean" $O/rep_*.log | head -12
