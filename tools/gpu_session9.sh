#!/bin/bash
O=gpurun_out/s9; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 -s -k "tf32x3 or ma_" > $O/pytest_new.log 2>&1; echo "pytest rc $?" >> $O/pytest_new.log
CVAD_PROFILE_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --profile-calls $O/calls_ma_train.md > $O/bench_ma_train.json 2> $O/bench_ma_train.err
CVAD_TF32X3=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_ma_train_notf32.json 2> $O/bench_ma_train_notf32.err
grep -h "tf32x3\]" $O/pytest_new.log; tail -n 3 $O/pytest_new.log
