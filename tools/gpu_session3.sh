#!/bin/bash
O=gpurun_out/s3; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 600 > $O/pytest_main.log 2>&1; echo "pytest rc $?" >> $O/pytest_main.log
CVAD_PROFILE_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --profile-calls $O/calls_ma_train.md > $O/bench_ma_train.json 2> $O/bench_ma_train.err
CVAD_WGRAD_OVERLAP=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_ma_train_wgovl.json 2> $O/bench_ma_train_wgovl.err
for w in mc_infer mb_train mc_long me_windows; do
  timeout 400 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu --profile-calls $O/calls_$w.md > $O/bench_$w.json 2> $O/bench_$w.err
done
grep -h "passed\|failed" $O/pytest_main.log | tail -3
