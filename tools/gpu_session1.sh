#!/bin/bash
# GPU session 1 of round 2: full parity suite + benches of all workloads on main, then the wip conv variants.
O=gpurun_out/s1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 600 > $O/pytest_main.log 2>&1; echo "pytest rc $?" >> $O/pytest_main.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-calls $O/calls_main.md > $O/bench_main.json 2> $O/bench_main.err
timeout 600 python bench.py --steps 20 --warmup 5 --frames f32 --no-cpu > $O/bench_main_f32frames.json 2> $O/bench_main_f32frames.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err
for w in mc_infer mb_train me_windows mc_long; do
  timeout 400 python bench.py --workload $w --steps 50 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 300 python bench.py --workload mb_train --batch 4 --steps 50 --warmup 5 --no-cpu > $O/bench_mb_train_b4.json 2> $O/bench_mb_train_b4.err
for v in kwstack kwstack1 dgrad; do
  CVAD_B200_LIB=$PWD/variants/$v.so timeout 900 python -m pytest tests/test_flat_gpu.py -q -s --timeout 300 -k "fwd_dgrad_wgrad or fused_bn" > $O/pytest_$v.log 2>&1; echo "pytest rc $?" >> $O/pytest_$v.log
  CVAD_B200_LIB=$PWD/variants/$v.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --profile-calls $O/calls_$v.md > $O/bench_$v.json 2> $O/bench_$v.err
done
timeout 600 python tools/stock_torch_b200.py > $O/stock_torch_b200.json 2> $O/stock_torch_b200.err
ls -la $O
tail -5 $O/pytest_main.log
