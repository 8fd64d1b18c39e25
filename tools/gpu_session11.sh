#!/bin/bash
O=gpurun_out/s11; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --profile-calls $O/calls_ma_train.md > $O/bench_ma_train.json 2> $O/bench_ma_train.err
CVAD_STEM_FUSED_POOL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_ma_train_nofusedpool.json 2> $O/bench_ma_train_nofusedpool.err
REPS=5 timeout 300 python tools/conv_probe.py 512 > $O/conv_layers.txt 2>&1
tail -n 3 $O/pytest.log; tail -2 $O/conv_layers.txt
