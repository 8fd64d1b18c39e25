#!/bin/bash
O=gpurun_out/s10; mkdir -p $O
timeout 1800 python -m pytest tests/test_flat_gpu.py tests/test_models_gpu.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --profile-calls $O/calls_ma_train.md > $O/bench_ma_train.json 2> $O/bench_ma_train.err
CVAD_STEM_FUSED_POOL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --profile-calls $O/calls_ma_train_nofusedpool.md > $O/bench_ma_train_nofusedpool.json 2> $O/bench_ma_train_nofusedpool.err
tail -n 3 $O/pytest.log; grep -h "stem\|maxpool" $O/calls_ma_train.md $O/calls_ma_train_nofusedpool.md
