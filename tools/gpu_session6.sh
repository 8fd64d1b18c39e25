#!/bin/bash
O=gpurun_out/s6; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -s --timeout 600 -k "mlp_chain or fused_bn_backward" > $O/pytest_new.log 2>&1; echo "pytest rc $?" >> $O/pytest_new.log
timeout 1800 python -m pytest tests/test_models_gpu.py -m gpu -q -s --timeout 600 -k "ma_" > $O/pytest_models.log 2>&1; echo "pytest rc $?" >> $O/pytest_models.log
CVAD_PROFILE_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --profile-calls $O/calls_ma_train.md > $O/bench_ma_train.json 2> $O/bench_ma_train.err
CVAD_FUSED_CHAINS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_ma_train_nochain.json 2> $O/bench_ma_train_nochain.err
CVAD_FUSED_BN_BWD=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_ma_train_nobnfuse.json 2> $O/bench_ma_train_nobnfuse.err
tail -n 3 $O/pytest_new.log $O/pytest_models.log
