#!/bin/bash
O=gpurun_out/s9; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 -s -k "tf32x3 or ma_ or pad_bn or mc_ or batchnorm or avgpool or long_sequence" > $O/pytest_new.log 2>&1; echo "pytest rc $?" >> $O/pytest_new.log
CVAD_PROFILE_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --profile-calls $O/calls_ma_train.md > $O/bench_ma_train.json 2> $O/bench_ma_train.err
CVAD_TF32X3=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_ma_train_notf32.json 2> $O/bench_ma_train_notf32.err
for R in 1 2 4; do for C in 4 8 16; do
  CVAD_BN_ROWS=$R CVAD_BN_CTAS=$C timeout 300 python tools/bw_probe.py 512 2>/dev/null | grep -E "apply|bwd|family" > $O/bw_R${R}_C${C}.md
  CVAD_BN_ROWS=$R CVAD_BN_CTAS=$C timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-check 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print('R$R C$C ms_per_step', round(d['ms_per_step'],4))" >> $O/bn_sweep.txt
done; done
for w in mc_infer mc_long; do
  timeout 400 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu --profile-calls $O/calls_$w.md > $O/bench_$w.json 2> $O/bench_$w.err
done
grep -h "tf32x3\]" $O/pytest_new.log; tail -n 3 $O/pytest_new.log; cat $O/bn_sweep.txt
