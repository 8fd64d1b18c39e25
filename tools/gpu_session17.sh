#!/bin/bash
O=gpurun_out/s17; mkdir -p $O
timeout 900 python -m pytest tests/test_flat_gpu.py tests/test_models_gpu.py -q --timeout 600 -k "pad_bn or stem or ma_ or graphed" > $O/pytest_bn.log 2>&1; echo "pytest rc $?" >> $O/pytest_bn.log
tail -n 4 $O/pytest_bn.log
for f in 0 1; do
CVAD_BN_FOLD=$f timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > $O/bench_fold$f.json 2> $O/bench_fold$f.err; python - $f <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/s17/bench_fold{sys.argv[1]}.json')); print('fold',sys.argv[1], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['self_check']['ok'], d['gpu_launches'])
PY
done
