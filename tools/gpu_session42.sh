#!/bin/bash
O=gpurun_out/s42; mkdir -p $O
timeout 900 python -m pytest tests/test_flat_gpu.py tests/test_models_gpu.py -q --timeout 600 -k "pad_bn or ma_c2 or ma_train or ma_bf16" > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -n 3 $O/pytest.log
for i in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > $O/bench_$i.json 2> $O/bench_$i.err; python - $i <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/s42/bench_{sys.argv[1]}.json')); print(d['ms_per_step'], d['value'], d['self_check']['ok'])
PY
done
