#!/bin/bash
O=gpurun_out/s33; mkdir -p $O
timeout 1200 python -m pytest tests/test_flat_gpu.py tests/test_models_gpu.py -q --timeout 600 -k "stem or pad_bn or ma_ or uint8" > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -n 3 $O/pytest.log
for f in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --profile-calls $O/calls_$f.md > $O/bench_$f.json 2> $O/bench_$f.err; python - $f <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/s33/bench_{sys.argv[1]}.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['self_check']['ok'], d['roofline']['frac'])
PY
done
grep stem8 $O/calls_1.md
