#!/bin/bash
O=gpurun_out/s32; mkdir -p $O
timeout 1200 python -m pytest tests/test_models_gpu.py -q --timeout 600 -k "ma_ or graphed" > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -n 3 $O/pytest.log
for f in 0 1 0 1; do
CVAD_BN_FOLD=$f timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > $O/bench_fold$f.json 2> $O/bench_fold$f.err; python - $f <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/s32/bench_fold{sys.argv[1]}.json')); print('fold',sys.argv[1], d['ms_per_step'], d['value'], d['e2e']['value'], d['self_check']['ok'], d['roofline']['traffic'])
PY
done
