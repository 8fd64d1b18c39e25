#!/bin/bash
O=gpurun_out/s14; mkdir -p $O
timeout 900 python -m pytest tests/test_flat_gpu.py -q -x --timeout 600 > $O/pytest_flat.log 2>&1; echo "pytest rc $?" >> $O/pytest_flat.log
tail -n 3 $O/pytest_flat.log
REPS=10 timeout 300 python tools/conv_probe.py 512 > $O/conv_layers.txt 2>&1; cat $O/conv_layers.txt
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s14/bench_n1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['per_kernel_ms'], d['self_check'])
PY
