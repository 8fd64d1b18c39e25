#!/bin/bash
O=gpurun_out/s45; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -n 3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s45/bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['self_check']['ok'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python tools/profile_step.py 2 > $O/ncu_launches.log 2>&1
grep -c "at::" $O/launches.csv
