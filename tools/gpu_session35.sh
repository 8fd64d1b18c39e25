#!/bin/bash
O=gpurun_out/s35; mkdir -p $O
timeout 1200 python -m pytest tests/test_flat_gpu.py tests/test_models_gpu.py -q --timeout 600 -k "stem or ma_ or uint8" > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -n 3 $O/pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --profile-calls $O/calls.md > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s35/bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['self_check'], d['roofline']['frac'])
PY
grep stem8 $O/calls.md
