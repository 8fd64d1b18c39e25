#!/bin/bash
O=gpurun_out/s26; mkdir -p $O
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_models_gpu.py -q --timeout 600 -k "ma_ or ma0 or streaming or graphed or mean_mid or softmax or kernels" > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -n 4 $O/pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --profile-calls $O/calls_in_graph.md > $O/bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s26/bench_n1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['self_check']['ok'], d['gpu_launches'])
PY
grep "stem\|avgpool" $O/calls_in_graph.md
