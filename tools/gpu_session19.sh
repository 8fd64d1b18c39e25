#!/bin/bash
O=gpurun_out/s19; mkdir -p $O
timeout 900 python -m pytest tests/test_flat_gpu.py tests/test_models_gpu.py -q --timeout 600 -k "stem or ma_ or uint8" > $O/pytest_stem.log 2>&1; echo "pytest rc $?" >> $O/pytest_stem.log
tail -n 4 $O/pytest_stem.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu --profile-calls $O/calls_in_graph.md > $O/bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s19/bench_n1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['self_check']['ok'], d['gpu_launches'])
PY
grep "stem" $O/calls_in_graph.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stem_pool_kernel" -c 1 -o $O/stem_pool_full -f python tools/profile_step.py 1 > $O/ncu.log 2>&1
