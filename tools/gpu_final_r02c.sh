#!/bin/bash
# light refresh of the round's final numbers on the last commit: full parity suite, smoke, both bench arms, per-call table
O=gpurun_out/r02c_final; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/smoke.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
CVAD_PROFILE_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --profile-calls $O/calls_in_graph.md > /dev/null 2> $O/calls.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python tools/profile_step.py 2 > $O/ncu_launches.log 2>&1
tail -n 3 $O/pytest_gpu.log; tail -2 $O/smoke.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_final/bench_n1.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'], d['roofline']['traffic'], d['self_check']['ok'], d['cpu_baseline'])
d=json.load(open('gpurun_out/r02c_final/bench_reference_arm.json')); print(d)
PY
