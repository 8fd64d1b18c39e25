#!/bin/bash
O=gpurun_out/s40; mkdir -p $O
for v in "k4:CVAD_BN_SMALL_GRID=4" "k8:CVAD_BN_SMALL_GRID=8" "k16:CVAD_BN_SMALL_GRID=16" "k4b:CVAD_BN_SMALL_GRID=4" "k8b:CVAD_BN_SMALL_GRID=8" "k16b:CVAD_BN_SMALL_GRID=16"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > $O/bench_$name.json 2> $O/bench_$name.err; python - $name "$envs" <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/s40/bench_{sys.argv[1]}.json')); print(sys.argv[1], sys.argv[2], d['ms_per_step'], d['value'], d['self_check']['ok'])
PY
done
