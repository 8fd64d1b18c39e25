#!/bin/bash
O=gpurun_out/s36; mkdir -p $O
for v in "a:" "b:CVAD_BN_CTAS=16" "c:CVAD_BN_CTAS=4" "d:CVAD_BN_ROWS=4" "e:CVAD_BN_ROWS=4 CVAD_BN_CTAS=4" "a2:"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > $O/bench_$name.json 2> $O/bench_$name.err; python - $name "$envs" <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/s36/bench_{sys.argv[1]}.json')); print(sys.argv[1], sys.argv[2], d['ms_per_step'], d['value'], d['self_check']['ok'])
PY
done
