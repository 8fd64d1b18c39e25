#!/bin/bash
O=gpurun_out/s37; mkdir -p $O
timeout 900 python -m pytest tests/test_flat_gpu.py -q --timeout 600 -k "pad_bn" > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -n 3 $O/pytest.log
for v in "occ2:CVAD_BN_OCC=2" "occ3:CVAD_BN_OCC=3" "occ2b:CVAD_BN_OCC=2" "occ3b:CVAD_BN_OCC=3"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > $O/bench_$name.json 2> $O/bench_$name.err; python - $name "$envs" <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/s37/bench_{sys.argv[1]}.json')); print(sys.argv[1], sys.argv[2], d['ms_per_step'], d['value'], d['self_check']['ok'])
PY
done
