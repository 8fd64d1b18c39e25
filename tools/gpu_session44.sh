#!/bin/bash
O=gpurun_out/s44; mkdir -p $O
for w in mc_infer mb_train me_windows mc_long; do
  timeout 400 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err; python - $w <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/s44/bench_{sys.argv[1]}.json')); print(sys.argv[1], d['metric'], round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('self_check',{}).get('ok'))
PY
done
