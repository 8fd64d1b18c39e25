#!/bin/bash
O=gpurun_out/s16; mkdir -p $O
timeout 900 python -m pytest tests/test_flat_gpu.py -q --timeout 600 -k "pad_bn or stem" > $O/pytest_bn.log 2>&1; echo "pytest rc $?" >> $O/pytest_bn.log
tail -n 4 $O/pytest_bn.log
CVAD_BN_REDUCE_ROWS=0 timeout 300 python tools/bw_probe.py 512 > $O/bw_flat.md 2> $O/bw_flat.err
timeout 300 python tools/bw_probe.py 512 > $O/bw_rows.md 2> $O/bw_rows.err
grep "pad_bn_stats\|pad_bn_relu_bwd\|BatchNorm family" $O/bw_flat.md; echo; grep "pad_bn_stats\|pad_bn_relu_bwd\|BatchNorm family" $O/bw_rows.md
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/s16/bench_n1.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['self_check']['ok'])
PY
