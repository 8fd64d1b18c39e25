#!/bin/bash
# weak-scaling run on one 8-GPU box: N = 1, 2, 4, 8 back to back (bash tools/gpu_scaling.sh <tag>)
T=${1:-r02_scale}; O=gpurun_out/$T; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 30 --warmup 5 > $O/bench_n$n.json 2> $O/bench_n$n.err
  echo "rc $?" >> $O/bench_n$n.err
done
CVAD_ALLREDUCE_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 8 --steps 30 --warmup 5 > $O/bench_n8_noovl.json 2> $O/bench_n8_noovl.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 30 --warmup 5 --frames f32 > $O/bench_n8_f32frames.json 2> $O/bench_n8_f32frames.err
for w in mb_train mc_long mc_infer me_windows; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29730 bench.py --gpus 8 --steps 30 --warmup 5 --workload $w > $O/bench_${w}_n8.json 2> $O/bench_${w}_n8.err
done
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -s --timeout 800 > $O/pytest_parallel.log 2>&1; echo "pytest rc $?" >> $O/pytest_parallel.log
python - <<'PY'
import json, glob, os
O = os.environ.get("O", "") or "gpurun_out/r02_scale"
PY
for f in $O/bench_n*.json $O/bench_*_n8.json; do python -c "
import json,sys
try:
    d=json.load(open('$f')); print('$f', d['n_gpus'], round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('$f', 'ERR', e)
"; done
tail -n 3 $O/pytest_parallel.log
