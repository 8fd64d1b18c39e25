#!/bin/bash
# two GPUs: the data-parallel tests (exchange modes, full-state broadcast) and a short weak-scaling bench with the parallel tail branches
O=gpurun_out/s31; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
timeout 900 python -m pytest tests/test_parallel_gpu.py -q --timeout 600 > $O/pytest_parallel.log 2>&1; echo "pytest rc $?" >> $O/pytest_parallel.log
tail -n 5 $O/pytest_parallel.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/s31/bench_n2.json')); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['self_check']['ok'])
PY
tail -3 $O/bench_n2.err
