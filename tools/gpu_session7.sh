#!/bin/bash
# 2-GPU session: overlapped all-reduce test + 2-GPU benches (overlap on / off) + 1-GPU bench with the prefetching fused BN-bwd
O=gpurun_out/s7; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -s --timeout 800 > $O/pytest_parallel.log 2>&1; echo "pytest rc $?" >> $O/pytest_parallel.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err
CVAD_FUSED_BN_BWD=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $O/bench_n1_nobnfuse.json 2> $O/bench_n1_nobnfuse.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err
CVAD_ALLREDUCE_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_noovl.json 2> $O/bench_n2_noovl.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 --workload mb_train > $O/bench_mb_n2.json 2> $O/bench_mb_n2.err
tail -n 3 $O/pytest_parallel.log
