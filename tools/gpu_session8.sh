#!/bin/bash
O=gpurun_out/s8; mkdir -p $O
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -q -s --timeout 800 > $O/pytest_parallel.log 2>&1; echo "pytest rc $?" >> $O/pytest_parallel.log
CVAD_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc $?" >> $O/bench_n2.err
CVAD_ALLREDUCE_OVERLAP=0 CVAD_BENCH_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2_noovl.json 2> $O/bench_n2_noovl.err; echo "rc $?" >> $O/bench_n2_noovl.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 --workload mb_train > $O/bench_mb_n2.json 2> $O/bench_mb_n2.err; echo "rc $?" >> $O/bench_mb_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 2 --steps 20 --warmup 5 --workload mc_long > $O/bench_mclong_n2.json 2> $O/bench_mclong_n2.err; echo "rc $?" >> $O/bench_mclong_n2.err
tail -n 3 $O/pytest_parallel.log; tail -n 5 $O/bench_n2.err
