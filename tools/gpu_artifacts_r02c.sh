#!/bin/bash
# Round-2 (second half) measurement artefacts on one B200, copied into profiles/r02c_* afterwards.  Outputs stay small (< 64 MiB pull limit).
T=${1:-r02c}; O=gpurun_out/$T; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/smoke.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
CVAD_PROFILE_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --profile-calls $O/calls_in_graph.md > /dev/null 2> $O/calls.err
for w in mc_infer mb_train me_windows mc_long; do
  timeout 400 python bench.py --workload $w --steps 50 --warmup 5 > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 300 python tools/bw_probe.py 512 > $O/bandwidth_kernels.md 2> $O/bw.err
REPS=5 timeout 300 python tools/conv_probe.py 512 > $O/conv_layers.txt 2> $O/conv_layers.err
# ---- ncu (numbers printed by these runs are never bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python tools/profile_step.py 2 > $O/ncu_launches.log 2>&1
STEP_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/conv_step_metrics.csv python tools/conv_probe.py 512 > $O/ncu_conv.log 2>&1
STEP_ONLY=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"flatconv_kernel|flatwgrad_kernel" -c 3 -o $O/flatconv_L0_full -f python tools/conv_probe.py 512 > $O/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stem8_stats_kernel|stem8_pool_kernel" -c 2 -o $O/stem8_full -f python tools/profile_step.py 1 > $O/ncu_full_stem.log 2>&1
du -sh $O; ls -la $O | head -60; tail -n 3 $O/pytest_gpu.log; cat $O/smoke.log | tail -3
