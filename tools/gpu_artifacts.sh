#!/bin/bash
# Measurement artefacts of the current build on one B200 (copied into profiles/ afterwards): full parity suite, every bench line, the
# per-call tables, the ncu launch list of one step, the convolution family's DRAM traffic, full ncu captures of the top kernels, the
# bandwidth table.  Usage: bash tools/gpu_artifacts.sh <tag>
T=${1:-r02a}; O=gpurun_out/$T; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/smoke.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
CVAD_PROFILE_SHAPES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --profile-calls $O/calls_in_graph.md > /dev/null 2> $O/calls.err
for w in mc_infer mb_train me_windows mc_long; do
  timeout 400 python bench.py --workload $w --steps 50 --warmup 5 --profile-calls $O/calls_$w.md > $O/bench_$w.json 2> $O/bench_$w.err
  timeout 400 python bench.py --workload $w --impl reference --steps 20 --warmup 3 > $O/bench_${w}_reference_arm.json 2> $O/bench_${w}_ref.err
done
timeout 300 python bench.py --workload mb_train --batch 4 --steps 50 --warmup 5 --no-cpu > $O/bench_mb_train_b4.json 2> /dev/null
timeout 300 python tools/bw_probe.py 512 > $O/bandwidth_kernels.md 2> $O/bw.err
REPS=5 timeout 300 python tools/conv_probe.py 512 > $O/conv_layers.txt 2> $O/conv_layers.err
timeout 600 python tools/stock_torch_b200.py > $O/stock_torch_b200.json 2> $O/stock.err
# ---- ncu (numbers printed by these runs are never bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python tools/profile_step.py 2 > $O/ncu_launches.log 2>&1
STEP_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/conv_step_metrics.csv python tools/conv_probe.py 512 > $O/ncu_conv.log 2>&1
STEP_ONLY=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"flatconv_kernel|flatwgrad_kernel" -c 3 -o $O/flatconv_L0_full -f python tools/conv_probe.py 512 > $O/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pad_bn_relu_bwd_apply|pad_bn_apply_relu|stem_tf32_kernel|maxpool3x3s2" -c 6 -o $O/bandwidth_full -f python tools/profile_step.py 1 > $O/ncu_full_bw.log 2>&1
ls -la $O | head -60; tail -n 3 $O/pytest_gpu.log; cat $O/smoke.log | tail -3
