#!/bin/bash
O=gpurun_out/s21; mkdir -p $O
timeout 600 python -m pytest tests/test_flat_gpu.py -q -x --timeout 300 -k "stem_f16" > $O/pytest_stem8.log 2>&1; echo "pytest rc $?" >> $O/pytest_stem8.log
grep "stem f16\|passed\|failed\|rc\|Error\|error" $O/pytest_stem8.log | head -40
