#!/bin/bash
O=gpurun_out/s15; mkdir -p $O
timeout 900 python -m pytest tests/test_models_gpu.py -q --timeout 600 -k "ma0 or streaming" > $O/pytest_ma0.log 2>&1; echo "pytest rc $?" >> $O/pytest_ma0.log
tail -n 30 $O/pytest_ma0.log
