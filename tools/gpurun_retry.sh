#!/bin/bash
# gpurun with retries while the pod answers "transient" (no box free): tools/gpurun_retry.sh [--gpus N] <timeout> <script>
gpus=""
if [ "$1" == "--gpus" ]; then gpus="--gpus $2"; shift 2; fi
t=$1; script=$2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $gpus --timeout $t -- "bash $script" > /tmp/gpurun_last.log 2>&1
  if grep -q "status=transient\|status=busy" /tmp/gpurun_last.log; then sleep 90; continue; fi
  break
done
tail -n 25 /tmp/gpurun_last.log
