#!/bin/bash
# one full ncu capture of the stem's pass 2 (conv + BN + ReLU + max-pool): why 460 us when the same MMAs take 194 us in pass 1
O=gpurun_out/s18; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stem_pool_kernel" -c 1 -o $O/stem_pool_full -f python tools/profile_step.py 1 > $O/ncu.log 2>&1
tail -3 $O/ncu.log; ls -la $O
