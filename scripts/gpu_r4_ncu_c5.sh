#!/bin/bash
# r4: full ncu capture of the C5 step kernel (64 spots, four lanes per env) on the plane-major state.
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 24 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 --legs none --workload c5"
timeout 300 $B > $out/ncu_plain_c5.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_simple -s 30 -c 1 -o $out/r4_step_c5_full -f $B > $out/ncu_c5.log 2>&1
tail -1 $out/ncu_c5.log
