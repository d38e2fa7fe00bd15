#!/bin/bash
# Full ncu captures (source counters on) of the step kernel at C4 and C5 and of the tensor-core policy kernel.
# Each capture only after the same command has exited 0 without ncu.
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 24 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 --legs none"
timeout 300 $B > $out/ncu_plain_c4.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 30 -c 1 -o $out/r3_step_c4_full -f $B > $out/ncu_c4.log 2>&1
tail -1 $out/ncu_c4.log
timeout 300 $B --workload c5 > $out/ncu_plain_c5.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 30 -c 1 -o $out/r3_step_c5_full -f $B --workload c5 > $out/ncu_c5.log 2>&1
tail -1 $out/ncu_c5.log
timeout 300 python scripts/policy_tc_prof.py > $out/ncu_plain_tc.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:policy_tc -s 4 -c 1 -o $out/r3_policy_tc -f python scripts/policy_tc_prof.py > $out/ncu_tc.log 2>&1
tail -1 $out/ncu_tc.log
echo done
