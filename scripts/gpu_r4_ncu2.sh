#!/bin/bash
# r4: launch list of this repo's kernels of the bench command (the launch-floor leg's null kernels filtered out) and a full
# ncu capture of the one-lane-per-spot kernel (4,096 envs, rollout form).  Each only after the same command exited 0 without ncu.
out=gpurun_out; mkdir -p $out
L="python bench.py --steps 48 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 --legs c2,c5,c3"
timeout 300 $L > $out/r4_launches_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:step_|policy_|reset_kernel|gae|or_reduce" -c 1500 --csv --log-file $out/r4_launches.csv $L > $out/r4_ncu_launches.log 2>&1
echo "launch list rc=$?"
S="python scripts/lanes_sweep.py --sizes 4096"
timeout 300 $S > $out/ncu_plain_lanes.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k step_lanes_kernel -s 20 -c 1 -o $out/r4_step_lanes_rollout_full -f $S > $out/ncu_lanes.log 2>&1
tail -1 $out/ncu_lanes.log
