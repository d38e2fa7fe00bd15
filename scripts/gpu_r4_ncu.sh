#!/bin/bash
# r4: launch list of this repo's kernels of the bench command (the launch-floor leg's null kernels filtered out; ncu matches
# the kernel's base name), full ncu captures of the plane-major step kernel (C4), of the C5 kernel and of the one-lane-per-spot
# kernel (4,096 envs, rollout form), CTA-size sweep of the latter.  Each capture only after the same command has exited 0
# without ncu.  Summaries: scripts/launch_summary.py, scripts/ncu_summary.py -> profiles/r4_*.
out=gpurun_out; mkdir -p $out
L="python bench.py --steps 48 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 --legs c2,c5,c3"
timeout 300 $L > $out/r4_launches_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:step_|policy_|reset_kernel|gae|or_reduce" -c 1500 --csv --log-file $out/r4_launches.csv $L > $out/r4_ncu_launches.log 2>&1
echo "launch list rc=$?"
B="python bench.py --steps 24 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 --legs none"
for wl in c4 c5; do
  timeout 300 $B --workload $wl > $out/ncu_plain_$wl.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_simple -s 30 -c 1 -o $out/r4_step_${wl}_full -f $B --workload $wl > $out/ncu_$wl.log 2>&1
  tail -1 $out/ncu_$wl.log
done
S="python scripts/lanes_sweep.py --sizes 4096"
timeout 300 $S > $out/ncu_plain_lanes.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k step_lanes_kernel -s 20 -c 1 -o $out/r4_step_lanes_rollout_full -f $S > $out/ncu_lanes.log 2>&1
tail -1 $out/ncu_lanes.log
for w in 1 2 4; do echo "warps per CTA $w"; python scripts/lanes_sweep.py --sizes 2048,4096,8192 --warps $w 2>&1 | head -3; done
