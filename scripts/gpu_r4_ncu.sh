#!/bin/bash
# r4: launch list of the sng:: kernels of the bench command (the launch-floor leg's null kernels filtered out), full ncu
# captures of the plane-major step kernel (C4) and of the one-lane-per-spot kernel (4,096 envs, rollout form), CTA-size sweep
# of the latter.  Each capture only after the same command has exited 0 without ncu.
out=gpurun_out; mkdir -p $out
L="python bench.py --steps 48 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 --legs c2,c5,c3"
timeout 300 $L > $out/r4_launches_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sng:: -c 1500 --csv --log-file $out/r4_launches.csv $L > $out/r4_ncu_launches.log 2>&1
echo "launch list rc=$?"
B="python bench.py --steps 24 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 --legs none"
timeout 300 $B > $out/ncu_plain_c4.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_simple -s 30 -c 1 -o $out/r4_step_c4_full -f $B > $out/ncu_c4.log 2>&1
tail -1 $out/ncu_c4.log
S="python scripts/lanes_sweep.py --sizes 4096"
timeout 300 $S > $out/ncu_plain_lanes.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_lanes_kernel.*true -s 20 -c 1 -o $out/r4_step_lanes_rollout_full -f $S > $out/ncu_lanes.log 2>&1
tail -1 $out/ncu_lanes.log
for w in 1 2 4; do echo "warps per CTA $w"; python scripts/lanes_sweep.py --sizes 2048,4096,8192 --warps $w 2>&1 | head -3; done
