#!/bin/bash
# Runs on the GPU box (via gpurun): ncu captures of the step kernel's DRAM traffic per workload and batch size,
# steady state (launches 30..34 of a run of plain per-step launches, so the dirty L2 lines of the previous launch
# are written back inside the measured window).  CSVs land in gpurun_out/traffic_<workload>_<envs>.csv;
# scripts/ncu_traffic_update.py folds them into profiles/roofline_traffic.json.
# ncu runs only after the same command exited 0 without it (B200_PROFILING.md).
out=gpurun_out
mkdir -p $out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum
for spec in "c4 1048576" "c4 524288" "c4 262144" "c4 131072" "c4 65536" "c4 4096" "c5 262144" "c4_h5 1048576" "c4_pv2d 1048576" "c4_nopv 1048576"; do
  set -- $spec
  cmd="python bench.py --workload $1 --envs $2 --steps 40 --warmup 24 --graph-steps 0 --no-cpu --e2e-steps 1 --legs none"
  $cmd > $out/traffic_plain_$1_$2.log 2>&1 &&
  ncu --metrics $M --clock-control none -k "regex:step_(simple|lanes)" -s 30 -c 5 --csv --log-file $out/traffic_$1_$2.csv $cmd > $out/traffic_ncu_$1_$2.log 2>&1
  echo "$1 $2 rc=$?"
done
# launch list of the default bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/${TAG:-r4}_launches.csv \
    python bench.py --steps 48 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 --legs c2,c5,c3 > $out/${TAG:-r4}_ncu_launches.log 2>&1
echo "launch list rc=$?"
