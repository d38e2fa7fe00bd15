#!/bin/bash
# Round-4 record: the driver's own sequence on a fresh box (timed), the steady-state DRAM traffic captures, the launch list,
# the wide parity sweep and soak on the final kernels, the batch-size sweep of the two step kernels.
out=gpurun_out; mkdir -p $out
s=$(date +%s)
timeout 900 python bench.py > $out/r4_bench_1gpu.json 2> $out/r4_bench_1gpu.err; echo "bench rc=$? in $(( $(date +%s) - s )) s"
s=$(date +%s)
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $out/r4_bench_reference_arm.json 2>> $out/r4_bench_1gpu.err; echo "reference arm rc=$? in $(( $(date +%s) - s )) s"
cut -c1-300 $out/r4_bench_reference_arm.json
TAG=r4 bash scripts/ncu_traffic.sh
timeout 600 python scripts/parity_sweep.py > $out/r4_parity_sweep.log 2>&1; echo "sweep rc=$?"; tail -1 $out/r4_parity_sweep.log
timeout 600 python scripts/parity_soak.py > $out/r4_parity_soak.log 2>&1; echo "soak rc=$?"; tail -2 $out/r4_parity_soak.log
timeout 300 python scripts/lanes_sweep.py > $out/r4_lanes_sweep.log 2>&1; echo "lanes sweep rc=$?"
