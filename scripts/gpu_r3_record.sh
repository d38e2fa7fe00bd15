#!/bin/bash
# The driver's own sequence on a fresh box (timed), then the steady-state DRAM traffic captures and the launch list.
out=gpurun_out; mkdir -p $out
s=$(date +%s)
timeout 900 python bench.py > $out/r3_bench_1gpu.json 2> $out/r3_bench_1gpu.err; echo "bench rc=$? in $(( $(date +%s) - s )) s"
s=$(date +%s)
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $out/r3_bench_reference_arm.json 2>> $out/r3_bench_1gpu.err; echo "reference arm rc=$? in $(( $(date +%s) - s )) s"
cut -c1-400 $out/r3_bench_reference_arm.json
bash scripts/ncu_traffic.sh
