#!/bin/bash
# One full ncu capture of the step kernel (after a plain run of the same command exits 0).
# Usage: scripts/gpu_ncu.sh <tag> [bench args]
tag=${1:-p}; shift
out=gpurun_out; mkdir -p $out
timeout 300 python bench.py --steps 24 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 "$@" > $out/${tag}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 30 -c 1 -o $out/${tag}_step_full -f \
    python bench.py --steps 24 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 "$@" > $out/${tag}_ncu_full.log 2>&1
tail -2 $out/${tag}_ncu_full.log
