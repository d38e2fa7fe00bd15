#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, bench, ncu launch list, one full ncu capture.
# Usage: scripts/gpu_check.sh [tag]   -> everything lands in gpurun_out/<tag>_*
tag=${1:-run}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cat $out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 48 --warmup 3 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 48 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 > $out/${tag}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_ -s 30 -c 1 -o $out/${tag}_step_full -f \
    python bench.py --steps 24 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 > $out/${tag}_ncu_full.log 2>&1
echo done
