#!/bin/bash
# ncu full capture (source counters) of the C4 step kernel, steady state, after the same command exited 0 without ncu.
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 24 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 --legs none"
timeout 300 $B > $out/r3_plain_c4.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 30 -c 1 -o $out/r3_step_c4_full -f $B > $out/r3_ncu_c4.log 2>&1
tail -1 $out/r3_ncu_c4.log
timeout 300 python bench.py --no-cpu --legs none --steps 1200 --warmup 120 --e2e-steps 2 > $out/r3_c4_base.json 2>$out/r3_c4_base.err
cat $out/r3_c4_base.json | cut -c1-400
echo done
