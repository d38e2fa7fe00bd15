#!/bin/bash
# ncu captures of the secondary kernels: the two-lane 64-spot step kernel (C5) and the fused policy kernel (C3).
out=gpurun_out; mkdir -p $out
timeout 300 python bench.py --workload c5 --steps 24 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 > $out/x_plain_c5.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 30 -c 1 -o $out/x_c5_step_full -f \
    python bench.py --workload c5 --steps 24 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 > $out/x_ncu_c5.log 2>&1
timeout 300 python scratch/pol_time.py > $out/x_plain_pol.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:policy_forward -s 10 -c 1 -o $out/x_policy_full -f \
    python scratch/pol_time.py > $out/x_ncu_pol.log 2>&1
echo done
