#!/bin/bash
# C5 (64 spots) kernel sweep: lanes per env (variant 0 = 4 lanes, 3 = 2 lanes, 2 = 1 lane) x register caps (library builds in scratch/)
out=gpurun_out; mkdir -p $out
run() { # name libpath extra-args
  SNG_LIB_PATH=$2 timeout 300 python bench.py --no-cpu --legs none --workload c5 --steps 480 --warmup 48 --e2e-steps 2 ${@:3} > $out/c5_$1.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('$out/c5_$1.json').read().strip().splitlines()[-1])
print('$1: ms/step %.5f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))
PY
}
D=smart_nanogrid_gym_b200/libsng.so
run l4_m6 $D
run l2 $D --variant 3
run l1 $D --variant 2
run l4_m8 scratch/libsng_m8.so
run l4_m5 scratch/libsng_m5.so
run l4_m4 scratch/libsng_m4.so
run l4_m6_w1 $D --warps 1
run l4_m6_w4 $D --warps 4
run l4_m8_w4 scratch/libsng_m8.so --warps 4
