#!/bin/bash
# C4 step-kernel sweep over register caps (library builds in scratch/) and warps per CTA
out=gpurun_out; mkdir -p $out
run() { SNG_LIB_PATH=$2 timeout 300 python bench.py --no-cpu --legs none --steps 1200 --warmup 120 --e2e-steps 2 ${@:3} > $out/c4_$1.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('$out/c4_$1.json').read().strip().splitlines()[-1])
print('$1: ms/step %.5f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))
PY
}
D=smart_nanogrid_gym_b200/libsng.so
run b8 $D
run b7 scratch/libsng_b7.so
run b6 scratch/libsng_b6.so
run b8_w4 $D --warps 4
run b7_w4 scratch/libsng_b7.so --warps 4
run b8_again $D
