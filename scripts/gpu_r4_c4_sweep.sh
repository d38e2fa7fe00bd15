#!/bin/bash
# C4 step kernel on the plane-major layout: row staging and CTA shape re-checked
out=gpurun_out; mkdir -p $out
run() { timeout 300 python bench.py --no-cpu --legs none --steps 1200 --warmup 120 --e2e-steps 2 ${@:2} > $out/r4c4_$1.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('$out/r4c4_$1.json').read().strip().splitlines()[-1])
print('$1: ms/step %.5f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))
PY
}
run default
run bulk3 --bulk 3
run bulk0 --bulk 0
run w4 --warps 4
run w1 --warps 1
run w4_bulk3 --warps 4 --bulk 3
run default_again
