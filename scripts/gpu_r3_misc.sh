#!/bin/bash
out=gpurun_out; mkdir -p $out
for w in 2 4 1; do
  timeout 300 python bench.py --no-cpu --legs none --steps 1200 --warmup 120 --e2e-steps 2 --warps $w > $out/warps_$w.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('$out/warps_$w.json').read().strip().splitlines()[-1])
print('warps/CTA $w: ms/step %.5f frac %.3f' % (d['ms_per_step'], d['roofline']['frac']))
PY
done
