#!/bin/bash
out=gpurun_out; mkdir -p $out
for sets in 1 2; do
  export SNG_POLICY_IO_SETS=$sets
  timeout 300 python bench.py --no-cpu --legs c3,c3_sb3 --steps 48 --warmup 24 --e2e-steps 2 > $out/c3c_$sets.json 2>$out/c3c_$sets.err
  python - <<PY
import json
try:
    d=json.loads(open('$out/c3c_$sets.json').read().strip().splitlines()[-1])
    for k,v in d['legs'].items():
        if isinstance(v,dict) and 'value' in v: print('io sets $sets %-12s us/step %8.3f value %.3e' % (k, v['ms_per_step']*1e3, v['value']))
except Exception as e: print('FAILED', e)
PY
done
SNG_POLICY_IO_SETS=2 timeout 300 python -m pytest tests/test_gpu_rollout.py -m gpu -x -q 2>&1 | tail -3
