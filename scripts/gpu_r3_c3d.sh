#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_rollout.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --legs c3,c3_fused,c3_sb3 --steps 48 --warmup 24 --e2e-steps 2 > $out/c3d.json 2>$out/c3d.err
python - <<PY
import json
try:
    d=json.loads(open('$out/c3d.json').read().strip().splitlines()[-1])
    for k,v in d['legs'].items():
        if isinstance(v,dict) and 'value' in v: print('%-12s us/step %8.3f value %.3e pol %.4f' % (k, v['ms_per_step']*1e3, v['value'], v.get('policy_forward_ms',0)))
except Exception as e: print('FAILED', e)
PY
