#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 240 python -m pytest tests/test_gpu_rollout.py -m gpu -x -q -k "fused_policy_step" > $out/fused_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $out/fused_pytest.log
timeout 200 python bench.py --no-cpu --legs c3,c3_fused --steps 48 --warmup 24 --e2e-steps 2 > $out/fused_legs.json 2>$out/fused_legs.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('$out/fused_legs.json').read().strip().splitlines()[-1])
    for k,v in d['legs'].items():
        if isinstance(v,dict) and 'value' in v: print('%-24s us/step %8.3f value %.3e' % (k, v['ms_per_step']*1e3, v['value']))
        elif isinstance(v,dict): print(k, v)
except Exception as e: print('FAILED', e)
PY
tail -5 $out/fused_legs.err
