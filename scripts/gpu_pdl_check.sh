#!/bin/bash
# PDL: correctness (rollout + parity tests) and effect on small batches and on C3
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for pdl in 0 1; do for E in 4096 65536 131072 1048576; do
  timeout 200 python bench.py --no-cpu --legs none --envs $E --steps 2400 --warmup 240 --e2e-steps 2 --pdl $pdl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pdl=$pdl E=$E: %.3f us/step  %.4g env-steps/s' % (d['ms_per_step']*1e3, d['value']))
"
done; done
timeout 300 python bench.py --no-cpu --legs c3,c3_sb3,c2 --steps 240 --warmup 24 --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d['legs'].items():
    if isinstance(v,dict): print(k, '%.4g'%v.get('value',0), 'ms %.5f'%v.get('ms_per_step',0), v.get('launch','')[:70], v.get('error',''))
"
