#!/bin/bash
# C3 rollout leg with the programmatic-dependent-launch modes of the rollout loop
out=gpurun_out; mkdir -p $out; tag=${1:-c3}
for m in off policy policy+x both both+x step+x; do
  timeout 300 python bench.py --no-cpu --legs c3 --steps 48 --warmup 24 --e2e-steps 2 --rollout-pdl $m > $out/${tag}_$m.json 2>$out/${tag}_$m.err
  python - <<PY
import json
try:
    d=json.loads(open('$out/${tag}_$m.json').read().strip().splitlines()[-1])['legs']['c3']
    print('$m: us/step %.2f value %.3e pol_ms %.4f rew %.4f' % (d['ms_per_step']*1e3, d['value'], d.get('policy_forward_ms',0), d['mean_step_reward']))
except Exception as e:
    print('$m: FAILED', e)
PY
done
