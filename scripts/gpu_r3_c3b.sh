#!/bin/bash
out=gpurun_out; mkdir -p $out
python scripts/policy_tc_trace.py 2>&1 | tail -4 | cut -c1-700
python scripts/policy_tc_prof.py 2>&1 | tail -3
for cv in none 100; do
  for m in off policy; do
    if [ $cv = none ]; then unset SNG_CARVEOUT; else export SNG_CARVEOUT=$cv; fi
    timeout 300 python bench.py --no-cpu --legs c3 --steps 48 --warmup 24 --e2e-steps 2 --rollout-pdl $m > $out/c3b_${cv}_$m.json 2>$out/c3b_${cv}_$m.err
    python - <<PY
import json
try:
    d=json.loads(open('$out/c3b_${cv}_$m.json').read().strip().splitlines()[-1])['legs']['c3']
    print('carveout $cv pdl $m: us/step %.2f value %.3e pol_ms %.4f' % (d['ms_per_step']*1e3, d['value'], d.get('policy_forward_ms',0)))
except Exception as e:
    print('$cv $m: FAILED', e)
PY
  done
done
