#!/bin/bash
# e2e (sng_step_host): small first chunks on/off x host chunks
out=gpurun_out; mkdir -p $out
for r in 1 0 1 0; do for hc in 8; do
  SNG_HOST_RAMP=$r timeout 300 python bench.py --no-cpu --legs none --steps 240 --warmup 24 --e2e-steps 48 --host-chunks $hc > $out/e2e_r${r}_hc${hc}.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('$out/e2e_r${r}_hc${hc}.json').read().strip().splitlines()[-1])
e=d['e2e']; print('ramp=$r chunks=$hc e2e %.4g frac %.3f ceiling %.4g' % (e['value'], e['frac'], e['pcie_ceiling']['value']))
PY
done; done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rollout_equals or zero_copy" 2>&1 | tail -2
