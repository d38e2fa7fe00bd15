#!/bin/bash
# GPU suite + the rollout / small-batch legs of the bench
out=gpurun_out; mkdir -p $out; tag=${1:-r3b}
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py --no-cpu --legs c3,c2,rollout_kernel,c5 --steps 240 --warmup 24 --e2e-steps 2 > $out/${tag}_legs.json 2>$out/${tag}_legs.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('$out/${tag}_legs.json').read().strip().splitlines()[-1])
print('headline ms', d['ms_per_step'])
for k,v in d['legs'].items():
    if isinstance(v,dict) and 'value' in v: print('%-24s us/step %8.3f value %.3e' % (k, v['ms_per_step']*1e3, v['value']))
PY
