#!/bin/bash
# The driver's round-end sequence on a fresh box: GPU tests, smoke(), the reference arm, the default bench.
tag=${1:-final}; out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/${tag}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log
s=$(date +%s); timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench.err; echo "reference arm rc=$? in $(( $(date +%s) - s )) s"
s=$(date +%s); timeout 900 python bench.py > $out/${tag}_bench_1gpu.json 2>> $out/${tag}_bench.err; echo "bench rc=$? in $(( $(date +%s) - s )) s"
python - <<PY
import json
d=json.loads(open('$out/${tag}_bench_1gpu.json').read().strip().splitlines()[-1])
print('value %.4e ms %.5f frac %.3f e2e %.3e launches %d' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches']))
for k,v in d['legs'].items():
    if isinstance(v,dict) and 'value' in v: print('%-24s us/step %8.3f value %.3e' % (k, v['ms_per_step']*1e3, v['value']))
    elif isinstance(v,dict) and 'error' in v: print(k, v)
PY
