#!/bin/bash
# Quick iteration check: selected GPU tests, C4 bench with chosen legs, TC policy check + trace.
tag=${1:-q2}; legs=${2:-c3}; tests=${3:-"tests/test_gpu_rollout.py tests/test_gpu_parity.py"}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest $tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py --no-cpu --legs $legs > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('$out/${tag}_bench.json').read().strip().splitlines()[-1])
print('value %.4g  ms/step %.5f  frac %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))
for k,v in d.get('legs',{}).items():
    if isinstance(v,dict): print(k, '%.4g'%v.get('value',0), 'ms %.5f'%v.get('ms_per_step',0), (v.get('roofline') or {}).get('frac'), v.get('error',''))
PY
timeout 300 python scripts/policy_tc_check.py > $out/${tag}_tc_check.log 2>&1; echo "tc_check rc=$?"; grep "us per\|ALL OK\|FAIL" $out/${tag}_tc_check.log | tail -6
timeout 120 python scripts/policy_tc_trace.py > $out/${tag}_tc_trace.log 2>&1; cat $out/${tag}_tc_trace.log
