#!/bin/bash
# Policy-kernel iteration check: numerics + timing, in-kernel trace, rollout tests, C3 legs
out=gpurun_out; mkdir -p $out; tag=${1:-pq}
timeout 120 python scripts/policy_tc_check.py > $out/${tag}_tc_check.log 2>&1; echo "tc_check rc=$?"; grep "us per\|ALL OK\|FAIL\|Error\|error" $out/${tag}_tc_check.log | tail -8
timeout 60 python scripts/policy_tc_trace.py 2>&1 | tail -3
timeout 400 python -m pytest tests/test_gpu_rollout.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --legs c3,c3_sb3 --steps 240 --warmup 24 --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d['legs'].items():
    if isinstance(v,dict): print(k, '%.4g'%v.get('value',0), 'ms %.5f'%v.get('ms_per_step',0), v.get('policy_forward_ms'), v.get('error',''))
"
