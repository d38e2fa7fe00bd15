#!/bin/bash
# C3 legs with variant libraries in scratch/ (policy-kernel experiments).  Usage: gpu_r3_c3lib.sh name1 name2 ... ("base" = in-tree)
out=gpurun_out; mkdir -p $out
for n in "$@"; do
  lib=smart_nanogrid_gym_b200/libsng.so; [ "$n" != base ] && lib=scratch/libsng_$n.so
  SNG_LIB_PATH=$lib timeout 300 python bench.py --no-cpu --legs c3,c3_sb3 --steps 48 --warmup 24 --e2e-steps 2 > $out/c3lib_$n.json 2>$out/c3lib_$n.err
  python - <<PY
import json
try:
    d=json.loads(open('$out/c3lib_$n.json').read().strip().splitlines()[-1])
    print('$n: ' + '  '.join('%s %.2f us (pol %.4f ms)' % (k, v['ms_per_step']*1e3, v.get('policy_forward_ms',0)) for k,v in d['legs'].items() if isinstance(v,dict) and 'value' in v))
except Exception as e: print('$n FAILED', e)
PY
done
last="${@: -1}"
if [ "$last" != base ]; then SNG_LIB_PATH=$PWD/scratch/libsng_$last.so timeout 300 python -m pytest tests/test_gpu_rollout.py -m gpu -x -q 2>&1 | tail -2; fi
