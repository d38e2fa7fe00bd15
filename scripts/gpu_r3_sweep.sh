#!/bin/bash
# C4 step-kernel sweep over variant libraries in scratch/ (built by scripts/build_variant.sh); parity of the last one.
# Usage: scripts/gpu_r3_sweep.sh tag name1 name2 ...   ("base" = the in-tree library)
out=gpurun_out; mkdir -p $out
tag=$1; shift
run() { lib=smart_nanogrid_gym_b200/libsng.so; [ "$1" != base ] && lib=scratch/libsng_$1.so
  SNG_LIB_PATH=$lib timeout 300 python bench.py --no-cpu --legs none --steps 1200 --warmup 120 --e2e-steps 2 ${@:2} > $out/${tag}_$1.json 2>$out/${tag}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('$out/${tag}_$1.json').read().strip().splitlines()[-1])
    print('$1: ms/step %.5f frac %.3f ret %.3f' % (d['ms_per_step'], d['roofline']['frac'], d['mean_episode_return']))
except Exception as e:
    print('$1: FAILED', e)
PY
}
for n in "$@"; do run $n; done
run base
last="${@: -1}"
if [ "$last" != base ]; then
  SNG_LIB_PATH=$PWD/scratch/libsng_$last.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/${tag}_pytest_$last.log 2>&1; echo "pytest($last) rc=$?"; tail -3 $out/${tag}_pytest_$last.log
fi
