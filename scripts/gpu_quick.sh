#!/bin/bash
# Quick GPU check: parity tests + bench line (no ncu).  Usage: scripts/gpu_quick.sh <tag> [bench args]
tag=${1:-q}; shift
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
timeout 600 python bench.py --no-cpu "$@" > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('$out/${tag}_bench.json'))
print('value %.4g  ms/step %.4f  frac %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))
PY
