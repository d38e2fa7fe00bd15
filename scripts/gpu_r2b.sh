#!/bin/bash
# Round-2 checkpoint on the GPU box: parity tests, smoke, bench (all legs), TC policy check + timing trace.
tag=${1:-r2b}
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python scripts/policy_tc_check.py > $out/${tag}_tc_check.log 2>&1; echo "tc_check rc=$?"; tail -8 $out/${tag}_tc_check.log
timeout 120 python scripts/policy_tc_trace.py > $out/${tag}_tc_trace.log 2>&1; cat $out/${tag}_tc_trace.log
echo done
