#!/bin/bash
# Full GPU check of the in-tree build: the -m gpu suite (timed), smoke, short C4 bench, ncu source-level capture of the step kernel.
out=gpurun_out; mkdir -p $out; tag=${1:-r3}
s=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$? in $(( $(date +%s) - s )) s"; tail -3 $out/${tag}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
B="python bench.py --steps 24 --warmup 24 --no-cpu --e2e-steps 2 --graph-steps 0 --legs none"
timeout 300 $B > $out/${tag}_plain_c4.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_ -s 30 -c 1 -o $out/${tag}_step_c4_full -f $B > $out/${tag}_ncu_c4.log 2>&1
tail -1 $out/${tag}_ncu_c4.log
timeout 300 python bench.py --no-cpu --legs none --steps 1200 --warmup 120 --e2e-steps 2 > $out/${tag}_c4.json 2>$out/${tag}_c4.err
cut -c1-300 $out/${tag}_c4.json
