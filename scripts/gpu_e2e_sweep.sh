#!/bin/bash
# e2e (sng_step_host) sweep: direct PCIe observation stores on/off x host chunks
out=gpurun_out; mkdir -p $out
for d in 0 1; do for hc in 4 8 16 32; do
  SNG_HOST_DIRECT=$d timeout 300 python bench.py --no-cpu --legs none --steps 240 --warmup 24 --e2e-steps 48 --host-chunks $hc > $out/e2e_d${d}_hc${hc}.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('$out/e2e_d${d}_hc${hc}.json').read().strip().splitlines()[-1])
e=d['e2e']; print('direct=$d chunks=$hc e2e %.4g frac %.3f ceiling %.4g d2h %.1f GB/s' % (e['value'], e['frac'], e['pcie_ceiling']['value'], e['pcie_ceiling']['d2h_gbs_per_gpu']))
PY
done; done
