#!/usr/bin/env python
"""Folds the ncu CSVs written by scripts/ncu_traffic.sh (gpurun_out/traffic_<workload>_<envs>.csv) into
profiles/roofline_traffic.json: per workload and batch size, the mean DRAM read / write bytes per launch of the step
kernel over the profiled steady-state launches.  Runs where ncu ran or here (it only parses)."""
import csv
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1.0, "msecond": 1e6}


def parse(path):
    rows = []
    with open(path) as fp:
        lines = [l for l in fp if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    per = {}
    for r in rows:
        key = r["ID"]
        per.setdefault(key, {"kernel": r["Kernel Name"]})
        per[key][r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    return list(per.values())


def main():
    import bench
    from smart_nanogrid_gym_b200.config import NanogridConfig
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    with open(path) as fp:
        t = json.load(fp)
    src_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
    for f in sorted(glob.glob(os.path.join(src_dir, "traffic_*.csv"))):
        m = re.match(r"traffic_(\w+?)_(\d+)\.csv", os.path.basename(f))
        if not m:
            continue
        wl, envs = m.group(1), int(m.group(2))
        launches = parse(f)
        if not launches:
            continue
        n = len(launches)
        rd = sum(x["dram__bytes_read.sum"] for x in launches) / n
        wr = sum(x["dram__bytes_write.sum"] for x in launches) / n
        cfg = NanogridConfig(**bench.WORKLOADS[wl]["kw"])
        alg = bench.algorithmic_bytes_per_env_step(cfg.n_spots, int(cfg.batt), int(cfg.pv), cfg.hours_ahead) * envs
        e = {"kernel": launches[0]["kernel"], "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
             "algorithmic_bytes_per_launch": alg, "launches_profiled": n,
             "ncu_duration_us": sum(x.get("gpu__time_duration.sum", 0.0) for x in launches) / n / 1e3,
             "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, launches 30..%d of plain per-step launches (steady state), %s" % (
                 29 + n, os.path.basename(f))}
        if "lts__t_bytes.sum" in launches[0]:
            e["l2_bytes_per_launch"] = sum(x["lts__t_bytes.sum"] for x in launches) / n
        t["entries"]["%s:%d" % (wl, envs)] = e
        print("%s:%d  read %.1f MB  write %.1f MB  total %.1f MB  (algorithmic %.1f MB)" % (wl, envs, rd / 1e6, wr / 1e6, (rd + wr) / 1e6, alg / 1e6))
    with open(path, "w") as fp:
        json.dump(t, fp, indent=1)


if __name__ == "__main__":
    main()
