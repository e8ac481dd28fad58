#!/usr/bin/env python
"""ncu --csv launch list with dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum -> the JSON bench.py
reads as roofline.traffic:  python profiles/traffic_json.py gpurun_out/r2y_traffic.csv > profiles/r02_traffic_steady.json"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
h = rows[hi]
kn, mn, mu, mv, idc = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("ID")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "ms": 1e6}
acc = collections.defaultdict(lambda: collections.defaultdict(float))
ids = collections.defaultdict(set)
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    acc[r[kn]][r[mn]] += float(r[mv].replace(",", "")) * UNIT[r[mu]]
    ids[r[kn]].add(r[idc])
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none "
                 "-k regex:step_kernel -s 300 -c 192 on `bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-socket --no-policy --min-time-ms 5 "
                 "--e2e-steps 3` (profiles/run_r2y.sh, final round-2 tree): steady state, the profiled launches (kernel nodes of the replayed "
                 "graphs) rotate over 6 shards, caches not flushed between them; raw rows in r02_traffic_steady.csv",
       "algorithmic_bytes_per_launch": {"step_kernel<float,AUTO,OBS>": 146 * 1048576, "step_kernel<float,AUTO,!OBS>": 86 * 1048576},
       "kernels": {}}
for k, m in acc.items():
    n = len(ids[k])
    out["kernels"][k] = {"launches": n, "dram_read_bytes_per_launch": m["dram__bytes_read.sum"] / n,
                         "dram_write_bytes_per_launch": m["dram__bytes_write.sum"] / n,
                         "dram_bytes_per_launch": (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / n,
                         "ns_per_launch_under_ncu": m["gpu__time_duration.sum"] / n}
print(json.dumps(out, indent=1))
