#!/usr/bin/env python
"""Turn the ncu launch list of one `bench.py --steps 1 --warmup 1 --no-graph` run
(ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv) into
  * a per-kernel table of the LAST full forward pass (steady state, calibrated label map), and
  * the conv-trunk DRAM traffic per launch group (bench.py reads it for roofline.traffic).
usage: python profiles/step_breakdown.py launches.csv out_prefix"""
import collections
import csv
import json
import sys

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
by = collections.OrderedDict()
for r in rows:
    d = by.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * SCALE[r["Metric Unit"]]
ids = sorted(by)
starts = [i for i in ids if "k_conv1_tc" in by[i]["name"]]
ends = [i for i in ids if "k_nms_pose" in by[i]["name"]]
# last forward that ran to the end (conv1 ... nms_pose)
s = max(i for i in starts if any(e > i for e in ends))
e = min(x for x in ends if x > s)
lines, total, trunk_t, trunk_b = [], 0.0, 0.0, 0.0
for i in ids:
    if i < s or i > e:
        continue
    d = by[i]
    t = d["gpu__time_duration.sum"]
    b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    total += t
    if i <= s + 13:                      # conv1_1 ... conv5_3 incl. the one un-fused max pool
        trunk_t += t; trunk_b += b
    lines.append(f"{t:9.1f} us {b / 1e6:9.1f} MB  {d['name'][:90]}")
out = sys.argv[2]
with open(out + "_step_breakdown.txt", "w") as f:
    f.write("one full-pipeline step, batch 32 x 640x480, eager launches under ncu (cold caches, serialised: compare shares)\n")
    f.write("\n".join(lines))
    f.write(f"\ntotal {total:.1f} us over {len(lines)} launches; conv trunk {trunk_t:.1f} us, DRAM traffic {trunk_b / 1e9:.3f} GB\n")
json.dump({"trunk_dram_bytes_per_launch_group": trunk_b, "trunk_us_under_ncu": trunk_t, "batch": 32,
           "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the 14 trunk launches of one step"},
          open(out + "_trunk_traffic.json", "w"), indent=1)
print(open(out + "_step_breakdown.txt").read())
