"""DRAM traffic of the conv_tc launches of ONE swap step, from an ncu metric log of tools/one_step.py
(the same three counters an `ncu --set full` capture holds; collecting only them costs 2 replays per kernel
instead of ~40 -- a full capture of all 73 launches is a 150 MB report):
    ncu --profile-from-start off -k regex:conv_tc --clock-control none --csv \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --log-file gpurun_out/conv_tc_dram.csv python tools/one_step.py
    python tools/ncu_traffic.py gpurun_out/conv_tc_dram.csv > profiles/rNN_conv_tc_traffic.json
bench.py reads the JSON for roofline.traffic (bytes per launch, averaged over the launches of a step)."""
import csv
import json
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, iid = hdr.index("Kernel Name"), hdr.index("ID")
im, iu, iv = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
launches = {}
for r in rows[1:]:
    if "conv_tc" not in r[ik]:
        continue
    L = launches.setdefault(r[iid], {"name": re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("ppst::", "")})
    L[r[im]] = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
n = len(launches)
rd = sum(L.get("dram__bytes_read.sum", 0.0) for L in launches.values())
wr = sum(L.get("dram__bytes_write.sum", 0.0) for L in launches.values())
us = sum(L.get("gpu__time_duration.sum", 0.0) for L in launches.values())
per = {}
for L in launches.values():
    p = per.setdefault(L["name"], {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "ncu_us": 0.0})
    p["launches"] += 1
    p["dram_read"] += L.get("dram__bytes_read.sum", 0.0)
    p["dram_write"] += L.get("dram__bytes_write.sum", 0.0)
    p["ncu_us"] += L.get("gpu__time_duration.sum", 0.0)
print(json.dumps({"source": sys.argv[1], "kernel": "conv_tc (all launches of one swap step, 8 pairs @512^2)",
                  "launches": n, "dram_bytes_read": rd, "dram_bytes_write": wr,
                  "traffic_bytes_per_launch": (rd + wr) / max(n, 1), "ncu_duration_us_total": us,
                  "by_template": per}, indent=1))
