"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: launches, total us, share."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("ppst::", "")
    name = re.sub(r"<.*", lambda mt: mt.group(0)[:40], name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
print("%d launches, %.1f us total (cold-cache, serialised per-launch times)" % (sum(v[0] for v in agg.values()), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%9.1f us  %5.1f%%  x%-4d %s" % (v[1], 100 * v[1] / tot, v[0], k))
