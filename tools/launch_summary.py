#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): count, mean, total (microseconds)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[h]
iN, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) > iV:
        agg.setdefault(re.sub(r"\(.*", "", r[iN]), []).append(float(r[iV].replace(",", "")) / 1000.0)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-70s n=%4d  mean %9.2f us  total %10.2f us" % (k[:70], len(v), sum(v) / len(v), sum(v)))
