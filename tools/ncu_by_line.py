#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel: joins `ncu --page source --csv` (SASS view: executed counts, stall
samples) with `nvdisasm -g` line info of the cubin inside libmetad_b200.so.

    python tools/ncu_by_line.py <report.ncu-rep> <kernel regex> <mangled-name substring> [units] [top]

units: divide thread-instruction counts by this number (e.g. particles per launch)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kre, sub = sys.argv[1:4]
units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "metadynamics_plugin_b200", "libmetad_b200.so")

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
iA, iS, iE, iSm = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
seen, sass = set(), []
for r in rows[2:]:
    if len(r) <= iE or r[iA] in seen or r[0] == "Address" or not r[iA].startswith("0x"):
        continue
    if len(seen) and r[iA] == rows[2][iA]:
        break
    seen.add(r[iA])
    sass.append((r[iS].strip(), int(r[iE]), int(r[iSm])))

with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, capture_output=True)
    dis = None
    for f in os.listdir(d):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        if sub in out:
            dis = out
            break
lines, cur, infn = [], ("?", 0), False
for ln in dis.splitlines():
    if ln.lstrip().startswith(".section"):
        infn = (".text." in ln and sub in ln)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m:
        lines.append((cur, m.group(1).strip()))
if len(lines) != len(sass):
    print("warning: %d disassembled instructions vs %d profiled" % (len(lines), len(sass)))
by = collections.OrderedDict()
for (loc, _), (op, e, s) in zip(lines, sass):
    a = by.setdefault(loc, [0, 0])
    a[0] += e
    a[1] += s
tot = sum(v[0] for v in by.values())
ts = sum(v[1] for v in by.values()) or 1
print("total warp instructions %d (%.1f thread instructions per unit), %d samples" % (tot, tot * 32 / units, ts))
src_cache = {}
def src(loc):
    for dd in ("metadynamics_plugin_b200/csrc", "include"):
        p = os.path.join(root, dd, loc[0])
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[loc[1] - 1].strip()[:90] if 0 < loc[1] <= len(L) else ""
    return ""
for loc, (e, s) in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s:%4d  inst %5.1f%% (%6.1f/unit)  stall samples %5.1f%%   %s" % (loc[0], loc[1], 100 * e / tot, e * 32 / units, 100 * s / ts, src(loc)))
