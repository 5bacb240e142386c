#!/usr/bin/env python
"""Instruction counts per kernel from `cuobjdump -sass` of the in-tree library (no GPU needed).

    python tools/sass_summary.py [path/to/libmetad_b200.so] [--filter substr] > profiles/rNN_sass_summary.txt

Per kernel: total instructions, the mnemonics that prove which engine moves the tiles (UTMALDG / UTMAREDG = tensor-map
TMA, UBLKCP / UBLKRED = bulk TMA, LDGSTS = cp.async, ATOMS / RED / ATOM), packed fp32 (FFMA2 / FMUL2 / FADD2), and the
loops that carry the per-particle work: the backward branches with the largest spans, their instruction counts and mix.
For an issue-bound kernel the loop size is the figure of merit (instructions per particle and thread)."""
import collections
import re
import subprocess
import sys

KEYS = ("UTMALDG", "UTMAREDG", "UTMASTG", "UBLKCP", "UBLKRED", "LDGSTS", "ATOMS", "ATOMG", "REDG", "RED", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL",
        "FADD", "LDS", "STS", "LDG", "STG", "BAR", "SYNCS", "SHFL", "DFMA", "DADD", "DMUL", "MUFU", "I2F", "F2I", "BRA")


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    name, rows = None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                yield name, rows
            name, rows = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            rows.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        yield name, rows


def opcode(text):
    t = text.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0] if t else ""


def main():
    argv = sys.argv[1:]
    filt = None
    if "--filter" in argv:
        i = argv.index("--filter")
        filt = argv[i + 1]
        del argv[i:i + 2]
    path = argv[0] if argv else "metadynamics_plugin_b200/libmetad_b200.so"
    for mangled, rows in kernels(path):
        name = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        if filt and filt not in name:
            continue
        ops = collections.Counter(opcode(t) for _, t in rows)
        total = sum(ops.values())
        addr = {a: i for i, (a, _) in enumerate(rows)}
        loops = []
        for i, (a, t) in enumerate(rows):
            if opcode(t) != "BRA":
                continue
            m = re.search(r"0x([0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt in addr and addr[tgt] < i:
                    loops.append((i - addr[tgt] + 1, addr[tgt], i))
        print("%-100s total %5d" % (name[:100], total))
        print("    " + "  ".join("%s=%d" % (k, ops[k]) for k in KEYS if ops.get(k)))
        for span, a, b in sorted(loops, reverse=True)[:3]:
            lo = collections.Counter(opcode(t) for _, t in rows[a:b + 1])
            print("    loop %4d instr: " % span + "  ".join("%s=%d" % (k, lo[k]) for k in KEYS if lo.get(k)))


if __name__ == "__main__":
    main()
