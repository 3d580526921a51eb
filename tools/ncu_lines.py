#!/usr/bin/env python3
"""Attribute an ncu source-page CSV (SASS view) to CUDA source lines using nvdisasm -gi output.

usage: ncu_lines.py <src.csv from `ncu -i X --page source --csv`> <disasm from `nvdisasm -gi -c cubin`>
                    <mangled kernel name> <source file> [--func]
Prints the innermost source line (or, with --func, the innermost two frames) ranked by
warp-instructions executed, with stall-sample share.
"""
import collections
import csv
import re
import sys

src_csv, disasm, kname, srcfile = sys.argv[1:5]
by_func = "--func" in sys.argv
lines = open(disasm).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(kname + ":"))
fn_re = re.compile(r'//## File "([^"]+)", line (\d+)')
ins_re = re.compile(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);")
block, addr2 = [], {}
prev_ins = False
for l in lines[start + 1:]:
    if l.startswith(".text.") and addr2:
        break
    m = fn_re.search(l)
    if m:
        if prev_ins:
            block = []
        block.append((m.group(1).split('/')[-1], int(m.group(2))))
        prev_ins = False
        continue
    m = ins_re.match(l)
    if m:
        addr2[int(m.group(1), 16)] = tuple(block)
        prev_ins = True
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
agg, samp = collections.Counter(), collections.Counter()
tot = ts = 0
base = None
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        ie, ns, a = int(r[idx["Instructions Executed"]]), int(r[idx["# Samples"]]), int(r[idx["Address"]], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    chain = addr2.get(a - base, ())
    key = chain[:3] if by_func else chain[:2]
    agg[key] += ie
    samp[key] += ns
    tot += ie
    ts += ns
text = open(srcfile).read().split("\n")
print(f"total warp-instructions {tot}, samples {ts}")
for key, v in agg.most_common(60):
    where = " <- ".join(f"{f.split('.')[0][:12]}:{ln}" for f, ln in key)
    t = ""
    for f, ln in key:
        if f == srcfile.split("/")[-1] and 0 < ln <= len(text):
            t = text[ln - 1].strip()[:80]
            break
    print(f"{100 * v / tot:5.1f}% inst {100 * samp[key] / max(1, ts):5.1f}% samp  L{where:18s} {t}")
