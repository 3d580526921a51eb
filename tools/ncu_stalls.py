#!/usr/bin/env python3
"""Rank SASS instructions of an ncu source-page CSV by a stall reason, with the CUDA line each maps to.
usage: ncu_stalls.py <src.csv> <disasm from nvdisasm -gi -c> <mangled kernel .text label> <stall column> [N]"""
import collections
import csv
import re
import sys

src_csv, disasm, kname, col = sys.argv[1:5]
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 30
lines = open(disasm).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(kname + ":"))
fn_re = re.compile(r'//## File "([^"]+)", line (\d+)')
ins_re = re.compile(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);")
block, addr2, prev = [], {}, False
for l in lines[start + 1:]:
    if l.startswith(".text.") and addr2:
        break
    m = fn_re.search(l)
    if m:
        if prev:
            block = []
        block.append((m.group(1).split("/")[-1], int(m.group(2))))
        prev = False
        continue
    m = ins_re.match(l)
    if m:
        addr2[int(m.group(1), 16)] = tuple(block)
        prev = True
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
base, items, tot = None, [], 0
for r in rows[2:]:
    try:
        a, v = int(r[idx["Address"]], 16), int(r[idx[col]])
    except (ValueError, IndexError):
        continue
    if base is None:
        base = a
    tot += v
    items.append((v, a - base, r[idx["Source"]].strip(), addr2.get(a - base, ())))
byline = collections.Counter()
for v, a, s, ch in items:
    byline[ch[:2]] += v
print(f"{col}: total {tot}")
for key, v in byline.most_common(topn):
    print(f"{100 * v / max(tot, 1):5.1f}%  " + " <- ".join(f"{f.split('.')[0][:10]}:{ln}" for f, ln in key))
print("--- top instructions")
for v, a, s, ch in sorted(items, reverse=True)[:topn]:
    print(f"{100 * v / max(tot, 1):5.1f}%  {a:6x}  {s[:60]:60s} " + " <- ".join(f"{f.split('.')[0][:10]}:{ln}" for f, ln in ch[:3]))
