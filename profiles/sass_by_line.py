#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

usage: sass_by_line.py <ncu source-page csv (--page source --csv)> <nvdisasm -g -c output of the cubin> <kernel substring> [top]
The ncu CLI exports per-instruction counters only at SASS level; nvdisasm -g carries the line table, both list the
kernel's instructions in address order, so they are joined by position.
"""
import csv
import re
import sys
from collections import defaultdict

ncu_csv, sass_path, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
insts = [(r[ci["Source"]].strip(), int(r[ci["Instructions Executed"]]), int(r[ci["Warp Stall Sampling (All Samples)"]])) for r in rows[2:] if len(r) > 5]

# nvdisasm: find function section, collect (line, opcode) for each instruction
lines = open(sass_path).read().split("\n")
cur = None
infn = False
seq = []
inl = None
for ln in lines:
    if ln.startswith(".text.") or ln.startswith("\t.section\t.text."):
        infn = kern in ln
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        seq.append((cur, m.group(2).strip()))
print(f"ncu instructions: {len(insts)}, nvdisasm instructions: {len(seq)}", file=sys.stderr)
n = min(len(insts), len(seq))
by = defaultdict(lambda: [0, 0, 0])
tot = 0
tots = 0
for (src, cnt, stall), (loc, op) in zip(insts[:n], seq[:n]):
    by[loc][0] += cnt
    by[loc][1] += stall
    by[loc][2] += 1
    tot += cnt
    tots += stall
print(f"total warp instructions {tot}, stall samples {tots}")
for loc, (c, s, k) in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100.0*c/tot:6.2f}% inst  {100.0*s/max(1,tots):6.2f}% stall  {k:5d} sass  {loc}")
