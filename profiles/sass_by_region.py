#!/usr/bin/env python
"""Executed warp instructions per code region (function of spg_kernel.cuh) and per opcode, from an ncu source-page CSV
joined by position with `nvdisasm -g -c` of the same cubin.
usage: sass_by_region.py <ncu source csv> <nvdisasm sass> <kernel substring> <pairs per launch>"""
import csv, re, sys
from collections import defaultdict

ncu_csv, sass_path, kern, pairs = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
src = open(sys.argv[5] if len(sys.argv) > 5 else "ngs-bits_b200/csrc/spg_kernel.cuh").read().split("\n")
starts = []
for i, ln in enumerate(src, 1):
    m = re.match(r"^(?:__device__|__global__).*?\b([a-z_0-9]+)\s*\(", ln)
    if m:
        starts.append((i, m.group(1)))
def region_of(line):
    name = "preamble"
    for st, nm in starts:
        if line >= st - 1:
            name = nm
    return name
rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
insts = [(int(r[ci["Instructions Executed"]]), int(r[ci["Warp Stall Sampling (All Samples)"]])) for r in rows[2:] if len(r) > 5]
infn = False; seq = []; cur = None
for ln in open(sass_path).read().split("\n"):
    if ln.startswith(".text.") or ln.startswith("\t.section\t.text."):
        infn = kern in ln; continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        chain = [(m.group(1).split("/")[-1], int(m.group(2)))]
        for mm in re.finditer(r'inlined at "([^"]+)", line (\d+)', m.group(3)):
            chain.append((mm.group(1).split("/")[-1], int(mm.group(2))))
        cur = chain; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        toks = m.group(2).split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        seq.append((cur, op.split(".")[0]))
assert len(seq) == len(insts), (len(seq), len(insts))
by = defaultdict(lambda: [0, 0]); ops = defaultdict(int); tot = 0; stot = 0
for (c, s), (chain, op) in zip(insts, seq):
    r = "nolineinfo"
    for f, l in (chain or []):
        if f == "spg_kernel.cuh":
            r = region_of(l); break
    by[r][0] += c; by[r][1] += s; tot += c; stot += s; ops[op] += c
print(f"warp instructions per pair: {tot / pairs:.1f}   (stall samples {stot})")
for r, (c, s) in sorted(by.items(), key=lambda kv: -kv[1][0]):
    print(f"  {r:24s} {100 * c / tot:6.2f}%  {c / pairs:8.1f} inst/pair   {100 * s / max(1, stot):6.2f}% of stall samples")
print("opcodes:", ", ".join(f"{k} {v / pairs:.0f}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:22]))
