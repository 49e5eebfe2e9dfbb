#!/usr/bin/env python
"""Executed warp instructions and stall samples per source function (spg_kernel.cuh / spg_lanes.cuh) and per opcode, from an ncu
source-page CSV joined by position with `nvdisasm -g -c` of the same cubin. An instruction belongs to the innermost frame of its
inline chain that lies in one of the two files.
usage: sass_regions2.py <ncu source csv> <nvdisasm sass> <kernel substring> <pairs per launch>"""
import csv, re, sys, os
from collections import defaultdict

ncu_csv, sass_path, kern, pairs = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
HERE = os.path.dirname(os.path.abspath(__file__))
starts = {}
for fn in ("spg_kernel.cuh", "spg_lanes.cuh", "spg_qc.cuh", "spg_fastq.cuh"):
    path = os.path.join(HERE, "..", "ngs-bits_b200", "csrc", fn)
    if not os.path.exists(path):
        continue
    st = []
    for i, ln in enumerate(open(path).read().split("\n"), 1):
        m = re.match(r"^(?:__device__|__global__).*?\b([a-zA-Z_0-9]+)\s*\(", ln)
        if m:
            st.append((i, m.group(1)))
    starts[fn] = st
def region_of(fn, line):
    name = fn + ":preamble"
    for st, nm in starts[fn]:
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
by = defaultdict(lambda: [0, 0, 0]); ops = defaultdict(int); tot = 0; stot = 0
outer = defaultdict(lambda: [0, 0])
for (c, s), (chain, op) in zip(insts, seq):
    r = "nolineinfo"; o = "nolineinfo"
    for f, l in (chain or []):
        if f in starts:
            r = region_of(f, l); break
    for f, l in reversed(chain or []):
        if f == "spg_lanes.cuh":
            o = f"lanes.cuh:{l}"; break
    by[r][0] += c; by[r][1] += s; by[r][2] += 1; tot += c; stot += s; ops[op] += c
    outer[o][0] += c; outer[o][1] += s
print(f"warp instructions per pair: {tot / pairs:.1f}   (stall samples {stot}; static instructions {len(seq)})")
for r, (c, s, k) in sorted(by.items(), key=lambda kv: -kv[1][0]):
    print(f"  {r:28s} {100 * c / tot:6.2f}%  {c / pairs:8.1f} inst/pair   {100 * s / max(1, stot):6.2f}% of stall samples   {k:5d} static")
print("opcodes:", ", ".join(f"{k} {v / pairs:.1f}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:26]))
if "--lines" in sys.argv:
    print("by line of the kernel body (outermost frame in spg_lanes.cuh):")
    for o, (c, s) in sorted(outer.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"  {o:20s} {c / pairs:8.1f} inst/pair  {100 * s / max(1, stot):6.2f}% stalls")
