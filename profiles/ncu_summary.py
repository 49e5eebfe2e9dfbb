#!/usr/bin/env python
"""Text summary of one `ncu --set full` capture of the trimming kernel: key raw metrics (per launch), stall reasons,
pipe utilisation. usage: ncu_summary.py <file.ncu-rep> <pairs per launch>"""
import csv
import subprocess
import sys

rep, pairs = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, unit, val = rows[0], rows[1], rows[-1]
m = {h: (v, u) for h, u, v in zip(hdr, unit, val)}
def g(k):
    return m.get(k, ("n/a", ""))
def f(k):
    try:
        return float(g(k)[0].replace(",", ""))
    except ValueError:
        return float("nan")
print(f"kernel            : {g('Kernel Name')[0]}  grid {g('launch__grid_size')[0]} x block {g('launch__block_size')[0]}, {g('launch__registers_per_thread')[0]} regs/thread")
dur = f("gpu__time_duration.sum")
du = g("gpu__time_duration.sum")[1]
print(f"duration          : {dur} {du} (under the profiler: cold cache, serialised -- not a bench value)")
rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
ru, wu = g("dram__bytes_read.sum")[1], g("dram__bytes_write.sum")[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rb, wb = rd * scale.get(ru, 1), wr * scale.get(wu, 1)
print(f"dram traffic      : read {rb/1e6:.1f} MB + write {wb/1e6:.1f} MB per launch = {(rb+wb)/pairs:.1f} B per pair (algorithmic 608 B at 2x150)")
print(f"dram throughput   : {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[0]} % of peak")
inst = f("smsp__inst_executed.sum")
print(f"warp instructions : {inst:.4g} per launch = {inst/pairs:.0f} per pair")
print(f"issue slots busy  : {g('smsp__issue_active.avg.pct_of_peak_sustained_active')[0]} %   warps active {g('sm__warps_active.avg.pct_of_peak_sustained_active')[0]} % of 64/SM")
for p in ("alu", "fma", "xu", "lsu", "adu", "cbu", "uniform"):
    print(f"pipe {p:8s}     : {g(f'sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active')[0]} % of peak")
print("stall reasons (warps per issue-active cycle):")
for s in ("wait", "not_selected", "long_scoreboard", "short_scoreboard", "math_pipe_throttle", "no_instruction", "branch_resolving", "barrier", "mio_throttle", "lg_throttle", "dispatch_stall", "sleeping", "membar"):
    print(f"  {s:20s} {g(f'smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio')[0]}")
