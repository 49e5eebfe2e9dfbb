#!/usr/bin/env python
"""Text summary of one `ncu --set full` capture of the trimming kernel: key raw metrics (per launch), stall reasons,
pipe utilisation. usage: ncu_summary.py <file.ncu-rep | raw-page .csv> <pairs per launch> [algorithmic bytes per pair] [--traffic out.json capture-name]"""
import csv
import subprocess
import sys

rep, pairs = sys.argv[1], float(sys.argv[2])
out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
balg = float(sys.argv[3]) if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else 608.0
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
print(f"dram traffic      : read {rb/1e6:.1f} MB + write {wb/1e6:.1f} MB per launch = {(rb+wb)/pairs:.1f} B per pair (algorithmic {balg:.0f} B)")
print(f"dram throughput   : {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[0]} % of peak")
inst = f("smsp__inst_executed.sum")
print(f"warp instructions : {inst:.4g} per launch = {inst/pairs:.0f} per pair")
print(f"issue slots busy  : {g('smsp__issue_active.avg.pct_of_peak_sustained_active')[0]} %   warps active {g('sm__warps_active.avg.pct_of_peak_sustained_active')[0]} % of 64/SM")
for p in ("alu", "fma", "xu", "lsu", "adu", "cbu", "uniform"):
    print(f"pipe {p:8s}     : {g(f'sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active')[0]} % of peak")
print("stall reasons (warps per issue-active cycle):")
for s in ("wait", "not_selected", "long_scoreboard", "short_scoreboard", "math_pipe_throttle", "no_instruction", "branch_resolving", "barrier", "mio_throttle", "lg_throttle", "dispatch_stall", "sleeping", "membar"):
    print(f"  {s:20s} {g(f'smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio')[0]}")

if "--traffic" in sys.argv:
    import json
    i = sys.argv.index("--traffic")
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    json.dump({"dram_bytes_per_pair": (rb + wb) / pairs, "kernel": g("Kernel Name")[0], "capture": sys.argv[i + 2], "commit": commit,
               "pairs_per_launch": pairs, "how": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of bench.py's headline config"},
              open(sys.argv[i + 1], "w"), indent=1)
