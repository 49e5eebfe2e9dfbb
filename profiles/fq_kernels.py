#!/usr/bin/env python
"""Per-kernel device time and DRAM traffic of the FASTQ stream path (spg_fq_*: framing, packing, trimming, output assembly)
on plain-text FASTQ input, from an ncu launch list of the command line itself:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv seqpurge_b200 ...
Numbers under the profiler are serialised and cold-cache: read them as per-kernel shares and bytes, not as throughput claims.
usage: python profiles/fq_kernels.py [pairs] [block_size]   -> gpurun_out/fq_kernels.json + a table on stdout"""
import csv
import io
import json
import os
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g

g.build()
import helpers as H

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
block = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
L = 150
b = H.synth_batch_numpy(n, L, seed=5)
d = tempfile.mkdtemp()
text_bytes = 0
for r, (B, Q) in enumerate(((b.bases1, b.quals1), (b.bases2, b.quals2)), start=1):
    with open(f"{d}/in{r}.fastq", "wb") as f:
        for i in range(n):
            f.write(b"@SIM:1:B200:1:%d:%d %d:N:0:ACGT\n" % (i // 100000, i, r))
            f.write(B[i, :L].tobytes() + b"\n+\n" + Q[i, :L].tobytes() + b"\n")
    text_bytes += os.path.getsize(f"{d}/in{r}.fastq")
cli = os.path.join(ROOT, "ngs-bits_b200", "bin", "seqpurge_b200")
cmd = ["ncu", "--metrics", "gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--csv",
       cli, "-in1", f"{d}/in1.fastq", "-in2", f"{d}/in2.fastq", "-out1", f"{d}/o1.fastq.gz", "-out2", f"{d}/o2.fastq.gz", "-compression_level", "0",
       "-block_size", str(block), "-threads", "8"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = [r for r in csv.reader(io.StringIO(out[out.index('"ID"'):]))]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: {"launches": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}
seen = set()
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "").replace("spg::", "")
    v = float(r[ci["Metric Value"]].replace(",", "")) * scale.get(r[ci["Metric Unit"]], 1)
    a = agg[name]
    m = r[ci["Metric Name"]]
    if m == "gpu__time_duration.sum":
        a["ns"] += v
        a["launches"] += 1
    elif m == "dram__bytes_read.sum":
        a["rd"] += v
    elif m == "dram__bytes_write.sum":
        a["wr"] += v
total_ns = sum(a["ns"] for a in agg.values())
res = {"pairs": n, "block_size": block, "input_text_bytes": text_bytes, "kernels": {}}
print(f"{n} pairs of 2x{L}, {text_bytes / n:.0f} B of input text per pair, chunks of {block} pairs; device time total {total_ns / 1e6:.2f} ms = {n / total_ns * 1e3:.0f} Mpairs/s if the kernels ran back to back")
print(f"{'kernel':44s} {'launches':>8s} {'ms':>8s} {'share':>6s} {'B/pair rd':>10s} {'B/pair wr':>10s} {'DRAM GB/s':>10s} {'of 6548':>8s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    gbs = (a["rd"] + a["wr"]) / a["ns"] if a["ns"] else 0.0
    print(f"{name[:44]:44s} {a['launches']:8d} {a['ns'] / 1e6:8.3f} {100 * a['ns'] / total_ns:5.1f}% {a['rd'] / n:10.1f} {a['wr'] / n:10.1f} {gbs:10.1f} {gbs / 6547.8:8.3f}")
    res["kernels"][name] = {"launches": a["launches"], "ms": a["ns"] / 1e6, "share": a["ns"] / total_ns, "dram_read_bytes_per_pair": a["rd"] / n,
                            "dram_write_bytes_per_pair": a["wr"] / n, "dram_gbs": gbs, "frac_of_measured_hbm": gbs / 6547.8}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "fq_kernels.json"), "w"), indent=1)

if os.environ.get("FQ_NCU_FULL"):  # one full capture each of the two copy kernels (raw and source pages as CSV under gpurun_out/)
    import gzip
    for k in ("fq_pack", "fq_out_write"):
        rep = f"/tmp/ncu_{k}"
        subprocess.run(["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "-k", f"regex:{k}", "-c", "1", "-o", rep, "-f"] + cmd[cmd.index(cli):],
                       capture_output=True, text=True)
        raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        open(os.path.join(ROOT, "gpurun_out", f"ncu_{k}_raw.csv"), "w").write(raw)
        src = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "source", "--csv"], capture_output=True, text=True).stdout
        gzip.open(os.path.join(ROOT, "gpurun_out", f"ncu_{k}_src.csv.gz"), "wt").write(src)
