#!/usr/bin/env python
"""End-to-end command-line comparison (the X boundary of SURVEY.md 8d: gz in -> trim -> gz out, informational):
seqpurge_b200 (CUDA engine) vs the CPU oracle CLI with all host threads, same synthetic FASTQ.gz input, outputs compared byte for byte.
usage: python profiles/cli_throughput.py [pairs] [copies]   (the generated files are concatenated `copies` times: multi-member gzip)"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ngs-bits_b200"))
import torch

import __graft_entry__ as g

g.build()
import seqpurge_b200 as sp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
L = 150
dev = torch.device("cuda:0")
t = {k: torch.empty((n, L), dtype=torch.uint8, device=dev) for k in ("bases1", "quals1", "bases2", "quals2")}
l1 = torch.empty(n, dtype=torch.int16, device=dev)
l2 = torch.empty(n, dtype=torch.int16, device=dev)
sp.synth_device(sp.SynthConfig(read_len=L), 0, n, t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2)
torch.cuda.synchronize()
host = {k: v.cpu().numpy() for k, v in t.items()}
d = tempfile.mkdtemp()
for r, (bk, qk) in enumerate((("bases1", "quals1"), ("bases2", "quals2")), start=1):
    p = subprocess.Popen(["gzip", "-1", "-c"], stdin=subprocess.PIPE, stdout=open(f"{d}/in{r}.fastq.gz", "wb"))
    B, Q = host[bk], host[qk]
    for i in range(n):
        p.stdin.write(b"@SIM:1:B200:1:%d:%d %d:N:0:ACGT\n" % (i // 100000, i, r))
        p.stdin.write(B[i].tobytes() + b"\n+\n" + Q[i].tobytes() + b"\n")
    p.stdin.close()
    p.wait()
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if copies > 1:
    for r in (1, 2):
        subprocess.run("cat " + " ".join([f"{d}/in{r}.fastq.gz"] * copies) + f" > {d}/inx{r}.fastq.gz && mv {d}/inx{r}.fastq.gz {d}/in{r}.fastq.gz", shell=True, check=True)
    n *= copies
os.environ["SPG_TIMING"] = "1"
threads = len(os.sched_getaffinity(0))
try:  # the cgroup limit is what the box really gives
    quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
    if quota != "max":
        threads = max(1, min(threads, int(int(quota) / int(period))))
except OSError:
    pass
cli = os.path.join(ROOT, "ngs-bits_b200", "bin", "seqpurge_b200")
plain = {}
for r in (1, 2):  # uncompressed copies of the inputs: the inflate-free upper bound of the pipeline
    plain[r] = f"{d}/in{r}.fastq"
    subprocess.run(f"gzip -dc {d}/in{r}.fastq.gz > {plain[r]}", shell=True, check=True)
bgz = {}
pipe = os.path.join(ROOT, "ngs-bits_b200", "bin", "gzpipe")
for r in (1, 2):  # the same inputs as BGZF (blocked gzip): inflated by the -threads pool
    bgz[r] = f"{d}/in{r}.bgzf.fastq.gz"
    subprocess.run([pipe, plain[r], bgz[r], "-bgzf", "-threads", str(threads)], check=True, stderr=subprocess.DEVNULL)
runs = {
    "oracle_cli": ([os.path.join(ROOT, "oracle", "build", "seqpurge_oracle"), "-threads", str(threads)], "gz"),
    "b200_host_framing": ([cli, "-host_framing"], "gz"),
    "seqpurge_b200": ([cli], "gz"),
    f"b200_deflate_threads_{threads}": ([cli, "-threads", str(threads), "-block_size", "32768"], "gz"),
    f"b200_plain_in_deflate_threads_{threads}": ([cli, "-threads", str(threads), "-block_size", "32768"], "plain"),
    f"b200_plain_in_level0_out_threads_{threads}": ([cli, "-threads", str(threads), "-block_size", "32768", "-compression_level", "0"], "plain"),
    f"b200_bgzf_in_threads_{threads}": ([cli, "-threads", str(threads), "-block_size", "32768"], "bgzf"),
    f"b200_bgzf_in_bgzf_out_threads_{threads}": ([cli, "-threads", str(threads), "-block_size", "32768", "-bgzf"], "bgzf"),
}
if os.environ.get("SPG_CLI_ONLY"):  # comma-separated run names (the oracle CLI alone takes half a minute)
    keep = set(os.environ["SPG_CLI_ONLY"].split(",")) | {"oracle_cli"}
    runs = {k: v for k, v in runs.items() if k in keep or any(k.startswith(x) for x in keep)}
for name, (cmd, kind) in runs.items():
    os.makedirs(f"{d}/{name}")
    in1, in2 = (f"{d}/in1.fastq.gz", f"{d}/in2.fastq.gz") if kind == "gz" else ((bgz[1], bgz[2]) if kind == "bgzf" else (plain[1], plain[2]))
    t0 = time.perf_counter()
    subprocess.run(cmd + ["-in1", in1, "-in2", in2, "-out1", f"{d}/{name}/o1.fastq.gz", "-out2", f"{d}/{name}/o2.fastq.gz", "-summary", f"{d}/{name}/s.txt"], check=True)
    el = time.perf_counter() - t0
    print(f"{name}: {el:.2f} s for {n} pairs = {n / el / 1e6:.3f} Mpairs/s end to end ({kind} in, gz out)", flush=True)
for name in runs:
    if name == "oracle_cli":
        continue
    same = all(open(f"{d}/oracle_cli/{f}", "rb").read() == open(f"{d}/{name}/{f}", "rb").read() for f in ("o1.fastq.gz", "o2.fastq.gz"))
    content = all(subprocess.run(f"bash -c 'cmp <(gzip -dc {d}/oracle_cli/{f}) <(gzip -dc {d}/{name}/{f})'", shell=True).returncode == 0 for f in ("o1.fastq.gz", "o2.fastq.gz"))
    print(f"{name}: .gz bytes identical to the oracle: {same}; decompressed content identical: {content}")

# ---- the sibling tools on the same inputs (reads/s; ReadQC and FastqTrim of ngs-bits have no CPU restatement with a command line here)
if os.environ.get("SPG_CLI_TOOLS", "1") != "0":
    bin_dir = os.path.join(ROOT, "ngs-bits_b200", "bin")
    tools = {
        "readqc_b200 gz in, paired": ([f"{bin_dir}/readqc_b200", "-in1", f"{d}/in1.fastq.gz", "-in2", f"{d}/in2.fastq.gz", "-txt", "-out", f"{d}/qc1.txt"], 2 * n),
        f"readqc_b200 bgzf in, paired, -threads {threads}": ([f"{bin_dir}/readqc_b200", "-in1", bgz[1], "-in2", bgz[2], "-txt", "-out", f"{d}/qc2.txt", "-threads", str(threads)], 2 * n),
        "fastqtrim_b200 gz in -> gz out": ([f"{bin_dir}/fastqtrim_b200", "-in", f"{d}/in1.fastq.gz", "-out", f"{d}/ft1.fastq.gz", "-start", "5", "-end", "5"], n),
        f"fastqtrim_b200 bgzf in -> bgzf out, -threads {threads}": ([f"{bin_dir}/fastqtrim_b200", "-in", bgz[1], "-out", f"{d}/ft2.fastq.gz", "-start", "5", "-end", "5", "-threads", str(threads), "-bgzf"], n),
    }
    for name, (cmd, reads) in tools.items():
        t0 = time.perf_counter()
        subprocess.run(cmd, check=True)
        el = time.perf_counter() - t0
        print(f"{name}: {el:.2f} s for {reads} reads = {reads / el / 1e6:.3f} Mreads/s end to end", flush=True)
    print("readqc outputs identical:", open(f"{d}/qc1.txt").read() == open(f"{d}/qc2.txt").read(), "| fastqtrim outputs identical:",
          subprocess.run(f"bash -c 'cmp <(gzip -dc {d}/ft1.fastq.gz) <(gzip -dc {d}/ft2.fastq.gz)'", shell=True).returncode == 0)
    print(open(f"{d}/qc1.txt").read())
