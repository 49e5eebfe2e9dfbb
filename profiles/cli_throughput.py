#!/usr/bin/env python
"""End-to-end command-line comparison (the X boundary of SURVEY.md 8d: gz in -> trim -> gz out, informational):
seqpurge_b200 (CUDA engine) vs the CPU oracle CLI with all host threads, same synthetic FASTQ.gz input, outputs compared byte for byte.
usage: python profiles/cli_throughput.py [pairs]"""
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
threads = len(os.sched_getaffinity(0))
runs = {
    "oracle_cli": [os.path.join(ROOT, "oracle", "build", "seqpurge_oracle"), "-threads", str(threads)],
    "seqpurge_b200": [os.path.join(ROOT, "ngs-bits_b200", "bin", "seqpurge_b200")],
}
for name, cmd in runs.items():
    os.makedirs(f"{d}/{name}")
    t0 = time.perf_counter()
    subprocess.run(cmd + ["-in1", f"{d}/in1.fastq.gz", "-in2", f"{d}/in2.fastq.gz", "-out1", f"{d}/{name}/o1.fastq.gz", "-out2", f"{d}/{name}/o2.fastq.gz",
                          "-summary", f"{d}/{name}/s.txt"], check=True)
    el = time.perf_counter() - t0
    print(f"{name}: {el:.2f} s for {n} pairs = {n / el / 1e6:.3f} Mpairs/s end to end (gz in, gz out)")
same = all(open(f"{d}/oracle_cli/{f}", "rb").read() == open(f"{d}/seqpurge_b200/{f}", "rb").read() for f in ("o1.fastq.gz", "o2.fastq.gz"))
print("outputs byte-identical:", same)
