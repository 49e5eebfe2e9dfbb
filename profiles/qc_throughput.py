#!/usr/bin/env python
"""Throughput of the -qc statistics kernel (spg::qc_kernel) on one GPU: device-resident synthetic batches, CUDA-event timing.
usage: python profiles/qc_throughput.py [pairs] [read_len]   -> one JSON line (Mpairs/s, GB/s of algorithmic bytes 4*L+4 per pair)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ngs-bits_b200"))
import torch

import __graft_entry__ as g

g.build()
import seqpurge_b200 as sp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
stride = L + (L & 1)
dev = torch.device("cuda:0")
cfg = sp.SynthConfig(read_len=L)
bufs = []
for b in range(2):  # 2 x 2.4 GB at the defaults: larger than L2, alternated between launches
    t = {k: torch.empty((n, stride), dtype=torch.uint8, device=dev) for k in ("bases1", "quals1", "bases2", "quals2")}
    l1 = torch.empty(n, dtype=torch.int16, device=dev)
    l2 = torch.empty(n, dtype=torch.int16, device=dev)
    sp.synth_device(cfg, b * n, n, t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2)
    bufs.append((t, l1, l2))
eng = sp.Engine(sp.TrimmingParameters(qc=int(os.environ.get("SPG_QC_FLAGS", "1"))), devices=(0,))
if os.environ.get("SPG_QC_WARP"):  # the warp-per-pair form of the kernel (round 1)
    eng.set_option(sp.OPT_KERNEL, sp.KERNEL_WARP_PER_PAIR)


def run(i):
    t, l1, l2 = bufs[i % 2]
    eng.qc_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2)


for i in range(3):
    run(i)
torch.cuda.synchronize()
reps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    run(i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
st = eng.qc_stats()
assert st["reads_forward"] == (reps + 3) * n and st["errors"] == 0
print(json.dumps({"kernel": "qc_kernel (warp per pair)" if os.environ.get("SPG_QC_WARP") else "qc_lanes_kernel", "qc_flags": int(os.environ.get("SPG_QC_FLAGS", "1")), "pairs": n, "read_len": L, "ms_per_launch": round(ms, 4), "Mpairs_per_s": round(n / ms / 1e3, 1),
                  "GB_per_s": round(n * (4 * L + 4) / ms / 1e6, 1)}))
