#!/usr/bin/env python
"""Launch-geometry sweep of the trimming kernel on one GPU (device-resident batches, CUDA-event timing).
usage: python profiles/sweep.py [pairs] [read_len]   -> one line per (consumer warps, tile pairs, stages, CTAs/SM cap)"""
import itertools
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
stride = (L + 15) // 16 * 16
dev = torch.device("cuda:0")
workload = os.environ.get("SPG_SWEEP_CONFIG", "C2" if L == 150 else "C3")
params = sp.TrimmingParameters()
if workload == "C2":
    cfg = sp.SynthConfig(read_len=L)
elif workload == "C3":  # high overlap: every pair has an insert match
    cfg = sp.SynthConfig(read_len=L, insert_mean=150, insert_sd=40, insert_max=L - 1)
elif workload == "C4":  # 2 % errors, long low-quality tails, N runs
    cfg = sp.SynthConfig(read_len=L, error_rate=0.02, lowq_tail_mean=20.0, n_run_rate=0.005)
else:  # C5: NovaSeq-like, binned qualities
    cfg = sp.SynthConfig(read_len=L, insert_mean=350, insert_sd=100, error_rate=0.002, lowq_tail_mean=2.0, binned_quals=True)
print("workload", workload, flush=True)
bufs = []
for b in range(2):
    t = {k: torch.empty((n, stride), dtype=torch.uint8, device=dev) for k in ("bases1", "quals1", "bases2", "quals2")}
    l1 = torch.empty(n, dtype=torch.int16, device=dev)
    l2 = torch.empty(n, dtype=torch.int16, device=dev)
    sp.synth_device(cfg, b * n, n, t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2)
    bufs.append((t, l1, l2))
res = torch.empty((n, 8), dtype=torch.uint8, device=dev)
eng = sp.Engine(params, devices=(0,))
eng.set_option(sp.OPT_FULL_LEN, int(os.environ.get("SPG_SWEEP_FULL", L)))  # kernel variant compiled for this read length (0: general kernel)
grid = list(itertools.product((2, 3, 4), (16, 32), (2, 3), (0,)))
if len(sys.argv) > 3:
    grid = [tuple(int(x) for x in c.split(",")) for c in sys.argv[3:]]
ref = None
for cw, tp, ns, cap in grid:
    try:
        eng.set_option(sp.OPT_MIN_BLOCKS, cw)
        eng.set_option(sp.OPT_TILE_PAIRS, tp)
        eng.set_option(sp.OPT_STAGES, ns)
        eng.set_option(sp.OPT_GRID_CTAS_PER_SM, cap)

        def run(i):
            t, l1, l2 = bufs[i % 2]
            eng.trim_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, res)

        for i in range(2):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(6):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 6
        chk = int(res.view(torch.int64).sum().item())
        if ref is None:
            ref = chk
        print(f"minb={cw} tile={tp} stages={ns} cap={cap}: {ms:.3f} ms  {n / ms / 1e3:.1f} Mpairs/s  {'ok' if chk == ref else 'RESULT MISMATCH'}", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"minb={cw} tile={tp} stages={ns} cap={cap}: failed: {ex}", flush=True)
