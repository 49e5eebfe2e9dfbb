#!/usr/bin/env python
"""bench.py -- Mread-pairs/s trimmed, 2x150 bp synthetic pairs, on N B200 (one process per GPU; read pairs are independent,
so ranks shard the stream with no collective on the data path: weak scaling).

  python bench.py --gpus 1 --steps 10 --warmup 3            # this repo's CUDA engine
  python bench.py --impl reference --gpus 1 --steps 3 --warmup 1   # CPU baseline arm (oracle port of the reference, all host threads)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the trimming hot path over one batch of `--pairs-per-step` synthetic pairs. The headline is BASELINE
config 2 (C2: 2x150 bp, insert ~ N(250,80), default Illumina adapters, SeqPurge defaults -qcut 15 -ncut 7); configs 3, 4 and 5 are
measured the same way and reported under "configs". The batches are generated on the device and stay resident in HBM (a pool of
distinct batches, 6.0 GB each at the default size, so every launch reads inputs far larger than the 126 MB L2). Reported:
  value     whole-job Mpairs/s, inputs resident in HBM (kernel launches only), max over ranks of CUDA-event time
  e2e       same metric through the C ABI with HOST buffers: pinned slot -> H2D -> kernel -> D2H of the result records (spg_submit/spg_wait)
  e2e_roundrobin  (N > 1, rank 0 alone) ONE engine over all N devices, slots dealt round robin, retired in submission order, C5 data
  roofline  algorithmic bytes (2*(L1+L2)+8 = 608 B/pair) / mean kernel time, against the measured HBM copy bandwidth
  parity    after the timed region a random slice of >= 1 M pairs of every config is checked record by record against the CPU oracle
  cpu_baseline  the CPU oracle (restatement of the reference's AnalysisWorker, all host threads) on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ngs-bits_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "Mread-pairs/s trimmed, 2x150bp synthetic"
UNIT = "Mpairs/s"

# BASELINE.json configs 2-5: synthetic stream + SeqPurge flags. C1 (10 k pairs, CPU -threads 1) is a parity case.
CONFIGS = {
    "C2": dict(read_len=150, synth=dict(insert_mean=250.0, insert_sd=80.0, error_rate=0.001, n_rate=1e-4, lowq_tail_mean=3.0), params=dict(),
               desc="C2: synthetic 2x150bp pairs, insert~N(250,80), 0.1% substitutions, default Illumina adapters, SeqPurge defaults (-qcut 15 -ncut 7)"),
    "C3": dict(read_len=250, synth=dict(insert_mean=150.0, insert_sd=40.0, insert_max=249, error_rate=0.001, n_rate=1e-4, lowq_tail_mean=3.0), params=dict(),
               desc="C3: synthetic 2x250bp pairs, insert~N(150,40) < 250 (every pair overlaps, overlaps > 170 take the halving path), SeqPurge defaults"),
    "C4": dict(read_len=150, synth=dict(insert_mean=250.0, insert_sd=80.0, error_rate=0.02, n_rate=1e-4, lowq_tail_mean=20.0, n_run_rate=0.005),
               params=dict(qcut=15, ncut=7), desc="C4: synthetic 2x150bp pairs, 2% substitutions, low-quality tails of mean 20, N runs in 0.5% of reads, -qcut 15 -ncut 7"),
    "C5": dict(read_len=150, synth=dict(insert_mean=350.0, insert_sd=100.0, error_rate=0.002, n_rate=1e-4, lowq_tail_mean=2.0, binned_quals=True), params=dict(),
               desc="C5: synthetic 2x150bp NovaSeq-like pairs, insert~N(350,100), binned qualities, SeqPurge defaults"),
    # not a BASELINE config: long reads (the widest plane variant of the warp-per-pair kernel), for `--only L600`
    "L600": dict(read_len=600, synth=dict(insert_mean=500.0, insert_sd=200.0, error_rate=0.005, n_rate=1e-4, lowq_tail_mean=10.0), params=dict(),
                 desc="synthetic 2x600bp pairs, insert~N(500,200) (reads beyond the register-plane widths of round 1)"),
}


def b_alg(read_len):
    """Algorithmic bytes per pair: bases + qualities of both reads as FASTQ delivers them + one 8-byte result record."""
    return 2 * (read_len + read_len) + 8


def row_stride(read_len):
    return (read_len + 1) // 2 * 2  # read length rounded up to an even number


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def host_batch_from_device(np, H, tensors, l1, l2, first, n, stride):
    b = H.Batch(n, stride)
    for k in ("bases1", "quals1", "bases2", "quals2"):
        getattr(b, k)[:n] = tensors[k][first : first + n].cpu().numpy()
    b.len1[:n] = l1[first : first + n].cpu().numpy().view(np.uint16)
    b.len2[:n] = l2[first : first + n].cpu().numpy().view(np.uint16)
    return b


def time_oracle(H, batch, threads, **params):
    t0 = time.perf_counter()
    H.oracle_trim(batch, threads=threads, **params)
    return time.perf_counter() - t0


def host_threads():
    """Threads the CPU arm may use: the CPUs this process is allowed on (cgroup quota taken into account)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


def host_topology():
    """What the box lets this process use: NUMA nodes, allowed memory nodes (the 8-GPU e2e number is a host-feed number)."""
    out = {"cpus": host_threads()}
    try:
        out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
        out["mems_allowed"] = [ln.split()[1] for ln in open("/proc/self/status") if ln.startswith("Mems_allowed_list")][0]
    except Exception:
        pass
    return out


def best_oracle_threads(H, probe, **params):
    """The oracle's block-parallel pool does not always scale to every hardware thread of the box (memory-bound byte loops,
    SMT): time a probe at a few thread counts and use the fastest one, so the baseline is not handicapped."""
    n = host_threads()
    cands = sorted({n, max(1, n // 2), max(1, n // 4), min(n, 32), min(n, 16)}, reverse=True)
    best = (0.0, 1)
    for t in cands:
        rate = probe.n / time_oracle(H, probe, t, **params)
        if rate > best[0]:
            best = (rate, t)
    return best[1], best[0]


def git_head():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True, timeout=5).stdout.strip() or None
    except Exception:
        return None


def measured_traffic(kernel_name, pairs):
    """DRAM bytes per launch from this round's ncu capture of the same kernel on the same workload (profiles/traffic_r2.json, written by
    profiles/ncu_summary.py --traffic from `ncu --set full`); None if the capture is of another kernel than the one that just ran."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic_r2.json")))
        short = kernel_name.split("<")[0].split("::")[-1]
        if short not in t.get("kernel", ""):
            return None, None
        return t["dram_bytes_per_pair"] * pairs, {"dram_bytes_per_pair": t["dram_bytes_per_pair"], "capture": t.get("capture"), "commit": t.get("commit"), "kernel": t.get("kernel")}
    except Exception:
        return None, None


_REAL_STDOUT = None


def emit(line):
    """The one JSON line of this run, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries write to fd 1 on their own (NCCL prints its version line there under torchrun): keep stdout for the JSON line only
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-step", type=int, default=10_000_000)
    ap.add_argument("--pool", type=int, default=10, help="distinct resident batches of the headline config (pool*pairs_per_step = the 100M-pair config by default)")
    ap.add_argument("--configs", default="C3,C4,C5,qc", help="further BASELINE configs measured after the headline (comma separated, '' = none); qc = the -qc statistics kernel")
    ap.add_argument("--config-steps", type=int, default=5)
    ap.add_argument("--config-pool", type=int, default=3)
    ap.add_argument("--parity-pairs", type=int, default=1_000_000, help="pairs of every config checked against the CPU oracle after timing (0 = off)")
    ap.add_argument("--e2e-pairs", type=int, default=1_000_000, help="pairs per pinned slot for the end-to-end measurement")
    ap.add_argument("--no-qual-tails", action="store_true", help="end to end without the slots' quality tails (SPG_OPT_QUAL_TAILS 0: the kernel fetches every trimming point over PCIe)")
    ap.add_argument("--e2e-slots", type=int, default=4)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--only", default="", help="profiling aid: measure just this config (C2..C5) device-resident and print its result")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference" and rank != 0:
        return 0  # the CPU arm runs on rank 0 only

    import numpy as np

    import __graft_entry__ as g

    g.build()
    import helpers as H

    head = CONFIGS["C2"]

    # ------------------------------------------------------------------------------------------------ reference (CPU) arm
    if args.impl == "reference":
        # no CUDA library on this arm: the input comes from the numpy generator of tests/helpers.py (same model as the device generator)
        L = head["read_len"]
        probe = H.synth_batch_numpy(200_000, L, seed=1, stride=row_stride(L), **head["synth"])
        threads, rate = best_oracle_threads(H, probe, **head["params"])
        budget = 150.0 / max(1, args.steps + args.warmup)  # whole run within a few minutes
        n = int(max(20_000, min(rate * min(budget, 10.0), 4_000_000))) // 8 * 8
        # the per-pair cost of the oracle does not depend on which pairs follow which: a sample of n pairs is 400 k generated pairs repeated
        base = H.synth_batch_numpy(min(n, 400_000), L, seed=2, stride=row_stride(L), **head["synth"])
        sample = H.Batch(n, base.stride)
        for k in ("bases1", "quals1", "bases2", "quals2", "len1", "len2"):
            src, dst = getattr(base, k), getattr(sample, k)
            for o in range(0, n, base.n):
                m = min(base.n, n - o)
                dst[o : o + m] = src[:m]
        for _ in range(args.warmup):
            time_oracle(H, sample, threads, **head["params"])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            time_oracle(H, sample, threads, **head["params"])
        el = time.perf_counter() - t0
        val = args.steps * n / el / 1e6
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": head["desc"], "pairs_per_step": n, "note": "CPU oracle (C restatement of AnalysisWorker::run, validated on the reference's 23 golden files); "
                       "the Qt reference itself cannot be built in this image; input from the numpy generator (no CUDA library on this arm)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "host_cpus": host_threads(),
                             "sample": f"{n} pairs per step of the same synthetic model, in-memory SoA batches, {threads} threads (fastest of a thread-count probe)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit(line)
        return 0

    # ------------------------------------------------------------------------------------------------ this repo's engine
    import torch

    import seqpurge_b200 as sp
    from seqpurge_b200 import sharding

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path in the product)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    cpu_group = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")  # host-side barrier while rank 0 drives all devices (no NCCL kernel spinning on them)

    def alloc(n, stride):
        t = {k: torch.empty((n, stride), dtype=torch.uint8, device=dev) for k in ("bases1", "quals1", "bases2", "quals2")}
        return t, torch.empty(n, dtype=torch.int16, device=dev), torch.empty(n, dtype=torch.int16, device=dev)

    total_launches = 0
    parity_threads = max(1, host_threads() // world)

    def measure(name, steps, warmup, B, pool_n, clocks_wanted):
        """Device-resident throughput of one config + the parity gate. Returns (result dict, pool of batches, engine)."""
        nonlocal total_launches
        c = CONFIGS[name]
        L, stride = c["read_len"], row_stride(c["read_len"])
        cfg = sp.SynthConfig(read_len=L, **c["synth"])
        eng = sp.Engine(sp.TrimmingParameters(**c["params"]), devices=(local_rank,), n_slots=0, max_pairs=1, max_len=L)
        pool = []
        shard = sharding.shard_batches(pool_n, B, rank, world)  # every rank trims its own contiguous shard of the stream
        for first in shard.first_pair:
            t, l1, l2 = alloc(B, stride)
            sp.synth_device(cfg, first, B, t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, device_id=local_rank)
            pool.append((t, l1, l2))
        res = torch.empty((B, 8), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()

        def step(i):
            t, l1, l2 = pool[i % pool_n]
            eng.trim_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, res, device_index=0)

        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        launches0 = eng.launch_count
        sampler = ClockSampler(local_rank) if clocks_wanted else None
        if sampler:
            sampler.start()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        torch.cuda.synchronize()
        evs[0].record()
        for i in range(steps):
            step(warmup + i)
            evs[i + 1].record()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        clocks = sampler.finish() if sampler else None
        total_ms = evs[0].elapsed_time(evs[-1])
        kernel_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        launches = eng.launch_count - launches0
        total_launches += launches
        total_ms = sharding.reduce_max(total_ms, dist, dev)  # max over ranks of the device-side time
        value = sharding.aggregate_throughput(steps * B, world, total_ms * 1e-3) / 1e6
        kernel = eng.last_kernel

        # sanity on the last result: every record must be status 0 with plausible lengths
        chk = sp.results_from_tensor(res[:100000])
        assert (chk["status"] == 0).all() and (chk["len1"] <= L).all()
        frac_insert = float((chk["flags"] & 1).mean())

        # parity gate: a random slice of the pool, trimmed again by the kernel that was just timed, against the CPU oracle on the same bytes
        parity = None
        if args.parity_pairs > 0:
            rng = np.random.default_rng(1234 + rank)
            m = max(min(args.parity_pairs // world, B), min(100_000, B)) // 8 * 8  # N ranks check N different slices (>= parity_pairs together)
            bi = int(rng.integers(0, pool_n))
            st = int(rng.integers(0, B - m + 1)) // 8 * 8
            t, l1, l2 = pool[bi]
            eng.trim_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, res, device_index=0)
            torch.cuda.synchronize()
            total_launches += 1
            got = sp.results_from_tensor(res[st : st + m]).copy()
            batch = host_batch_from_device(np, H, t, l1, l2, st, m, stride)
            t0 = time.perf_counter()
            want, _ = H.oracle_trim(batch, threads=parity_threads, **c["params"])
            mism = int((got.view(np.uint64) != want.view(np.uint64)).sum())
            parity = {"pairs": m, "mismatches": mism, "batch": int(shard.batches[bi]), "first_pair": int(shard.first_pair[bi] + st),
                      "oracle_seconds": round(time.perf_counter() - t0, 2), "oracle_threads": parity_threads}
            assert mism == 0, f"{name}: {mism} of {m} result records differ from the CPU oracle"

        peak, peak_src = hbm_peak()
        mean_kernel_ms = sum(kernel_ms) / len(kernel_ms)
        achieved = B * b_alg(L) / (mean_kernel_ms * 1e-3) / 1e9
        out = {"workload": c["desc"], "value": value, "unit": UNIT, "steps": steps, "warmup": warmup, "pairs_per_step": B, "resident_batches": pool_n,
               "read_len": L, "row_stride": stride, "ms_per_step": total_ms / steps, "kernel": kernel, "insert_hit_fraction": frac_insert, "parity": parity,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                            "algorithmic_bytes_per_pair": b_alg(L), "kernel_ms_mean": mean_kernel_ms, "kernel": kernel}}
        if clocks is not None:
            out["clocks"] = clocks
        return out, pool, eng

    B = (args.pairs_per_step + 7) // 8 * 8
    pool_n = max(1, min(args.pool, args.steps + args.warmup))
    if args.only:
        r, _, _ = measure(args.only, args.steps, args.warmup, B if CONFIGS[args.only]["read_len"] <= 160 else B // 2, pool_n, True)
        if rank == 0:
            emit(r)
        return 0
    headline, pool, eng0 = measure("C2", args.steps, args.warmup, B, pool_n, True)
    L0, stride0 = head["read_len"], row_stride(head["read_len"])

    # ---- end to end through the C ABI: pinned host slots -> H2D -> kernel -> D2H results (one engine per rank, its own GPU)
    e2e = None
    if not args.no_e2e:
        ns, n_e = args.e2e_slots, args.e2e_pairs
        eng = sp.Engine(sp.TrimmingParameters(**head["params"]), devices=(local_rank,), n_slots=ns, max_pairs=n_e, max_len=L0)
        tails = not args.no_qual_tails
        eng.set_option(sp.OPT_QUAL_TAILS, 1 if tails else 0)
        t, l1, l2 = pool[0]
        for s in range(ns):
            sl = eng.slot(s)
            o = (s * n_e) % max(1, B - n_e)
            for k in ("bases1", "quals1", "bases2", "quals2"):
                getattr(sl, k)[:n_e] = t[k][o : o + n_e].cpu().numpy()
            sl.len1[:n_e] = l1[o : o + n_e].cpu().numpy().view(np.uint16)
            sl.len2[:n_e] = l2[o : o + n_e].cpu().numpy().view(np.uint16)
            if tails:
                sl.fill_qtails(n_e)

        def e2e_step():
            for s in range(ns):
                eng.submit(s, n_e)
            for s in range(ns):
                eng.wait(s)

        for _ in range(3):
            e2e_step()
        if dist:
            dist.barrier()
        e_steps = max(3, min(args.steps, 10))
        l0 = eng.launch_count
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        el = time.perf_counter() - t0
        total_launches += eng.launch_count - l0
        el = sharding.reduce_max(el, dist, dev)
        zero_copy = eng.zero_copy_quals
        copied = ns * (n_e * (2 if zero_copy else 4) * stride0 + 2 * 2 * n_e + (2 * sp.QTAIL * n_e if zero_copy and tails else 0))
        # parity of the end-to-end path: the records the last step left in slot 0 against the oracle on the slot's own rows
        e2e_parity = None
        if rank == 0 and args.parity_pairs > 0:
            sl = eng.slot(0)
            npar = min(n_e, args.parity_pairs)
            eng.submit(0, n_e)
            got = eng.wait(0)[:npar].copy()
            hb = H.Batch(npar, stride0)
            for k in ("bases1", "quals1", "bases2", "quals2"):
                getattr(hb, k)[:npar] = getattr(sl, k)[:npar]
            hb.len1[:npar] = sl.len1[:npar]
            hb.len2[:npar] = sl.len2[:npar]
            want, _ = H.oracle_trim(hb, threads=host_threads(), **head["params"])
            mism = int((got.view(np.uint64) != want.view(np.uint64)).sum())
            e2e_parity = {"pairs": npar, "mismatches": mism}
            assert mism == 0, f"end-to-end records differ from the oracle in {mism} of {npar} pairs"
        e2e = {"value": sharding.aggregate_throughput(e_steps * ns * n_e, world, el) / 1e6, "unit": UNIT, "h2d_bytes_per_step": copied,
               "d2h_bytes_per_step": ns * n_e * 8, "steps": e_steps, "pairs_per_step": ns * n_e, "kernel": eng.last_kernel,
               "h2d_copied_gbs_per_gpu": copied * e_steps / el / 1e9,
               "boundary": "spg_submit/spg_wait on pinned host SoA slots (ASCII rows as FASTQ delivers them), wall clock incl. H2D + kernel + D2H; the slots are filled once "
                           "and resubmitted every step (the E boundary of SURVEY.md 8d starts at the filled slot)",
               "host": host_topology()}
        if e2e_parity:
            e2e["parity"] = e2e_parity
        if zero_copy:
            e2e["zero_copy"] = ("the two quality planes stay in the pinned slot: the kernel reads the sectors that hold the trimming points over PCIe itself; "
                                "h2d_bytes_per_step counts what is copied (bases, lengths" + (", quality tails" if tails else "") + ")")
            e2e["qual_tails"] = (f"SPG_OPT_QUAL_TAILS: the stager also writes the last {sp.QTAIL} qualities of every read into the slot (2 x {sp.QTAIL} B per pair, copied with the bases); "
                                 "only reads cut by the adapter steps or trimmed deeper than 12 windows go to their quality row over PCIe") if tails else "off"
        eng.close()

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        t, l1, l2 = pool[0]
        probe = host_batch_from_device(np, H, t, l1, l2, 0, 200_000, stride0)
        threads, rate = best_oracle_threads(H, probe, **head["params"])
        n = int(max(40_000, min(rate * args.cpu_seconds, B, 8_000_000))) // 8 * 8
        sample = host_batch_from_device(np, H, t, l1, l2, 0, n, stride0)
        el = time_oracle(H, sample, threads, **head["params"])
        cpu = {"value": n / el / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
               "host_cpus": host_threads(),
               "sample": f"first {n} pairs of batch 0 of the same stream, in-memory SoA, oracle/ (C restatement of the reference) with {threads} threads "
                         f"(fastest of a thread-count probe), {el:.1f} s"}
    del pool
    eng0.close()
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs, same measurement, smaller pools
    others = {}
    for name in [c for c in args.configs.split(",") if c and c != "qc"]:
        c = CONFIGS[name]
        Bc = B if c["read_len"] <= 160 else B // 2  # config 3 is 50 M pairs of 2x250: half the pairs per batch
        r, p2, e2 = measure(name, args.config_steps, 3, Bc, args.config_pool, False)
        others[name] = r
        del p2
        e2.close()
        torch.cuda.empty_cache()

    # ---- the -qc statistics kernel on the headline workload: device-resident throughput + every accumulator against the oracle
    if "qc" in args.configs.split(","):
        c = CONFIGS["C2"]
        L, stride = c["read_len"], row_stride(c["read_len"])
        Bq = B
        engq = sp.Engine(sp.TrimmingParameters(qc=1), devices=(local_rank,), n_slots=0, max_pairs=1, max_len=L)
        cfg = sp.SynthConfig(read_len=L, **c["synth"])
        poolq = []
        for b in range(2):
            t, l1, l2 = alloc(Bq, stride)
            sp.synth_device(cfg, (rank * 2 + b) * Bq, Bq, t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, device_id=local_rank)
            poolq.append((t, l1, l2))
        torch.cuda.synchronize()

        def qstep(i):
            t, l1, l2 = poolq[i % 2]
            engq.qc_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2)

        for i in range(3):
            qstep(i)
        torch.cuda.synchronize()
        qs = args.config_steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = engq.launch_count
        e0.record()
        for i in range(qs):
            qstep(i)
        e1.record()
        torch.cuda.synchronize()
        total_launches += engq.launch_count - l0
        ms = sharding.reduce_max(e0.elapsed_time(e1), dist, dev) / qs
        engq.close()
        # parity: a fresh context over a slice, every accumulator against the oracle's restatement of StatisticsReads::update
        m = 200_000
        engp = sp.Engine(sp.TrimmingParameters(qc=5), devices=(local_rank,), n_slots=0, max_pairs=1, max_len=L)
        t, l1, l2 = poolq[0]
        engp.qc_device(t["bases1"][:m], t["quals1"][:m], t["bases2"][:m], t["quals2"][:m], l1[:m], l2[:m], n_pairs=m)
        got = engp.qc_stats()
        engp.close()
        want = H.oracle_qc(host_batch_from_device(np, H, t, l1, l2, 0, m, stride))
        bad = [k for k in want if not np.array_equal(np.asarray(got[k]), np.asarray(want[k]))]
        assert not bad, f"-qc statistics differ from the oracle: {bad}"
        peak, peak_src = hbm_peak()
        balg = 4 * L + 4
        ach = Bq * balg / (ms * 1e-3) / 1e9
        others["qc"] = {"workload": "-qc raw-read statistics (StatisticsReads::update of both reads) of the C2 batches, device resident", "kernel": "spg::qc_lanes_kernel",
                        "value": world * Bq / ms / 1e3, "unit": UNIT, "steps": qs, "pairs_per_step": Bq, "ms_per_step": ms,
                        "parity": {"pairs": m, "accumulators": len(want), "mismatches": 0},
                        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "algorithmic_bytes_per_pair": balg,
                                     "kernel_ms_mean": ms}}
        del poolq
        torch.cuda.empty_cache()

    # ---- N > 1: the north-star split. ONE host process (rank 0) drives all N devices: slots dealt round robin (slot % N), retired in
    # submission order like ThreadCoordinator hands jobs to the writer; the other ranks stay off their GPUs meanwhile.
    rr = None
    if world > 1 and not args.no_e2e:
        dist.barrier(group=cpu_group)
        if rank == 0:
            c = CONFIGS["C5"]
            L5, s5 = c["read_len"], row_stride(c["read_len"])
            n_e = args.e2e_pairs
            ns = (4 if world <= 2 else 2) * world  # pinned slots of 1 M pairs each: 2 per device are enough to keep copies and kernels overlapped
            eng = sp.Engine(sp.TrimmingParameters(**c["params"]), devices=tuple(range(world)), n_slots=ns, max_pairs=n_e, max_len=L5)
            eng.set_option(sp.OPT_QUAL_TAILS, 0 if args.no_qual_tails else 1)
            cfg = sp.SynthConfig(read_len=L5, **c["synth"])
            t, l1, l2 = alloc(n_e, s5)
            for s in range(ns):
                sp.synth_device(cfg, 500_000_000 - (s + 1) * n_e, n_e, t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, device_id=local_rank)
                torch.cuda.synchronize()
                sl = eng.slot(s)
                for k in ("bases1", "quals1", "bases2", "quals2"):
                    getattr(sl, k)[:n_e] = t[k][:n_e].cpu().numpy()
                sl.len1[:n_e] = l1[:n_e].cpu().numpy().view(np.uint16)
                sl.len2[:n_e] = l2[:n_e].cpu().numpy().view(np.uint16)
                if not args.no_qual_tails:
                    sl.fill_qtails(n_e)
            # parity of the multi-device path: the records of slot ns-1 (device (ns-1) % N) against the oracle
            rounds = max(3, min(args.steps, 10))
            checks = []

            def ring(n_rounds, keep=False):
                inflight = []
                for k in range(n_rounds * ns):
                    s = k % ns
                    if len(inflight) == ns:  # retire the oldest first: in-order retirement
                        r = eng.wait(inflight.pop(0))
                        if keep and len(checks) < ns:
                            checks.append(r.copy())
                    eng.submit(s, n_e)
                    inflight.append(s)
                for s in inflight:
                    r = eng.wait(s)
                    if keep and len(checks) < ns:
                        checks.append(r.copy())

            ring(2)
            l0 = eng.launch_count
            t0 = time.perf_counter()
            ring(rounds)
            el = time.perf_counter() - t0
            total_launches += eng.launch_count - l0
            ring(1, keep=True)
            mism = 0
            checked = 0
            for s in (0, ns - 1):  # one slot of the first and one of the last device against the oracle
                sl = eng.slot(s)
                m = min(n_e, 250_000)
                b = H.Batch(m, s5)
                for k in ("bases1", "quals1", "bases2", "quals2"):
                    getattr(b, k)[:m] = getattr(sl, k)[:m]
                b.len1[:m] = sl.len1[:m]
                b.len2[:m] = sl.len2[:m]
                want, _ = H.oracle_trim(b, threads=host_threads(), **c["params"])
                mism += int((checks[s][:m].view(np.uint64) != want.view(np.uint64)).sum())
                checked += m
            assert mism == 0, f"round-robin engine: {mism} records differ from the oracle"
            zc = eng.zero_copy_quals
            copied = n_e * (2 if zc else 4) * s5 + 2 * 2 * n_e + (2 * sp.QTAIL * n_e if zc and not args.no_qual_tails else 0)
            rr = {"value": rounds * ns * n_e / el / 1e6, "unit": UNIT, "workload": c["desc"], "devices": world, "slots": ns, "pairs_per_slot": n_e, "rounds": rounds,
                  "h2d_bytes_per_slot": copied, "h2d_copied_gbs_total": rounds * ns * copied / el / 1e9, "kernel": eng.last_kernel,
                  "parity": {"pairs": checked, "mismatches": mism}, "qual_tails": not args.no_qual_tails,
                  "boundary": "ONE process, spg_create over all devices, slot s on device s % N, spg_submit ring with in-order spg_wait (what ThreadCoordinator::analyze becomes); "
                              "the other ranks idle on a host-side barrier",
                  "host": host_topology()}
            eng.close()
        dist.barrier(group=cpu_group)

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0

    roof = dict(headline["roofline"])
    traffic, traffic_src = measured_traffic(headline["kernel"], B)
    roof["traffic"] = traffic
    roof["traffic_source"] = traffic_src
    roof["note"] = "integer-issue bound, not HBM bound (DESIGN.md section 4): DRAM traffic equals the algorithmic bytes, the kernel's ALU pipe is the busy one"
    line = {
        "metric": METRIC, "value": headline["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": headline["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": head["desc"], "pairs_per_step": B, "resident_batches": pool_n, "read_len": L0, "row_stride": stride0,
                   "l2": "inputs larger than L2 (each step reads a distinct 6.0 GB batch at the default size)", "parallelism": f"replicated engine x{world}, stream sharded by rank, no collective",
                   "insert_hit_fraction": headline["insert_hit_fraction"], "commit": git_head()},
        "roofline": roof, "parity": headline["parity"], "clocks": headline.get("clocks"), "gpu_launches": total_launches,
    }
    if others:
        line["configs"] = others
    if e2e:
        line["e2e"] = e2e
    if rr:
        line["e2e_roundrobin"] = rr
    if cpu:
        line["cpu_baseline"] = cpu
    emit(line)
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
