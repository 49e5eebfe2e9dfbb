"""Multi-GPU tests (need >= 2 CUDA devices; skipped otherwise): one context dealing slots round robin to two devices returns
the same records in the same order as one device, and the drop-in CLI with -gpus 0,1 reproduces the reference's golden files."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import __graft_entry__ as g

    g.build()
    import seqpurge_b200

    return seqpurge_b200


def test_round_robin_two_devices(sp):
    batch = H.golden_batch(5, 6)  # 10 526 pairs
    want, _ = H.oracle_trim(batch)
    eng = sp.Engine(sp.TrimmingParameters(), devices=(0, 1), n_slots=4, max_pairs=1500, max_len=batch.stride - 1)
    out = np.zeros(batch.n, sp.RESULT_DTYPE)
    inflight = []
    for k, st in enumerate(range(0, batch.n, 1500)):
        slot = k % 4
        if len(inflight) == 4:
            s0, st0, n0 = inflight.pop(0)
            out[st0 : st0 + n0] = eng.wait(s0)
        n = min(1500, batch.n - st)
        s = eng.slot(slot)
        for name in ("bases1", "quals1", "bases2", "quals2"):
            getattr(s, name)[:n] = getattr(batch, name)[st : st + n]
        s.len1[:n] = batch.len1[st : st + n]
        s.len2[:n] = batch.len2[st : st + n]
        eng.submit(slot, n)
        inflight.append((slot, st, n))
    for s0, st0, n0 in inflight:
        out[st0 : st0 + n0] = eng.wait(s0)
    eng.close()
    assert np.array_equal(out.view(np.uint64), want.view(np.uint64))


def test_cli_two_gpus_reproduces_golden(sp, tmp_path):
    G = H.GOLDEN
    cli = os.path.join(H.ROOT, "ngs-bits_b200", "bin", "seqpurge_b200")
    cmd = [cli, "-in1", f"{G}/SeqPurge_in5.fastq.gz", "-in2", f"{G}/SeqPurge_in6.fastq.gz", "-out1", str(tmp_path / "o1.fastq.gz"), "-out2", str(tmp_path / "o2.fastq.gz"),
           "-summary", str(tmp_path / "s.txt"), "-ncut", "0", "-qcut", "0", "-min_len", "15", "-block_size", "100", "-block_prefetch", "8", "-gpus", "0,1"]
    subprocess.run(cmd, check=True)
    for mine, gold in (("o1", 5), ("o2", 6)):
        with gzip.open(tmp_path / f"{mine}.fastq.gz", "rb") as a, gzip.open(f"{G}/SeqPurge_out{gold}.fastq.gz", "rb") as b:
            assert a.read() == b.read()


def test_slots_from_concurrent_threads():
    """The ABI is used per slot from different host threads (the reference's analysis pool runs jobs concurrently,
    ThreadCoordinator.cpp:40-42): four threads fill, submit and wait for their own slots in a loop; every result equals the oracle."""
    import threading

    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import __graft_entry__ as g

    g.build()
    import seqpurge_b200 as spm

    batch = H.golden_batch(1, 2)
    want, _ = H.oracle_trim(batch)
    n_threads, rounds, chunk = 4, 6, 600
    devices = (0, 1) if torch.cuda.device_count() >= 2 else (0,)
    eng = spm.Engine(spm.TrimmingParameters(), devices=devices, n_slots=n_threads, max_pairs=chunk, max_len=batch.stride - 1)
    errors = []

    def work(tid):
        try:
            s = eng.slot(tid)
            for r in range(rounds):
                st = ((tid * rounds + r) * 97) % (batch.n - chunk)
                for name in ("bases1", "quals1", "bases2", "quals2"):
                    getattr(s, name)[:chunk] = getattr(batch, name)[st : st + chunk]
                s.len1[:chunk] = batch.len1[st : st + chunk]
                s.len2[:chunk] = batch.len2[st : st + chunk]
                eng.submit(tid, chunk)
                got = eng.wait(tid)
                if not np.array_equal(got.view(np.uint64), want[st : st + chunk].view(np.uint64)):
                    errors.append((tid, r))
        except Exception as ex:  # noqa: BLE001
            errors.append((tid, repr(ex)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    eng.close()
    assert errors == []
