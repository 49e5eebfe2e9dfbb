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
