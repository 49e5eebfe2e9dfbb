"""GPU parity of the -qc raw-read statistics: spg::qc_kernel through the C ABI (spg_qc_device, and spg_submit with
spg_params.qc) against the oracle's restatement of StatisticsReads::update (src/cppNGS/StatisticsReads.cpp:30-82), accumulator by
accumulator, and the CLI's qcML against the reference's golden file the way the reference's own test compares it
(src/tools-TEST/SeqPurge_Test.cpp:99-113).
"""
import os
import re
import subprocess

import numpy as np
import pytest

import helpers as H
from test_gpu_parity import sp  # noqa: F401  (module-scoped fixture: builds and imports the CUDA library)

pytestmark = pytest.mark.gpu

G = H.GOLDEN
FIELDS = ("reads_forward", "reads_reverse", "bases_sequenced", "read_q20", "base_q20", "base_q30")
ARRAYS = ("read_lengths", "pileup", "qsum_forward", "qsum_reverse", "base_qualities", "read_qualities", "qscore_dist_forward", "qscore_dist_reverse")


def assert_qc_equal(got, want):
    for k in FIELDS:
        assert got[k] == want[k], k
    for k in ARRAYS:
        if not np.array_equal(got[k], want[k]):
            bad = np.argwhere(got[k] != want[k])
            raise AssertionError(f"{k}: {len(bad)} entries differ, first at {bad[0]}: gpu={got[k][tuple(bad[0])]} oracle={want[k][tuple(bad[0])]}")
    assert got["errors"] == 0 and want["errors"] == 0


# spg_params.qc = 5: the statistics (1) with the histograms behind the qcML plots (+4), so that every accumulator is compared


def qc_via_submit(sp, batch, chunk=None, n_slots=2, **params):  # noqa: F811
    """-qc on the slot path: H2D, qc_kernel on the raw reads, then the trimming kernel (which may edit them with -ec)."""
    chunk = chunk or batch.n
    eng = sp.Engine(sp.TrimmingParameters(qc=5, **params), devices=(0,), n_slots=n_slots, max_pairs=chunk, max_len=min(batch.stride, 999))
    out = np.zeros(batch.n, sp.RESULT_DTYPE)
    pending = []
    for k, st in enumerate(range(0, batch.n, chunk)):
        slot = k % n_slots
        if len(pending) == n_slots:
            s0, st0, n0 = pending.pop(0)
            out[st0 : st0 + n0] = eng.wait(s0)
        n = min(chunk, batch.n - st)
        s = eng.slot(slot)
        for name in ("bases1", "quals1", "bases2", "quals2"):
            getattr(s, name)[:n] = getattr(batch, name)[st : st + n]
        s.len1[:n] = batch.len1[st : st + n]
        s.len2[:n] = batch.len2[st : st + n]
        eng.submit(slot, n)
        pending.append((slot, st, n))
    for s0, st0, n0 in pending:
        out[st0 : st0 + n0] = eng.wait(s0)
    got = eng.qc_stats()
    eng.close()
    return got, out


def qc_via_device(sp, batch):  # noqa: F811
    import torch

    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(getattr(batch, k)).to(dev) for k in ("bases1", "quals1", "bases2", "quals2")}
    l1 = torch.from_numpy(batch.len1.view(np.int16)).to(dev)
    l2 = torch.from_numpy(batch.len2.view(np.int16)).to(dev)
    eng = sp.Engine(sp.TrimmingParameters(qc=5), devices=(0,))
    eng.qc_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, n_pairs=batch.n)  # Batch rows are padded to a multiple of 8
    got = eng.qc_stats()
    eng.close()
    return got


def test_golden_fixture_matches_oracle_and_reference_qcml(sp):  # noqa: F811
    batch = H.golden_batch(1, 2)
    want = H.oracle_qc(batch)
    got, recs = qc_via_submit(sp, batch, chunk=100, ncut=0, qcut=0)
    assert_qc_equal(got, want)
    # the trimming result is unchanged by -qc
    ref, _ = H.oracle_trim(batch, ncut=0, qcut=0)
    assert np.array_equal(recs.view(np.uint64), ref.view(np.uint64))
    from test_oracle_golden import _golden_qc_values

    assert H.qc_metrics(got) == _golden_qc_values()


@pytest.mark.parametrize("i1,i2", [(3, 4), (5, 6), (7, 8), (9, 10), (11, 12)])
def test_other_fixtures_device_path(sp, i1, i2):  # noqa: F811
    batch = H.golden_batch(i1, i2)
    assert_qc_equal(qc_via_device(sp, batch), H.oracle_qc(batch))


@pytest.mark.parametrize("L,stride,n,seed", [(150, 150, 5003, 1), (151, 160, 777, 2), (36, 36, 9, 3), (250, 256, 2001, 4), (300, 320, 1001, 5), (33, 48, 100, 6)])
def test_random_ragged_batches(sp, L, stride, n, seed):  # noqa: F811
    batch = H.random_batch(n, L, seed, ragged=True, n_runs=0.01, stride=stride)
    want = H.oracle_qc(batch)
    assert_qc_equal(qc_via_device(sp, batch), want)
    got, _ = qc_via_submit(sp, batch, chunk=max(1, n // 3))
    assert_qc_equal(got, want)


def test_qc_sees_the_reads_before_error_correction(sp):  # noqa: F811
    batch = H.golden_batch(1, 2)
    got, _ = qc_via_submit(sp, batch, chunk=500, ncut=0, qcut=0, ec=True)
    assert_qc_equal(got, H.oracle_qc(batch))


def test_long_reads_generic_kernel(sp):  # noqa: F811
    batch = H.random_batch(300, 700, 21, ragged=True, stride=1000)
    batch.set_pair(0, b"ACGTN" * 199 + b"ACGT", b"I" * 999, b"T" * 999, b"5" * 999)
    assert_qc_equal(qc_via_device(sp, batch), H.oracle_qc(batch))


def test_accumulates_over_calls_and_quality_extremes(sp):  # noqa: F811
    import torch

    rng = np.random.default_rng(5)
    batch = H.random_batch(4000, 100, 31, stride=112)
    # the whole usable quality range 0..94 ('!'..chr(127); a byte >= 128 is a negative char in the reference) and lower-case bases,
    # which Pileup::inc accepts
    q = rng.integers(33, 128, size=batch.quals1.shape, dtype=np.uint8)
    batch.quals1[:] = q
    batch.quals2[:] = q[::-1]
    batch.bases1[::7] = np.char.lower(batch.bases1[::7].view("S1")).view(np.uint8)
    want = H.oracle_qc(batch)
    assert want["errors"] == 0
    dev = torch.device("cuda:0")
    eng = sp.Engine(sp.TrimmingParameters(qc=5), devices=(0,))
    for st in range(0, batch.n, 1000):
        t = [torch.from_numpy(np.ascontiguousarray(getattr(batch, k)[st : st + 1000])).to(dev) for k in ("bases1", "quals1", "bases2", "quals2")]
        l1 = torch.from_numpy(batch.len1[st : st + 1000].view(np.int16)).to(dev)
        l2 = torch.from_numpy(batch.len2[st : st + 1000].view(np.int16)).to(dev)
        eng.qc_device(*t, l1, l2)
    got = eng.qc_stats()
    eng.close()
    assert_qc_equal(got, want)


def test_unknown_base_or_quality_sets_the_error_count(sp):  # noqa: F811
    for what in ("base", "quality_high", "quality_low"):
        batch = H.random_batch(64, 100, 41, stride=112)
        if what == "base":
            batch.bases2[5, 17] = ord("X")
        elif what == "quality_high":
            batch.quals1[9, 3] = 200  # negative as a signed char
        else:
            batch.quals1[9, 3] = 32
        assert H.oracle_qc(batch)["errors"] > 0
        assert qc_via_device(sp, batch)["errors"] > 0


def test_synthetic_config2_slice(sp):  # noqa: F811
    import torch

    cfg = sp.SynthConfig(read_len=150, insert_mean=350, insert_sd=100)
    n, stride = 200_000, 150
    dev = torch.device("cuda:0")
    t = {k: torch.empty((n, stride), dtype=torch.uint8, device=dev) for k in ("bases1", "quals1", "bases2", "quals2")}
    l1 = torch.empty(n, dtype=torch.int16, device=dev)
    l2 = torch.empty(n, dtype=torch.int16, device=dev)
    sp.synth_device(cfg, 0, n, t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2)
    eng = sp.Engine(sp.TrimmingParameters(qc=5), devices=(0,))
    eng.qc_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2)
    got = eng.qc_stats()
    eng.close()
    batch = H.Batch(n, stride)
    for k in t:
        getattr(batch, k)[:n] = t[k].cpu().numpy()
    batch.len1[:n] = l1.cpu().numpy().view(np.uint16)
    batch.len2[:n] = l2.cpu().numpy().view(np.uint16)
    assert_qc_equal(got, H.oracle_qc(batch))


def _qcml_lines(path):
    """The reference's comparison: drop the 'creation ', 'source file' and <binary> lines (SeqPurge_Test.cpp:108-112); here also the
    stylesheet block and the XML prolog lines that belong to it, which seqpurge_b200 does not embed (QcReport.h)."""
    keep = []
    with open(path, encoding="latin-1") as f:
        for line in f:
            if re.search(r"creation |source file|<binary>", line):
                continue
            keep.append(line.rstrip("\n"))
    text = "\n".join(keep)
    text = re.sub(r"<\?xml-stylesheet.*?\]>\n", "", text, flags=re.S)
    text = re.sub(r"  <xsl:stylesheet.*</xsl:stylesheet>\n", "", text, flags=re.S)
    return text.split("\n")


def test_cli_qcml_equals_reference_golden(sp, tmp_path):  # noqa: F811
    cli = os.path.join(H.ROOT, "ngs-bits_b200", "bin", "seqpurge_b200")
    subprocess.run(["make", "-s", "-C", os.path.join(H.ROOT, "ngs-bits_b200", "host")], check=True)
    qc = tmp_path / "out.qcML"
    cmd = [cli, "-in1", f"{G}/SeqPurge_in1.fastq.gz", "-in2", f"{G}/SeqPurge_in2.fastq.gz", "-out1", str(tmp_path / "o1.fastq.gz"), "-out2", str(tmp_path / "o2.fastq.gz"),
           "-ncut", "0", "-qcut", "0", "-min_len", "15", "-qc", str(qc), "-block_size", "100", "-block_prefetch", "1", "-summary", str(tmp_path / "s.txt")]
    subprocess.run(cmd, check=True)
    assert _qcml_lines(qc) == _qcml_lines(f"{G}/SeqPurge_out1.qcML")
    import gzip

    for mine, gold in (("o1", 1), ("o2", 2)):
        assert gzip.open(tmp_path / f"{mine}.fastq.gz").read() == gzip.open(f"{G}/SeqPurge_out{gold}.fastq.gz").read()


def test_plot_histograms_are_opt_in(sp):  # noqa: F811
    """spg_params.qc = 1 leaves the plot histograms alone (they cost an atomic per 32 bases); everything else is unchanged."""
    batch = H.random_batch(800, 150, seed=3)
    eng = sp.Engine(sp.TrimmingParameters(qc=1), devices=(0,), n_slots=1, max_pairs=batch.n, max_len=batch.stride)
    s = eng.slot(0)
    for name in ("bases1", "quals1", "bases2", "quals2"):
        getattr(s, name)[: batch.n] = getattr(batch, name)[: batch.n]
    s.len1[: batch.n] = batch.len1[: batch.n]
    s.len2[: batch.n] = batch.len2[: batch.n]
    eng.submit(0, batch.n)
    eng.wait(0)
    got, want = eng.qc_stats(), H.oracle_qc(batch)
    eng.close()
    for k in FIELDS:
        assert got[k] == want[k], k
    for k in ("read_lengths", "pileup", "qsum_forward", "qsum_reverse"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("base_qualities", "read_qualities", "qscore_dist_forward", "qscore_dist_reverse"):
        assert not got[k].any(), k
