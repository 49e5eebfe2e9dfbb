"""CPU tests that pin the oracle (oracle/) to the reference's own fixtures:
  * the 23 golden FASTQ outputs of the ten single-thread SeqPurge tool tests (src/tools-TEST/SeqPurge_Test.cpp:100-208),
    compared on decompressed content exactly like the reference's COMPARE_FILES (src/cppTFW/TestFramework.h:373-448);
  * the known answers of trimQuality / trimN (src/cppNGS-TEST/FastqFileStream_Test.cpp:9-128) and of
    factorial / matchProbability (src/cppCORE-TEST/BasicStatistics_Test.cpp:144-162).
"""
import ctypes as C
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

import helpers as H

G = H.GOLDEN
COMMON = ["-block_size", "100", "-block_prefetch", "1"]

# (test name, in1, in2, out1, out2, extra flags) -- flags copied from SeqPurge_Test.cpp
CASES = [
    ("test_01", 1, 2, 1, 2, ["-ncut", "0", "-qcut", "0", "-min_len", "15"]),
    ("test_02", 3, 4, 3, 4, ["-ncut", "0", "-qcut", "0", "-min_len", "15"]),
    ("test_03", 5, 6, 5, 6, ["-ncut", "0", "-qcut", "0", "-min_len", "15"]),
    ("test_04", 7, 8, 7, 8, ["-a1", "CTGTCTCTTATACACATCT", "-a2", "CTGTCTCTTATACACATCT", "-ncut", "0", "-qcut", "0", "-min_len", "15"]),
    ("test_05", 1, 2, 9, 10, ["-qcut", "15", "-ncut", "0", "-min_len", "15"]),
    ("test_06", 1, 2, 11, 12, ["-ncut", "7", "-qcut", "0", "-min_len", "15"]),
    ("test_07", 1, 2, 13, 14, ["-qcut", "25", "-out3", "OUT3"]),
    ("test_08", 9, 10, 16, 17, ["-min_len", "15"]),
    ("test_09", 11, 12, 18, 19, ["-min_len", "15"]),
    ("test_10", 1, 2, 20, 21, ["-ncut", "0", "-qcut", "0", "-ec", "-min_len", "15"]),
]


def _content(path):
    with gzip.open(path, "rb") as f:
        return f.read()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("threads", [1, 3])
def test_cli_reproduces_reference_goldens(case, threads, oracle_build, tmp_path):
    name, i1, i2, o1, o2, flags = case
    out1, out2 = tmp_path / "o1.fastq.gz", tmp_path / "o2.fastq.gz"
    flags = [str(tmp_path / "out15") if f == "OUT3" else f for f in flags]
    cmd = [os.path.join(oracle_build, "seqpurge_oracle"), "-in1", f"{G}/SeqPurge_in{i1}.fastq.gz", "-in2", f"{G}/SeqPurge_in{i2}.fastq.gz",
           "-out1", str(out1), "-out2", str(out2), "-summary", str(tmp_path / "summary.txt"), "-threads", str(threads)] + COMMON + flags
    subprocess.run(cmd, check=True)
    assert _content(out1) == _content(f"{G}/SeqPurge_out{o1}.fastq.gz")
    assert _content(out2) == _content(f"{G}/SeqPurge_out{o2}.fastq.gz")
    if name == "test_07":
        assert _content(tmp_path / "out15_R1.fastq.gz") == _content(f"{G}/SeqPurge_out15_R1.fastq.gz")
        assert _content(tmp_path / "out15_R2.fastq.gz") == _content(f"{G}/SeqPurge_out15_R2.fastq.gz")


def test_match_probability_known_answers():
    lib = H.oracle_lib()
    # BasicStatistics_Test.cpp:144-162 (F_EQUAL compares with 1e-5 tolerance)
    for n, want in [(0, 1.0), (1, 1.0), (2, 2.0), (3, 6.0), (4, 24.0)]:
        assert lib.spo_factorial(n) == want
    for args, want in [((0.1, 1, 1), 0.100), ((0.1, 1, 2), 0.190), ((0.1, 1, 3), 0.271), ((0.1, 1, 5), 0.40951), ((0.1, 5, 5), 0.00001)]:
        assert abs(lib.spo_match_probability(*args) - want) < 1e-5
    assert np.isfinite(lib.spo_factorial(170)) and np.isnan(lib.spo_factorial(171))
    # halving path: count > 170 (BasicStatistics.cpp:284-290)
    assert lib.spo_match_probability(0.25, 300, 300) == lib.spo_match_probability(0.25, 150, 150)


def _tq(quals, cutoff=15):
    lib = H.oracle_lib()
    n = C.c_int(len(quals))
    removed = lib.spo_trim_quality(quals, C.byref(n), cutoff, 5, 33)
    return removed, n.value


def _tn(bases, k=7):
    lib = H.oracle_lib()
    n = C.c_int(len(bases))
    removed = lib.spo_trim_n(bases, C.byref(n), k)
    return removed, n.value


def test_trim_quality_known_answers():
    # FastqFileStream_Test.cpp:9-67
    assert _tq(b"") == (0, 0)
    assert _tq(b"###") == (0, 3)
    assert _tq(b"IIIII") == (0, 5)
    assert _tq(b"#####") == (5, 0)
    assert _tq(b"I" * 32) == (0, 32)
    assert _tq(b"I" * 27 + b"#####") == (5, 27)
    assert _tq(b"?????????????????????:50+#######") == (8, 24)


def test_trim_n_known_answers():
    # FastqFileStream_Test.cpp:70-128
    assert _tn(b"") == (0, 0)
    assert _tn(b"ACG") == (0, 3)
    assert _tn(b"ACGTANNNNNN") == (0, 11)
    assert _tn(b"ACGTANNNNNNN") == (7, 5)
    assert _tn(b"ACGTANNNNNNANNNNNNN") == (7, 12)
    assert _tn(b"NNNNNNNACGTANNNNNNA") == (19, 0)
    assert _tn(b"ACGTANNNNNNNNNNNNNN") == (14, 5)


def test_batch_form_matches_cli_records(oracle_build, tmp_path):
    """The SoA batch entry point used by the GPU parity tests gives the same trimmed lengths as the CLI (test_05 flags)."""
    b = H.golden_batch(1, 2)
    rec, _ = H.oracle_trim(b, threads=2, qcut=15, ncut=0)
    out1 = tmp_path / "o1.fastq.gz"
    out2 = tmp_path / "o2.fastq.gz"
    subprocess.run([os.path.join(oracle_build, "seqpurge_oracle"), "-in1", f"{G}/SeqPurge_in1.fastq.gz", "-in2", f"{G}/SeqPurge_in2.fastq.gz", "-out1", str(out1),
                    "-out2", str(out2), "-summary", str(tmp_path / "s.txt"), "-qcut", "15", "-ncut", "0", "-min_len", "0"], check=True)
    got1 = [len(r[1]) for r in H.read_fastq(str(out1))]
    got2 = [len(r[1]) for r in H.read_fastq(str(out2))]
    assert got1 == rec["len1"].tolist() and got2 == rec["len2"].tolist()
    assert (rec["status"] == 0).all()


def _golden_qc_values():
    import re

    vals = []
    with open(f"{G}/SeqPurge_out1.qcML", encoding="latin-1") as f:
        for line in f:
            m = re.search(r'<qualityParameter ID="qp\d+" name="([^"]+)".* value="([^"]*)"', line)
            if m:
                vals.append((m.group(1), m.group(2)))
    return vals


def test_qc_statistics_match_the_reference_qcml():
    """-qc (SeqPurge_Test.cpp:99-113): the eight quality parameters of the reference's golden qcML from the oracle's restatement of
    StatisticsReads::update / getResult over the untrimmed reads of test_01."""
    want = _golden_qc_values()
    assert len(want) == 8
    d = H.oracle_qc(H.golden_batch(1, 2))
    assert d["errors"] == 0
    assert H.qc_metrics(d) == want


def test_qc_statistics_invariants():
    b = H.random_batch(500, 100, 11, ragged=True, n_runs=0.01)
    d = H.oracle_qc(b)
    assert d["reads_forward"] == d["reads_reverse"] == b.n
    assert d["bases_sequenced"] == int(b.len1.sum()) + int(b.len2.sum()) == int(d["pileup"].sum())
    assert int(d["read_lengths"].sum()) == 2 * b.n
    assert d["base_q30"] <= d["base_q20"] <= d["bases_sequenced"]
    # cycle c is covered by the reads longer than c
    cover = np.array([(b.len1 > c).sum() + (b.len2 > c).sum() for c in range(b.stride)])
    assert np.array_equal(d["pileup"].sum(axis=1)[: b.stride], cover)


# ---- the sibling tools on the same infrastructure: ReadQC and FastqTrim ---------------------------------------


def _qcml_values(path):
    vals = []
    with open(path, encoding="latin-1") as f:
        for line in f:
            m = re.search(r'<qualityParameter ID="qp\d+" name="([^"]+)".* value="([^"]*)"', line)
            if m:
                vals.append((m.group(1), m.group(2)))
    return vals


READQC_CASES = [
    ("base_test", [1], [2], "ReadQC_out1.qcML"),
    ("single_end", [1], [], "ReadQC_out3.qcML"),
    ("different_read_lengths", [3], [4], "ReadQC_out4.qcML"),  # holds the exact tie 0.125 MB -> "0.13"
    ("multiple_input_files", [1, 3], [2, 4], "ReadQC_out5.qcML"),
]


@pytest.mark.parametrize("case", READQC_CASES, ids=[c[0] for c in READQC_CASES])
def test_readqc_oracle_reproduces_reference_goldens(case):
    """src/tools-TEST/ReadQC_Test.cpp:8-44: the eight quality parameters of the reference's golden qcML files from the oracle's
    restatement of FastqEntry::validate + StatisticsReads::update / getResult."""
    _, in1, in2, golden = case
    d = H.oracle_readqc([f"{G}/ReadQC_in{k}.fastq.gz" for k in in1], [f"{G}/ReadQC_in{k}.fastq.gz" for k in in2])
    assert d["errors"] == 0
    want = _qcml_values(f"{G}/{golden}")
    assert len(want) == 8
    assert H.qc_metrics(d) == want


def test_readqc_oracle_txt_output():
    d = H.oracle_readqc([f"{G}/ReadQC_in1.fastq.gz"], [f"{G}/ReadQC_in2.fastq.gz"])
    text = "".join(f"{k}: {v}\n" for k, v in H.qc_metrics(d))
    assert text == open(f"{G}/ReadQC_out2.txt").read()  # ReadQC_Test.cpp:17-21


def test_validate_entry_codes():
    lib = H.oracle_lib()

    def v(h, b, h2, q):
        return lib.spo_validate_entry(h, len(h), b, len(b), h2, len(h2), q, len(q))

    assert v(b"@r", b"ACGTN", b"+", b"!IIIJ") == 0
    assert v(b"r", b"ACGT", b"+", b"IIII") == 1 and v(b"", b"", b"+", b"") == 1
    assert v(b"@r", b"ACGT", b"-", b"IIII") == 2 and v(b"@r", b"ACGT", b"", b"IIII") == 2
    assert v(b"@r", b"ACGT", b"+", b"III") == 3
    assert v(b"@r", b"ACgT", b"+", b"IIII") == 4 and v(b"@r", b"AC-T", b"+", b"IIII") == 4
    assert v(b"@r", b"ACGT", b"+", b"IIKI") == 5 and v(b"@r", b"ACGT", b"+", b"II I") == 5


FASTQTRIM_CASES = [  # src/tools-TEST/FastqTrim_Test.cpp:7-36
    ("start", 1, dict(start=5)),
    ("start_end", 2, dict(start=5, end=5)),
    ("start_len", 3, dict(start=5, max_bases=50)),
    ("max_len", 4, dict(end=5, max_len=80)),
    ("all", 5, dict(max_bases=50, start=5, end=5, max_len=80)),
]


@pytest.mark.parametrize("case", FASTQTRIM_CASES, ids=[c[0] for c in FASTQTRIM_CASES])
def test_fastqtrim_oracle_reproduces_reference_goldens(case):
    _, k, params = case
    recs = H.read_fastq4(f"{G}/FastqTrim_in1.fastq.gz")
    with gzip.open(f"{G}/FastqTrim_out{k}.fastq.gz", "rb") as f:
        want = f.read()
    assert H.oracle_fastq_trim(recs, **params) == want
