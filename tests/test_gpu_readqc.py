"""GPU tests of readqc_b200 (ngs-bits_b200/host/readqc_main.cpp): the reference's ReadQC tool tests (src/tools-TEST/ReadQC_Test.cpp:8-52)
re-run against the reference's golden files, compared like the reference compares them ('creation ' and <binary> lines dropped; here
also the embedded stylesheet, which this tool does not write), plus the statistics-only / single-end form of the FASTQ stream
(spg_fq_config.stats_only) against the CPU oracle."""
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
G = H.GOLDEN
TOOL = os.path.join(H.ROOT, "ngs-bits_b200", "bin", "readqc_b200")
PIPE = os.path.join(H.ROOT, "ngs-bits_b200", "bin", "gzpipe")


@pytest.fixture(scope="module")
def sp():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import __graft_entry__ as g

    g.build()
    import seqpurge_b200

    assert os.path.exists(TOOL)
    return seqpurge_b200


def qcml_lines(path):
    keep = []
    with open(path, encoding="latin-1") as f:
        for line in f:
            if re.search(r"creation |<binary>", line):  # ReadQC_Test.cpp:13-14
                continue
            keep.append(line.rstrip("\n"))
    text = "\n".join(keep)
    text = re.sub(r"<\?xml-stylesheet.*?\]>\n", "", text, flags=re.S)
    text = re.sub(r"  <xsl:stylesheet.*</xsl:stylesheet>\n", "", text, flags=re.S)
    return text.split("\n")


def run(*args):
    return subprocess.run([TOOL, *[str(a) for a in args]], capture_output=True, text=True)


def test_base_test(sp, tmp_path):
    r = run("-in1", f"{G}/ReadQC_in1.fastq.gz", "-in2", f"{G}/ReadQC_in2.fastq.gz", "-out", tmp_path / "o.qcML")
    assert r.returncode == 0, r.stderr
    assert qcml_lines(tmp_path / "o.qcML") == qcml_lines(f"{G}/ReadQC_out1.qcML")


def test_with_txt_parameter(sp, tmp_path):
    r = run("-in1", f"{G}/ReadQC_in1.fastq.gz", "-in2", f"{G}/ReadQC_in2.fastq.gz", "-out", tmp_path / "o.txt", "-txt")
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "o.txt").read() == open(f"{G}/ReadQC_out2.txt").read()
    r = run("-in1", f"{G}/ReadQC_in1.fastq.gz", "-in2", f"{G}/ReadQC_in2.fastq.gz", "-txt")  # STDOUT if -out is unset
    assert r.returncode == 0 and r.stdout == open(f"{G}/ReadQC_out2.txt").read()


def test_single_end(sp, tmp_path):
    r = run("-in1", f"{G}/ReadQC_in1.fastq.gz", "-out", tmp_path / "o.qcML")
    assert r.returncode == 0, r.stderr
    assert qcml_lines(tmp_path / "o.qcML") == qcml_lines(f"{G}/ReadQC_out3.qcML")


def test_different_read_lengths(sp, tmp_path):
    r = run("-in1", f"{G}/ReadQC_in3.fastq.gz", "-in2", f"{G}/ReadQC_in4.fastq.gz", "-out", tmp_path / "o.qcML", "-block_size", 100)
    assert r.returncode == 0, r.stderr
    assert qcml_lines(tmp_path / "o.qcML") == qcml_lines(f"{G}/ReadQC_out4.qcML")


@pytest.mark.parametrize("threads", [1, 4])
def test_multiple_input_files(sp, tmp_path, threads):
    r = run("-in1", f"{G}/ReadQC_in1.fastq.gz", f"{G}/ReadQC_in3.fastq.gz", "-in2", f"{G}/ReadQC_in2.fastq.gz", f"{G}/ReadQC_in4.fastq.gz", "-out", tmp_path / "o.qcML",
            "-threads", threads, "-block_size", 5000)
    assert r.returncode == 0, r.stderr
    assert qcml_lines(tmp_path / "o.qcML") == qcml_lines(f"{G}/ReadQC_out5.qcML")


def test_bgzf_inputs_parallel_inflate(sp, tmp_path):
    for k in (1, 2):
        subprocess.run([PIPE, f"{G}/ReadQC_in{k}.fastq.gz", str(tmp_path / f"in{k}.fastq.gz"), "-bgzf", "-threads", "2"], check=True)
    r = run("-in1", tmp_path / "in1.fastq.gz", "-in2", tmp_path / "in2.fastq.gz", "-txt", "-threads", 4)
    assert r.returncode == 0 and r.stdout == open(f"{G}/ReadQC_out2.txt").read()


def _write(path, records):
    with gzip.open(path, "wb") as f:
        for h, b, h2, q in records:
            f.write(h + b"\n" + b + b"\n" + h2 + b"\n" + q + b"\n")


def test_errors_of_the_reference(sp, tmp_path):
    r = run("-in1", f"{G}/ReadQC_in1.fastq.gz", "-in2", f"{G}/ReadQC_in4.fastq.gz", "-txt")
    assert r.returncode == 1 and "Differing number of reads in file" in r.stderr  # main.cpp:93-97
    r = run("-in1", f"{G}/ReadQC_in1.fastq.gz", f"{G}/ReadQC_in3.fastq.gz", "-in2", f"{G}/ReadQC_in2.fastq.gz", "-txt")
    assert r.returncode == 1 and "differ in counts" in r.stderr  # main.cpp:43-46
    good = (b"@r1", b"ACGTN", b"+", b"IIII!")
    cases = [
        ((b"r2", b"ACGT", b"+", b"IIII"), "First header line does not start with '@': 'r2'."),
        ((b"@r2", b"ACGT", b"-", b"IIII"), "Second header line does not start with '+': '-'."),
        ((b"@r2", b"ACGT", b"+", b"III"), "Differing length of bases (4) and qualities string (4) in sequence '@r2'."),
        ((b"@r2", b"ACXT", b"+", b"IIII"), "Invalid base 'X' encountered in sequence '@r2'."),
        ((b"@r2", b"ACgT", b"+", b"IIII"), "Invalid base 'g' encountered in sequence '@r2'."),
        ((b"@r2", b"ACGT", b"+", b"IIKI"), "Invalid quality character 'K' with value '75' encountered in sequence '@r2'."),
    ]
    for i, (bad, msg) in enumerate(cases):  # FastqEntry::validate, src/cppNGS/FastqFileStream.cpp:3-48
        p = tmp_path / f"bad{i}.fastq.gz"
        _write(p, [good] * 7 + [bad] + [good] * 3)
        r = run("-in1", p, "-txt")
        assert r.returncode == 1 and ("Invalid Fastq file entry: " + msg) in r.stderr, (msg, r.stderr)
        ok = tmp_path / f"ok{i}.fastq.gz"
        _write(ok, [good] * 11)
        r = run("-in1", ok, "-in2", p, "-txt")  # the same in the reverse file
        assert r.returncode == 1 and msg in r.stderr, (msg, r.stderr)
    _write(tmp_path / "ok.fastq.gz", [good] * 11)
    r = run("-in1", tmp_path / "ok.fastq.gz", "-txt")
    assert r.returncode == 0 and "read count: 11" in r.stdout


def _fastq_text(batch, which, n):
    b, q, ln = (batch.bases1, batch.quals1, batch.len1) if which == 1 else (batch.bases2, batch.quals2, batch.len2)
    out = []
    for i in range(n):
        out.append(b"@R:%d %d\n" % (i, which) + b[i, : ln[i]].tobytes() + b"\n+\n" + q[i, : ln[i]].tobytes() + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("single", [False, True], ids=["paired", "single_end"])
def test_stats_only_stream_against_oracle(sp, single):
    """spg_fq_* with stats_only: every accumulator of the device reduction equals the oracle's StatisticsReads::update restatement."""
    batch = H.random_batch(3000, 150, seed=31, ragged=True, n_rate=0.01)
    batch.quals1[:] = np.minimum(batch.quals1, 74)
    batch.quals2[:] = np.minimum(batch.quals2, 74)
    n = batch.n
    eng = sp.Engine(sp.TrimmingParameters(qc=6), devices=(0,))  # 2: the checks of FastqEntry::validate, +4: the plot histograms
    fq = sp.FastqStream(eng, n_slots=2, max_pairs=1024, max_len=160, text_cap=4 << 20, stats_only=True, single_end=single, validate=True)
    t1, t2 = _fastq_text(batch, 1, n), (b"" if single else _fastq_text(batch, 2, n))
    off1 = off2 = 0
    slot = 0
    while off1 < len(t1):
        c1, c2 = t1[off1 : off1 + (1 << 20)], t2[off2 : off2 + (1 << 20)]
        fq.submit(slot, c1, c2, final1=off1 + len(c1) >= len(t1), final2=single or off2 + len(c2) >= len(t2))
        ch = fq.wait(slot)
        assert ch.error_pair == -1 and ch.n_pairs > 0
        off1 += ch.consumed[0]
        off2 += ch.consumed[1]
        slot ^= 1
    got = eng.qc_stats()
    want = H.oracle_qc(batch)
    if single:  # the oracle counts the (empty) mates as reverse reads: take them out
        empty = H.Batch(n, batch.stride)
        empty.len1[:n] = batch.len1[:n]
        empty.bases1[:], empty.quals1[:] = batch.bases1, batch.quals1
        want = H.oracle_qc(empty)
        want["reads_reverse"] = 0
        want["read_lengths"][0] -= n
    for k, v in want.items():
        assert np.array_equal(np.asarray(got[k]), np.asarray(v)), k
    fq.close()
    eng.close()
