"""GPU tests of fastqtrim_b200 (ngs-bits_b200/host/fastqtrim_main.cpp): the reference's FastqTrim tool tests
(src/tools-TEST/FastqTrim_Test.cpp:7-36) against the reference's golden files (decompressed content, as COMPARE_FILES compares .gz files),
and a restatement of the rule of src/FastqTrim/main.cpp:47-77 on random reads (empty reads, reads that vanish, every boundary)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
G = H.GOLDEN
TOOL = os.path.join(H.ROOT, "ngs-bits_b200", "bin", "fastqtrim_b200")


@pytest.fixture(scope="module")
def tool():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import __graft_entry__ as g

    g.build()
    assert os.path.exists(TOOL)
    return TOOL


def content(path):
    with gzip.open(path, "rb") as f:
        return f.read()


CASES = [
    ("start", 1, ["-start", "5"]),
    ("start_end", 2, ["-start", "5", "-end", "5"]),
    ("start_len", 3, ["-start", "5", "-len", "50"]),
    ("max_len", 4, ["-end", "5", "-max_len", "80"]),
    ("all", 5, ["-len", "50", "-start", "5", "-end", "5", "-max_len", "80"]),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_tool_tests(tool, case, tmp_path):
    _, k, flags = case
    out = tmp_path / "o.fastq.gz"
    r = subprocess.run([tool, "-in", f"{G}/FastqTrim_in1.fastq.gz", "-out", str(out)] + flags, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert content(out) == content(f"{G}/FastqTrim_out{k}.fastq.gz")


def fastq_trim_rule(records, start, end, ln, max_len):
    """src/FastqTrim/main.cpp:47-77"""
    out = []
    for h, b, h2, q in records:
        if max_len > 0 and len(b) >= max_len:
            out.append((h, b, h2, q))
            continue
        if start > 0 or end > 0:
            n = len(b)
            if n <= start + end:
                continue
            b, q = b[start : n - end], q[start : n - end]
        if ln > 0 and len(b) > ln:
            b, q = b[:ln], q[:ln]
        out.append((h, b, h2, q))
    return out


def text_of(records):
    return b"".join(h + b"\n" + b + b"\n" + h2 + b"\n" + q + b"\n" for h, b, h2, q in records)


@pytest.mark.parametrize("params", [(0, 0, 0, 0), (3, 0, 0, 0), (0, 7, 0, 0), (5, 5, 20, 0), (10, 10, 0, 60), (0, 0, 1, 0), (200, 0, 0, 0), (4, 4, 30, 31), (1, 1, 148, 151)],
                         ids=lambda p: "s%d_e%d_l%d_m%d" % p)
def test_rule_on_random_reads(tool, params, tmp_path):
    start, end, ln, max_len = params
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGTN", np.uint8)
    recs = []
    for i in range(5000):
        n = int(rng.choice([0, 1, 2, 9, 10, 11, 20, 30, 31, 59, 60, 61, 100, 150, 151])) if i % 3 else int(rng.integers(0, 152))
        recs.append((b"@R%d extra:%d" % (i, n), acgt[rng.integers(0, 5, n)].tobytes(), b"+" if i % 2 else b"+R%d" % i, bytes(rng.integers(33, 75, n, dtype=np.uint8))))
    src = tmp_path / "in.fastq.gz"
    with gzip.open(src, "wb", compresslevel=1) as f:
        f.write(text_of(recs))
    for extra in ([], ["-threads", "4", "-bgzf", "-block_size", "777"]):
        out = tmp_path / "o.fastq.gz"
        cmd = [tool, "-in", str(src), "-out", str(out), "-start", str(start), "-end", str(end), "-len", str(ln), "-max_len", str(max_len)] + extra
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        want = H.oracle_fastq_trim(recs, start, end, ln, max_len)  # oracle/: C restatement pinned to the reference's goldens (test_oracle_golden.py)
        assert want == text_of(fastq_trim_rule(recs, start, end, ln, max_len))
        assert content(out) == want


def test_gz_bytes_of_the_serial_writer(tool, tmp_path):
    """-threads 1: the output file is what gzopen/gzbuffer(131072)/gzsetparams/gzwrite produce for that text, i.e. the bytes of the
    reference's FastqOutfileStream (src/cppNGS/FastqFileStream.cpp:160-198) given the same zlib."""
    out = tmp_path / "o.fastq.gz"
    subprocess.run([tool, "-in", f"{G}/SeqPurge_in1.fastq.gz", "-out", str(out), "-start", "2", "-len", "100"], check=True)
    pipe = os.path.join(H.ROOT, "ngs-bits_b200", "bin", "gzpipe")
    (tmp_path / "plain.fastq").write_bytes(content(out))
    subprocess.run([pipe, str(tmp_path / "plain.fastq"), str(tmp_path / "ref.gz"), "-level", "1"], check=True)
    assert open(out, "rb").read() == open(tmp_path / "ref.gz", "rb").read()


def test_errors(tool, tmp_path):
    r = subprocess.run([tool, "-in", str(tmp_path / "missing.fastq.gz"), "-out", str(tmp_path / "o.gz")], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not open file" in r.stderr
    with gzip.open(tmp_path / "bad.fastq.gz", "wb") as f:
        f.write(b"@a\nACGT\n+\nIIII\n@b\nACGT\n+\nIII\n")
    r = subprocess.run([tool, "-in", str(tmp_path / "bad.fastq.gz"), "-out", str(tmp_path / "o.gz"), "-start", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and "Differing length of bases and qualities string in sequence '@b'" in r.stderr
