"""CPU tests of the command line's text streams (ngs-bits_b200/host/TextSource.cpp, GzipTextWriter.cpp) through the host-only helper
bin/gzpipe: serial gzip (the reference's gzFile path, src/cppCORE/VersatileFile.cpp:274-308), parallel deflate, BGZF output, parallel
BGZF inflate, the fall-back to gzFile for an ordinary member inside a BGZF file, and the error paths."""
import gzip
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import helpers as H

PIPE = os.path.join(H.ROOT, "ngs-bits_b200", "bin", "gzpipe")


@pytest.fixture(scope="module")
def pipe():
    subprocess.run(["make", "-s", "-C", os.path.join(H.ROOT, "ngs-bits_b200", "host"), "../bin/gzpipe"], check=True)
    return PIPE


def fastq_text(n, seed=3):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    out = []
    for i in range(n):
        ln = int(rng.integers(30, 151))
        out.append(b"@R:%d 1:N:0\n" % i + acgt[rng.integers(0, 4, ln)].tobytes() + b"\n+\n" + b"I" * ln + b"\n")
    return b"".join(out)


def run(pipe, *args):
    return subprocess.run([pipe, *[str(a) for a in args]], capture_output=True, text=True)


def content(path):
    with gzip.open(path, "rb") as f:
        return f.read()


def bgzf_blocks(path):
    """(compressed size, uncompressed size) of every BGZF block; raises if the file is not a sequence of BGZF blocks."""
    d = open(path, "rb").read()
    off, blocks = 0, []
    while off < len(d):
        assert d[off : off + 4] == b"\x1f\x8b\x08\x04", "not a BGZF block"
        xlen = struct.unpack_from("<H", d, off + 10)[0]
        extra = d[off + 12 : off + 12 + xlen]
        assert extra[:4] == b"BC\x02\x00"
        bsize = struct.unpack_from("<H", extra, 4)[0] + 1
        isize = struct.unpack_from("<I", d, off + bsize - 4)[0]
        raw = zlib.decompress(d[off + 12 + xlen : off + bsize - 8], -15)
        assert len(raw) == isize and zlib.crc32(raw) == struct.unpack_from("<I", d, off + bsize - 8)[0]
        blocks.append((bsize, isize))
        off += bsize
    return blocks


def test_serial_gzip_round_trip(pipe, tmp_path):
    text = fastq_text(20000)
    src = tmp_path / "in.gz"
    with gzip.open(src, "wb", compresslevel=1) as f:
        f.write(text)
    r = run(pipe, src, tmp_path / "o.gz")
    assert r.returncode == 0 and "serial inflate" in r.stderr
    assert content(tmp_path / "o.gz") == text
    # plain text and multi-member gzip inputs (what gzFile accepts)
    (tmp_path / "plain.fastq").write_bytes(text)
    assert run(pipe, tmp_path / "plain.fastq", tmp_path / "p.gz").returncode == 0
    assert content(tmp_path / "p.gz") == text
    (tmp_path / "multi.gz").write_bytes(src.read_bytes() * 3)
    assert run(pipe, tmp_path / "multi.gz", tmp_path / "m.gz", "-threads", 4).returncode == 0
    assert content(tmp_path / "m.gz") == text * 3


@pytest.mark.parametrize("threads", [1, 4])
def test_bgzf_output_is_valid_bgzf(pipe, tmp_path, threads):
    text = fastq_text(30000)
    (tmp_path / "plain.fastq").write_bytes(text)
    out = tmp_path / "o.bgzf.gz"
    assert run(pipe, tmp_path / "plain.fastq", out, "-bgzf", "-threads", threads).returncode == 0
    assert content(out) == text  # any gzip reader
    assert subprocess.run(["gzip", "-t", str(out)]).returncode == 0
    blocks = bgzf_blocks(out)
    assert blocks[-1] == (28, 0), "end-of-file marker block"
    assert max(b[0] for b in blocks) <= 65536 and max(b[1] for b in blocks) <= 0xFF00
    assert sum(b[1] for b in blocks) == len(text)


def test_parallel_bgzf_inflate(pipe, tmp_path):
    text = fastq_text(60000)
    (tmp_path / "plain.fastq").write_bytes(text)
    bg = tmp_path / "in.bgzf.gz"
    assert run(pipe, tmp_path / "plain.fastq", bg, "-bgzf", "-threads", 4).returncode == 0
    r = run(pipe, bg, tmp_path / "o.gz", "-threads", 4)
    assert r.returncode == 0 and "parallel BGZF inflate" in r.stderr
    assert content(tmp_path / "o.gz") == text
    r = run(pipe, bg, tmp_path / "o1.gz", "-threads", 1)  # no pool: gzFile reads BGZF like any multi-member gzip
    assert r.returncode == 0 and "serial inflate" in r.stderr
    assert content(tmp_path / "o1.gz") == text


def test_bgzf_followed_by_ordinary_member_falls_back(pipe, tmp_path):
    text = fastq_text(20000)
    (tmp_path / "plain.fastq").write_bytes(text)
    bg = tmp_path / "a.bgzf.gz"
    assert run(pipe, tmp_path / "plain.fastq", bg, "-bgzf", "-threads", 4).returncode == 0
    with gzip.open(tmp_path / "b.gz", "wb", compresslevel=1) as f:
        f.write(text[:100000])
    (tmp_path / "mix.gz").write_bytes(bg.read_bytes() + (tmp_path / "b.gz").read_bytes())
    assert run(pipe, tmp_path / "mix.gz", tmp_path / "o.gz", "-threads", 4).returncode == 0
    assert content(tmp_path / "o.gz") == text + text[:100000]


def test_bgzf_errors(pipe, tmp_path):
    text = fastq_text(20000)
    (tmp_path / "plain.fastq").write_bytes(text)
    bg = tmp_path / "a.bgzf.gz"
    assert run(pipe, tmp_path / "plain.fastq", bg, "-bgzf").returncode == 0
    d = bg.read_bytes()
    (tmp_path / "trunc.gz").write_bytes(d[: len(d) // 2])
    r = run(pipe, tmp_path / "trunc.gz", tmp_path / "o.gz", "-threads", 4)
    assert r.returncode == 1 and "truncated BGZF block" in r.stderr
    bad = bytearray(d)
    bad[len(d) // 3] ^= 0x55
    (tmp_path / "bad.gz").write_bytes(bytes(bad))
    r = run(pipe, tmp_path / "bad.gz", tmp_path / "o.gz", "-threads", 4)
    assert r.returncode == 1 and "corrupt BGZF block" in r.stderr
    r = run(pipe, tmp_path / "missing.gz", tmp_path / "o.gz")
    assert r.returncode == 1 and "Could not open file" in r.stderr
