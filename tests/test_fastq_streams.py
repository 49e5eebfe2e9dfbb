"""CPU tests of the host-side FASTQ streams (ngs-bits_b200/host/FastqFileStream.*) against the reference's reader/writer fixtures
and expectations (src/cppNGS-TEST/FastqFileStream_Test.cpp:130-452): gz and plain input, trailing empty line, empty file, CRLF,
a truncated gz file (FileParseException), and the gz writer round trip."""
import gzip
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FQ = os.path.join(ROOT, "tests", "golden", "fastq")
DUMP = os.path.join(ROOT, "ngs-bits_b200", "bin", "fastq_dump")

HEADERS = ["@NG-5232_4_1_1022_17823#0/1", "@NG-5232_4_1_1025_18503#0/1", "@NG-5232_4_1_1026_21154#0/1", "@NG-5232_4_1_1028_9044#0/1", "@NG-5232_4_1_1031_3041#0/1",
           "@NG-5232_4_1_1031_18565#0/1", "@NG-5232_4_1_1031_20044#0/1", "@NG-5232_4_1_1032_18092#0/1", "@NG-5232_4_1_1033_5386#0/1", "@NG-5232_4_1_1033_2620#0/1"]
FIRST = ("@NG-5232_4_1_1022_17823#0/1",
         "NACTCCGGTGTCGGTCTCGTAGGCCATTTTAGAAGCGAATAAATCGATGNATTCGANCNCNNNNNNNNATCGNNAGAGCTCGTANGCCGTCTTCTGCTTGANNNNNNN",
         "+NG-5232_4_1_1022_17823#0/1",
         "#'''')(++)AAAAAAAAAA########################################################################################")


@pytest.fixture(scope="module")
def dump():
    # the helper only needs g++ and zlib (no CUDA library)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "ngs-bits_b200", "host"), "../bin/fastq_dump"], check=True)
    return DUMP


def run(dump, *args):
    r = subprocess.run([dump, *args], capture_output=True, text=True)
    return r.returncode, [l.split("\t") for l in r.stdout.rstrip("\n").split("\n")]


@pytest.mark.parametrize("name", ["example1.fastq.gz", "example2.fastq", "example3.fastq"])
def test_read_ten_entries(dump, name):
    """read_gzipped / read_plain / read_plain_emptylineatend: ten entries, atEnd false before each, then atEnd and an empty entry."""
    rc, rows = run(dump, os.path.join(FQ, name))
    assert rc == 0
    entries = rows[:-1]
    assert [e[1] for e in entries[:10]] == HEADERS
    assert all(e[0] == "0" for e in entries[:10])
    assert tuple(entries[0][1:5]) == FIRST
    assert entries[10] == ["1", "", "", "", ""]
    assert rows[-1] == ["END", "1"]


def test_read_plain_empty(dump):
    rc, rows = run(dump, os.path.join(FQ, "example4.fastq"))
    assert rc == 0
    assert rows[0][1:] == ["", "", "", ""]
    assert rows[-1] == ["END", "1"]


def test_read_plain_crlf(dump):
    rc, rows = run(dump, os.path.join(FQ, "example5.fastq"))
    assert rc == 0
    assert [e[1] for e in rows[:3]] == HEADERS[:3]
    assert tuple(rows[0][1:5]) == FIRST  # \r\n stripped
    assert rows[3] == ["1", "", "", "", ""]


def test_read_gzipped_corrupt(dump):
    """read_gzipped_corrupt: the truncated file delivers at least 316 good entries, then FileParseException."""
    rc, rows = run(dump, os.path.join(FQ, "example8.fastq.gz"))
    assert rc == 3 and rows[-1][0] == "FileParseException"
    good = rows[:-1]
    assert len(good) >= 316 and all(e[2] != "" for e in good[:316])


def test_write_gzipped_round_trip(dump, tmp_path):
    """write_gzipped: copy through FastqOutfileStream, read back: identical entries."""
    out = tmp_path / "copy.fastq.gz"
    rc, rows = run(dump, os.path.join(FQ, "example1.fastq.gz"), str(out))
    assert rc == 0
    rc2, rows2 = run(dump, str(out))
    assert rc2 == 0 and rows2 == rows
    with gzip.open(out, "rb") as f, gzip.open(os.path.join(FQ, "example1.fastq.gz"), "rb") as g:
        assert f.read() == g.read() + b"\n"  # the writer terminates the last record, the fixture does not
