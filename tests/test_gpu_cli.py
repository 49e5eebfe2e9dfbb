"""GPU tests of the drop-in command line (ngs-bits_b200/bin/seqpurge_b200 = reference flags + reference I/O surface + CUDA engine):
the reference's ten single-thread tool tests (src/tools-TEST/SeqPurge_Test.cpp:100-208) re-run against the reference's own golden
outputs, compared like the reference's COMPARE_FILES does (decompressed content), plus gz byte identity with the oracle CLI, which
issues the same zlib call sequence as the reference (src/cppNGS/FastqFileStream.cpp:160-193)."""
import gzip
import os
import subprocess

import pytest

import helpers as H
from test_oracle_golden import CASES, COMMON

pytestmark = pytest.mark.gpu
G = H.GOLDEN
ROOT = H.ROOT
CLI = os.path.join(ROOT, "ngs-bits_b200", "bin", "seqpurge_b200")


@pytest.fixture(scope="module")
def cli():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import __graft_entry__ as g

    g.build()
    assert os.path.exists(CLI)
    return CLI


def _content(path):
    with gzip.open(path, "rb") as f:
        return f.read()


def _raw(path):
    with open(path, "rb") as f:
        return f.read()


MODES = {"device_framing": [], "device_framing_4_deflate_threads": ["-threads", "4"], "host_framing": ["-host_framing"]}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cli_reproduces_reference_goldens(cli, case, mode, oracle_build, tmp_path):
    """device_framing: the default pipeline (FASTQ text framed, trimmed and laid out on the GPU); host_framing: the reference's block
    pipeline with records parsed on the host; -threads 4: parallel deflate of the output (same content, other .gz bytes)."""
    name, i1, i2, o1, o2, flags = case
    outs = {}
    for tool, exe in (("gpu", cli), ("oracle", os.path.join(oracle_build, "seqpurge_oracle"))):
        d = tmp_path / tool
        d.mkdir()
        fl = [str(d / "out15") if f == "OUT3" else f for f in flags]
        cmd = [exe, "-in1", f"{G}/SeqPurge_in{i1}.fastq.gz", "-in2", f"{G}/SeqPurge_in{i2}.fastq.gz", "-out1", str(d / "o1.fastq.gz"), "-out2", str(d / "o2.fastq.gz"),
               "-summary", str(d / "summary.txt")] + COMMON + fl + (MODES[mode] if tool == "gpu" else [])
        subprocess.run(cmd, check=True)
        outs[tool] = d
    g = outs["gpu"]
    # the reference's golden files, decompressed content (what COMPARE_FILES checks)
    assert _content(g / "o1.fastq.gz") == _content(f"{G}/SeqPurge_out{o1}.fastq.gz")
    assert _content(g / "o2.fastq.gz") == _content(f"{G}/SeqPurge_out{o2}.fastq.gz")
    if name == "test_07":
        assert _content(g / "out15_R1.fastq.gz") == _content(f"{G}/SeqPurge_out15_R1.fastq.gz")
        assert _content(g / "out15_R2.fastq.gz") == _content(f"{G}/SeqPurge_out15_R2.fastq.gz")
    # gz bytes: identical to the oracle CLI (one deflate stream per file, the reference's zlib call sequence)
    if "threads" not in mode:
        assert _raw(g / "o1.fastq.gz") == _raw(outs["oracle"] / "o1.fastq.gz")
        assert _raw(g / "o2.fastq.gz") == _raw(outs["oracle"] / "o2.fastq.gz")
    else:
        assert subprocess.run(["gzip", "-t", str(g / "o1.fastq.gz"), str(g / "o2.fastq.gz")]).returncode == 0
    # statistics summary (everything except the runtime line)
    def summ(p):
        return [l for l in open(p).read().split("\n") if not l.startswith("overall runtime")]
    assert summ(g / "summary.txt") == summ(outs["oracle"] / "summary.txt")


def test_cli_multiple_input_files_and_small_prefetch(cli, tmp_path):
    """A block may span an input-file boundary (InputWorker.cpp:26-38); few slots force slot reuse."""
    d = tmp_path
    cmd = [cli, "-in1", f"{G}/SeqPurge_in1.fastq.gz", f"{G}/SeqPurge_in7.fastq.gz", "-in2", f"{G}/SeqPurge_in2.fastq.gz", f"{G}/SeqPurge_in8.fastq.gz",
           "-out1", str(d / "o1.fastq.gz"), "-out2", str(d / "o2.fastq.gz"), "-summary", str(d / "s.txt"), "-block_size", "333", "-block_prefetch", "2",
           "-ncut", "0", "-qcut", "0", "-min_len", "15"]
    subprocess.run(cmd, check=True)
    want1 = _content(f"{G}/SeqPurge_out1.fastq.gz")
    got1 = _content(d / "o1.fastq.gz")
    assert got1.startswith(want1) and len(got1) > len(want1)


def test_cli_header_mismatch_is_an_error(cli, tmp_path):
    r = subprocess.run([cli, "-in1", f"{G}/SeqPurge_in1.fastq.gz", "-in2", f"{G}/SeqPurge_in4.fastq.gz", "-out1", str(tmp_path / "a.gz"), "-out2", str(tmp_path / "b.gz"),
                        "-block_size", "100", "-block_prefetch", "1"], capture_output=True, text=True)  # one job: the reader cannot run ahead into the length mismatch
    assert r.returncode == 1 and "Headers of reads do not match" in r.stderr


def test_cli_unequal_entry_counts_is_an_error(cli, tmp_path):
    """InputWorker.cpp:39-46: one stream ends before the other."""
    r = subprocess.run([cli, "-in1", f"{G}/SeqPurge_in1.fastq.gz", "-in2", f"{G}/SeqPurge_in4.fastq.gz", "-out1", str(tmp_path / "a.gz"), "-out2", str(tmp_path / "b.gz")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "has more entries than" in r.stderr


def test_cli_config1_synthetic_10k_gz_byte_identical(cli, oracle_build, tmp_path):
    """BASELINE config 1: 10 k synthetic 2x150 pairs (device generator), default adapters, FASTQ.gz in -> trimmed FASTQ.gz out;
    output files byte-identical to the CPU oracle run with -threads 1, also with two blocks in flight per slot ring."""
    import gzip as gz

    import numpy as np
    import torch

    import seqpurge_b200 as sp

    n, L = 10_000, 150
    dev = torch.device("cuda:0")
    t = {k: torch.empty((n, L), dtype=torch.uint8, device=dev) for k in ("bases1", "quals1", "bases2", "quals2")}
    l1 = torch.empty(n, dtype=torch.int16, device=dev)
    l2 = torch.empty(n, dtype=torch.int16, device=dev)
    sp.synth_device(sp.SynthConfig(read_len=L, error_rate=0.001, lowq_tail_mean=3.0), 0, n, t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2)
    torch.cuda.synchronize()
    host = {k: v.cpu().numpy() for k, v in t.items()}
    for r, (bk, qk) in enumerate((("bases1", "quals1"), ("bases2", "quals2")), start=1):
        with gz.open(tmp_path / f"in{r}.fastq.gz", "wb", compresslevel=1) as f:
            for i in range(n):
                f.write(b"@SIM:1:B200:1:%d:%d %d:N:0:ACGT\n" % (i // 1000, i, r))
                f.write(host[bk][i].tobytes() + b"\n+\n" + host[qk][i].tobytes() + b"\n")
    outs = {}
    for tool, exe, extra in (("gpu", cli, ["-block_size", "1500", "-block_prefetch", "3"]), ("oracle", os.path.join(oracle_build, "seqpurge_oracle"), ["-threads", "1"])):
        d = tmp_path / tool
        d.mkdir()
        subprocess.run([exe, "-in1", str(tmp_path / "in1.fastq.gz"), "-in2", str(tmp_path / "in2.fastq.gz"), "-out1", str(d / "o1.fastq.gz"), "-out2", str(d / "o2.fastq.gz"),
                        "-out3", str(d / "single"), "-summary", str(d / "s.txt")] + extra, check=True)
        outs[tool] = d
    for name in ("o1.fastq.gz", "o2.fastq.gz", "single_R1.fastq.gz", "single_R2.fastq.gz"):
        assert _raw(outs["gpu"] / name) == _raw(outs["oracle"] / name), name
    assert len(_content(outs["gpu"] / "o1.fastq.gz")) > 1_000_000


def test_cli_bgzf_out_and_parallel_bgzf_in(cli, tmp_path):
    """-bgzf writes blocked gzip (decompressed content = the reference's golden files); BGZF inputs are inflated by the -threads pool
    and give the same output as the original gzip inputs."""
    d = tmp_path
    base = [cli, "-summary", str(d / "s.txt"), "-ncut", "0", "-qcut", "0", "-min_len", "15"]
    subprocess.run(base + ["-in1", f"{G}/SeqPurge_in1.fastq.gz", "-in2", f"{G}/SeqPurge_in2.fastq.gz", "-out1", str(d / "o1.gz"), "-out2", str(d / "o2.gz"),
                           "-bgzf", "-threads", "4"], check=True)
    assert _content(d / "o1.gz") == _content(f"{G}/SeqPurge_out1.fastq.gz")
    assert _content(d / "o2.gz") == _content(f"{G}/SeqPurge_out2.fastq.gz")
    assert _raw(d / "o1.gz")[:4] == b"\x1f\x8b\x08\x04" and _raw(d / "o1.gz")[12:16] == b"BC\x02\x00"
    # the inputs re-blocked as BGZF (bin/gzpipe), then trimmed with parallel inflate
    pipe = os.path.join(ROOT, "ngs-bits_b200", "bin", "gzpipe")
    for r in (1, 2):
        subprocess.run([pipe, f"{G}/SeqPurge_in{r}.fastq.gz", str(d / f"in{r}.bgzf.gz"), "-bgzf", "-threads", "2"], check=True)
    subprocess.run(base + ["-in1", str(d / "in1.bgzf.gz"), "-in2", str(d / "in2.bgzf.gz"), "-out1", str(d / "p1.gz"), "-out2", str(d / "p2.gz"), "-threads", "4"], check=True)
    assert _content(d / "p1.gz") == _content(f"{G}/SeqPurge_out1.fastq.gz")
    assert _content(d / "p2.gz") == _content(f"{G}/SeqPurge_out2.fastq.gz")
