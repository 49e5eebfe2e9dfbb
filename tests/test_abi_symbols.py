"""CPU test: the C-ABI library builds for sm_100a, loads, and exports every function include/seqpurge_b200.h declares.
No compute calls here (no GPU in this container)."""
import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.build()
    return ctypes.CDLL(g.LIB)


def declared_functions():
    text = open(os.path.join(ROOT, "include", "seqpurge_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spg_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_seam():
    names = declared_functions()
    for required in ("spg_create", "spg_slot_buffers", "spg_submit", "spg_wait", "spg_trim_device", "spg_last_error", "spg_destroy"):
        assert required in names


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/seqpurge_b200.h but not exported"


def test_create_without_gpu_fails_loudly(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sys.path.insert(0, os.path.join(ROOT, "ngs-bits_b200"))
    import seqpurge_b200 as sp

    with pytest.raises(sp.SeqPurgeError, match="no CUDA device|CPU path"):
        sp.Engine(sp.TrimmingParameters(), devices=(0,), n_slots=1, max_pairs=8, max_len=150)


def test_struct_layouts_match_header():
    sys.path.insert(0, os.path.join(ROOT, "ngs-bits_b200"))
    import seqpurge_b200 as sp

    assert sp.RESULT_DTYPE.itemsize == 8
    assert ctypes.sizeof(sp._EcStats) == 3 * 1000 * 8
    assert ctypes.sizeof(sp._SlotView) == 6 * 8 + 8


def test_ctypes_structs_have_the_sizes_the_compiler_gives_the_header(tmp_path):
    """Every struct of include/seqpurge_b200.h that the Python binding mirrors: sizeof from gcc == ctypes.sizeof (guards the binding
    against fields added to the header)."""
    import subprocess

    sys.path.insert(0, os.path.join(ROOT, "ngs-bits_b200"))
    import seqpurge_b200 as sp

    pairs = {"spg_params": sp._Params, "spg_slot_view": sp._SlotView, "spg_ec_stats": sp._EcStats, "spg_qc_stats": sp._QcStats, "spg_fq_config": sp._FqConfig,
             "spg_fq_input": sp._FqInput, "spg_fq_output": sp._FqOutput, "spg_fq_stats": sp._FqStats, "spg_synth_config": sp._SynthConfig}
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "seqpurge_b200.h"\nint main(void){' + "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in pairs) + "return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    sizes = dict(line.split() for line in out.strip().split("\n"))
    for name, ct in pairs.items():
        assert int(sizes[name]) == ctypes.sizeof(ct), name


def test_command_lines_fail_loudly_without_a_gpu(lib, tmp_path):
    """seqpurge_b200 / readqc_b200 / fastqtrim_b200: usage errors are reported as such, and on a box without a CUDA device the tools stop
    with the engine's error instead of computing anything on the CPU."""
    import gzip
    import subprocess

    import torch

    bin_dir = os.path.join(ROOT, "ngs-bits_b200", "bin")
    fq = tmp_path / "r.fastq.gz"
    with gzip.open(fq, "wb") as f:
        f.write(b"@r1\nACGTACGTAC\n+\nIIIIIIIIII\n")
    for tool, usage_args, run_args in (
        ("seqpurge_b200", [], ["-in1", str(fq), "-in2", str(fq), "-out1", str(tmp_path / "a.gz"), "-out2", str(tmp_path / "b.gz")]),
        ("readqc_b200", [], ["-in1", str(fq), "-txt"]),
        ("fastqtrim_b200", ["-in", str(fq)], ["-in", str(fq), "-out", str(tmp_path / "c.gz"), "-start", "1"]),
    ):
        exe = os.path.join(bin_dir, tool)
        assert os.path.exists(exe), exe
        r = subprocess.run([exe] + usage_args, capture_output=True, text=True)
        assert r.returncode == 1 and "Mandatory parameter" in r.stderr, (tool, r.stderr)
        r = subprocess.run([exe, "-nonsense"], capture_output=True, text=True)
        assert r.returncode == 1 and "Unknown parameter" in r.stderr, (tool, r.stderr)
        assert subprocess.run([exe, "--help"], capture_output=True, text=True).returncode == 0
        if not torch.cuda.is_available():
            r = subprocess.run([exe] + run_args, capture_output=True, text=True)
            assert r.returncode == 1 and "CUDA" in r.stderr, (tool, r.stderr)
    # numbers are parsed strictly like ToolBase does (no silent 0 for '-qcut abc'), device lists must name each device once, adapters are trimmed
    exe = os.path.join(bin_dir, "seqpurge_b200")
    base = ["-in1", str(fq), "-in2", str(fq), "-out1", str(tmp_path / "a.gz"), "-out2", str(tmp_path / "b.gz")]
    for extra, msg in ((["-qcut", "abc"], "is not an integer"), (["-mep", "1e-6x"], "is not a number"), (["-gpus", "0,0"], "listed twice"), (["-gpus", "-1"], "not a CUDA device index"),
                       (["-progress", "100"], "not supported"), (["-a1", "  ACGT  "], "too short")):
        r = subprocess.run([exe] + base + extra, capture_output=True, text=True)
        assert r.returncode == 1 and msg in r.stderr, (extra, r.stderr)


def test_slot_quality_tails_helper(lib):
    """Slot.fill_qtails (what a stager does for SPG_OPT_QUAL_TAILS): row r of a tail plane = the last 16 qualities of read r, a shorter
    read right-aligned. Checked against a direct loop on a slot made of plain numpy arrays (no device needed)."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "ngs-bits_b200"))
    import seqpurge_b200 as sp  # the ctypes binding (loads the library `lib` built; needs no device for this)

    s = sp.Slot.__new__(sp.Slot)
    n, stride = 64, 40
    rng = np.random.default_rng(1)
    s.quals1 = rng.integers(33, 74, (n, stride)).astype(np.uint8)
    s.quals2 = rng.integers(33, 74, (n, stride)).astype(np.uint8)
    s.len1 = rng.integers(0, stride + 1, n).astype(np.uint16)
    s.len2 = rng.integers(0, stride + 1, n).astype(np.uint16)
    s.len1[:3] = (0, 16, stride)
    s.qtail1 = np.full((n, sp.QTAIL), 255, np.uint8)
    s.qtail2 = np.full((n, sp.QTAIL), 255, np.uint8)
    s.fill_qtails(n)
    for q, ln, qt in ((s.quals1, s.len1, s.qtail1), (s.quals2, s.len2, s.qtail2)):
        for i in range(n):
            L = int(ln[i])
            k = min(L, sp.QTAIL)
            assert (qt[i, sp.QTAIL - k :] == q[i, L - k : L]).all(), i
