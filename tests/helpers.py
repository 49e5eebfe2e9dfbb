"""Test helpers: ctypes binding of the CPU oracle (oracle/), FASTQ fixtures -> SoA batches, synthetic batches in numpy."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_DIR = os.path.join(ROOT, "oracle")

RESULT_DTYPE = np.dtype([("len1", "<u2"), ("len2", "<u2"), ("best_offset", "<i2"), ("flags", "u1"), ("status", "u1")])
MAXLEN = 1000
DEFAULT_A1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA"
DEFAULT_A2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT"


class _SpoParams(C.Structure):
    _fields_ = [
        ("a1", C.c_char_p), ("a1_len", C.c_int), ("a2", C.c_char_p), ("a2_len", C.c_int), ("a_size", C.c_int),
        ("adapter_overlap", C.c_int), ("match_perc", C.c_double), ("mep", C.c_double),
        ("qcut", C.c_int), ("qwin", C.c_int), ("qoff", C.c_int), ("ncut", C.c_int), ("ec", C.c_int),
    ]


class _SpoEc(C.Structure):
    _fields_ = [("mismatch_r1", C.c_int64 * MAXLEN), ("mismatch_r2", C.c_int64 * MAXLEN), ("errors_per_read", C.c_int64 * MAXLEN)]


class _SpoQc(C.Structure):
    _fields_ = [
        ("reads_forward", C.c_int64), ("reads_reverse", C.c_int64), ("bases_sequenced", C.c_int64), ("read_q20", C.c_int64),
        ("base_q20", C.c_int64), ("base_q30", C.c_int64), ("errors", C.c_int64),
        ("read_lengths", C.c_int64 * MAXLEN), ("pileup", (C.c_int64 * 5) * MAXLEN),
        ("qsum_forward", C.c_int64 * MAXLEN), ("qsum_reverse", C.c_int64 * MAXLEN),
        ("base_qualities", C.c_int64 * 100), ("read_qualities", C.c_int64 * 100), ("qscore_dist_forward", C.c_int64 * 60), ("qscore_dist_reverse", C.c_int64 * 60),
    ]


def _qc_to_dict(st):
    d = {k: int(getattr(st, k)) for k in ("reads_forward", "reads_reverse", "bases_sequenced", "read_q20", "base_q20", "base_q30", "errors")}
    d["read_lengths"] = np.array(st.read_lengths, dtype=np.int64)
    d["pileup"] = np.array([list(row) for row in st.pileup], dtype=np.int64)
    d["qsum_forward"] = np.array(st.qsum_forward, dtype=np.int64)
    d["qsum_reverse"] = np.array(st.qsum_reverse, dtype=np.int64)
    for k in ("base_qualities", "read_qualities", "qscore_dist_forward", "qscore_dist_reverse"):
        d[k] = np.array(getattr(st, k), dtype=np.int64)
    return d


def oracle_qc(batch):
    """StatisticsReads::update over a Batch (raw reads) by the CPU oracle -> dict."""
    lib = oracle_lib()
    st = _SpoQc()
    lib.spo_qc_update_batch(batch.bases1.ctypes.data, batch.quals1.ctypes.data, batch.bases2.ctypes.data, batch.quals2.ctypes.data,
                            batch.len1.ctypes.data, batch.len2.ctypes.data, batch.stride, batch.n, C.byref(st))
    return _qc_to_dict(st)


def read_fastq4(path):
    """All four lines of every record (header, bases, header2, qualities)."""
    with gzip.open(path, "rb") as f:
        lines = f.read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return [tuple(l.rstrip(b"\r") for l in (lines[i : i + 4] + [b""] * 4)[:4]) for i in range(0, len(lines), 4)]


def oracle_readqc(files1, files2=()):
    """The ReadQC tool by the oracle (src/ReadQC/main.cpp:58-101): every entry of every file is validated and goes through
    StatisticsReads::update; returns the accumulators as a dict. Raises ValueError(code) on the first invalid entry."""
    lib = oracle_lib()
    st = _SpoQc()
    for i, f1 in enumerate(files1):
        for direction, path in ((0, f1),) + (((1, files2[i]),) if i < len(files2) else ()):
            for h, b, h2, q in read_fastq4(path):
                code = lib.spo_validate_entry(h, len(h), b, len(b), h2, len(h2), q, len(q))
                if code:
                    raise ValueError(code)
                lib.spo_qc_update_read(b, q, len(b), direction, C.byref(st))
    return _qc_to_dict(st)


def oracle_fastq_trim(records, start=0, end=0, max_bases=0, max_len=0):
    """The FastqTrim tool by the oracle (src/FastqTrim/main.cpp:47-77) on (header, bases, header2, qualities) records -> FASTQ text."""
    lib = oracle_lib()
    out = []
    first, count = C.c_int(0), C.c_int(0)
    for h, b, h2, q in records:
        if lib.spo_fastq_trim(len(b), start, end, max_bases, max_len, C.byref(first), C.byref(count)):
            s, n = first.value, count.value
            out.append(h + b"\n" + b[s : s + n] + b"\n" + h2 + b"\n" + q[s : s + n] + b"\n")
    return b"".join(out)


def qc_metrics(d):
    """The eight paired-end quality parameters of StatisticsReads::getResult (src/cppNGS/StatisticsReads.cpp:140-200) as the strings
    a qcML file holds (QCValue::toString: integers as such, doubles with two decimals)."""
    total_reads = d["reads_forward"] + d["reads_reverse"]
    pile = d["pileup"]
    c_base_n = int(pile[:, 4].sum())
    c_base_gc = int(pile[:, 1].sum() + pile[:, 2].sum())
    bases_total = int(pile.sum())
    keys = [int(i) for i in np.nonzero(d["read_lengths"])[0]]
    lengths = ", ".join(str(k) for k in keys) if len(keys) < 4 else f"{keys[0]}-{keys[-1]}"
    def f2(v):  # QString::number(v, 'f', 2): an exact tie goes up (double-conversion), not to even like printf -- 0.125 -> "0.13"
        from decimal import ROUND_HALF_UP, Decimal

        return str(Decimal(float(v)).quantize(Decimal("0.01"), rounding=ROUND_HALF_UP))

    return [
        ("read count", str(total_reads)),
        ("read length", lengths),
        ("bases sequenced (MB)", f2(d["bases_sequenced"] / 1000000.0)),
        ("Q20 read percentage", f2(100.0 * d["read_q20"] / total_reads)),
        ("Q20 base percentage", f2(100.0 * d["base_q20"] / bases_total)),
        ("Q30 base percentage", f2(100.0 * d["base_q30"] / bases_total)),
        ("no base call percentage", f2(100.0 * c_base_n / bases_total)),
        ("gc content percentage", f2(100.0 * c_base_gc / (bases_total - c_base_n))),
    ]




_oracle = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    return os.path.join(ORACLE_DIR, "build")


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(build_oracle(), "libseqpurge_oracle.so")
        lib = C.CDLL(path)
        lib.spo_trim_batch.argtypes = [C.POINTER(_SpoParams)] + [C.c_void_p] * 6 + [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        lib.spo_trim_batch.restype = None
        lib.spo_qc_update_batch.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int64, C.POINTER(_SpoQc)]
        lib.spo_qc_update_batch.restype = None
        lib.spo_qc_update_read.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(_SpoQc)]
        lib.spo_qc_update_read.restype = None
        lib.spo_validate_entry.argtypes = [C.c_char_p, C.c_int] * 4
        lib.spo_validate_entry.restype = C.c_int
        lib.spo_fastq_trim.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_int)] * 2
        lib.spo_fastq_trim.restype = C.c_int
        lib.spo_match_probability.argtypes = [C.c_double, C.c_int, C.c_int]
        lib.spo_match_probability.restype = C.c_double
        lib.spo_factorial.argtypes = [C.c_int]
        lib.spo_factorial.restype = C.c_double
        lib.spo_trim_quality.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int]
        lib.spo_trim_quality.restype = C.c_int
        lib.spo_trim_n.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_int]
        lib.spo_trim_n.restype = C.c_int
        _oracle = lib
    return _oracle


def oracle_params(a1=DEFAULT_A1, a2=DEFAULT_A2, adapter_overlap=10, match_perc=80.0, mep=1e-6, qcut=15, qwin=5, qoff=33, ncut=7, ec=False):
    b1, b2 = a1.encode(), a2.encode()
    p = _SpoParams(b1, len(b1), b2, len(b2), min(20, len(b1), len(b2)), adapter_overlap, match_perc, mep, qcut, qwin, qoff, ncut, int(ec))
    p._keep = (b1, b2)
    return p


class Batch:
    """SoA batch in the slot layout of the C ABI: fixed-stride uint8 rows + uint16 lengths."""

    def __init__(self, n, stride):
        self.n, self.stride = n, stride
        cap = (n + 7) // 8 * 8
        self.bases1 = np.zeros((cap, stride), np.uint8)
        self.quals1 = np.zeros((cap, stride), np.uint8)
        self.bases2 = np.zeros((cap, stride), np.uint8)
        self.quals2 = np.zeros((cap, stride), np.uint8)
        self.len1 = np.zeros(cap, np.uint16)
        self.len2 = np.zeros(cap, np.uint16)

    def copy(self):
        b = Batch(self.n, self.stride)
        for k in ("bases1", "quals1", "bases2", "quals2", "len1", "len2"):
            getattr(b, k)[...] = getattr(self, k)
        return b

    def set_pair(self, i, r1, q1, r2, q2):
        assert len(r1) == len(q1) and len(r2) == len(q2)
        self.bases1[i, : len(r1)] = np.frombuffer(r1, np.uint8)
        self.quals1[i, : len(q1)] = np.frombuffer(q1, np.uint8)
        self.bases2[i, : len(r2)] = np.frombuffer(r2, np.uint8)
        self.quals2[i, : len(q2)] = np.frombuffer(q2, np.uint8)
        self.len1[i], self.len2[i] = len(r1), len(r2)


def oracle_trim(batch, threads=1, **params):
    """Run the CPU oracle on a Batch. Returns (records, ec_stats or None); with ec=True the batch rows are edited in place."""
    lib = oracle_lib()
    p = oracle_params(**params)
    out = np.zeros(batch.n, RESULT_DTYPE)
    ec = _SpoEc() if params.get("ec") else None
    lib.spo_trim_batch(C.byref(p), batch.bases1.ctypes.data, batch.quals1.ctypes.data, batch.bases2.ctypes.data, batch.quals2.ctypes.data,
                       batch.len1.ctypes.data, batch.len2.ctypes.data, batch.stride, batch.n, out.ctypes.data, C.addressof(ec) if ec else None, threads)
    ecd = None
    if ec:
        ecd = {k: np.array(getattr(ec, k), dtype=np.int64) for k in ("mismatch_r1", "mismatch_r2", "errors_per_read")}
    return out, ecd


def read_fastq(path):
    """4-line FASTQ records (headers, bases, quals as bytes); tolerates a missing last quality line like the reference's reader."""
    with gzip.open(path, "rb") as f:
        lines = f.read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    recs = []
    for i in range(0, len(lines), 4):
        chunk = [l.rstrip(b"\r") for l in lines[i : i + 4]] + [b""] * 4
        recs.append((chunk[0], chunk[1], chunk[3]))
    return recs


def golden_batch(i1, i2, stride=None):
    r1 = read_fastq(os.path.join(GOLDEN, f"SeqPurge_in{i1}.fastq.gz"))
    r2 = read_fastq(os.path.join(GOLDEN, f"SeqPurge_in{i2}.fastq.gz"))
    assert len(r1) == len(r2)
    maxlen = max(max(len(r[1]) for r in r1), max(len(r[1]) for r in r2))
    if stride is None:
        stride = (maxlen + 15) // 16 * 16
    b = Batch(len(r1), stride)
    for i, (a, c) in enumerate(zip(r1, r2)):
        b.set_pair(i, a[1], a[2], c[1], c[2])
    return b


def random_batch(n, L, seed, insert_mean=None, insert_sd=None, error_rate=0.01, n_rate=0.001, lowq_tail=5.0, a1=DEFAULT_A1, a2=DEFAULT_A2,
                 ragged=False, n_runs=0.0, stride=None):
    """Numpy restatement of the synthetic model (fragment + adapters + filler, errors, Ns, low-quality tails)."""
    rng = np.random.default_rng(seed)
    if stride is None:
        stride = (L + 15) // 16 * 16
    b = Batch(n, stride)
    comp = {65: 84, 67: 71, 71: 67, 84: 65, 78: 78}
    acgt = np.frombuffer(b"ACGT", np.uint8)
    mu = insert_mean if insert_mean is not None else 1.2 * L
    sd = insert_sd if insert_sd is not None else 0.5 * L
    for i in range(n):
        ins = int(np.clip(round(rng.normal(mu, sd)), 1, 4 * L))
        frag = acgt[rng.integers(0, 4, ins)]
        rc = np.array([comp[x] for x in frag[::-1]], np.uint8)
        reads = []
        for frag_o, ad in ((frag, a1), (rc, a2)):
            ln = L if not ragged else int(rng.integers(0, L + 1))
            ad_b = np.frombuffer(ad.encode(), np.uint8)
            seq = np.concatenate([frag_o, ad_b, acgt[rng.integers(0, 4, L)]])[:ln].copy()
            err = rng.random(ln) < error_rate
            seq[err] = acgt[rng.integers(0, 4, int(err.sum()))]
            seq[rng.random(ln) < n_rate] = 78
            if n_runs > 0 and ln > 20 and rng.random() < n_runs:
                s = int(rng.integers(0, ln - 12))
                seq[s : s + int(rng.integers(7, 12))] = 78
            q = np.full(ln, 73, np.uint8)
            q[rng.random(ln) < 0.05] = 40
            t = int(rng.exponential(lowq_tail)) if lowq_tail > 0 else 0
            if t > 0:
                q[max(0, ln - t):] = 35
            reads.append((seq.tobytes(), q.tobytes()))
        b.set_pair(i, reads[0][0], reads[0][1], reads[1][0], reads[1][1])
    return b


def synth_batch_numpy(n, L, seed=1, stride=None, insert_mean=250.0, insert_sd=80.0, insert_min=1, insert_max=2000, error_rate=0.001, n_rate=1e-4,
                      lowq_tail_mean=3.0, n_run_rate=0.0, binned_quals=False, a1=DEFAULT_A1, a2=DEFAULT_A2):
    """Vectorised numpy generator of the synthetic model of SURVEY.md section 8d (the parameters of seqpurge_b200.SynthConfig): fragment
    of insert ~ N(mu, sigma) clipped, read 1 = fragment + adapter 1 + random filler, read 2 = revcomp(fragment) + adapter 2 + filler,
    i.i.d. substitutions and N, qualities 'I' with a geometric low-quality tail ('#'), optional N runs and binned qualities. The same
    distribution as the device generator, not the same random stream (used where no CUDA library may be loaded: bench.py's CPU arm)."""
    rng = np.random.default_rng(seed)
    if stride is None:
        stride = (L + 1) // 2 * 2
    b = Batch(n, stride)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    comp = np.zeros(256, np.uint8)
    for x, y in zip(b"ACGTN", b"TGCAN"):
        comp[x] = y
    ins = np.clip(np.rint(rng.normal(insert_mean, insert_sd, n)), insert_min, insert_max).astype(np.int64)
    ins = np.minimum(ins, 2 * L)  # longer fragments: the two reads do not overlap, their bases are independent either way
    frag = acgt[rng.integers(0, 4, (n, 2 * L))]
    col = np.arange(L)[None, :]
    ad1 = np.concatenate([np.frombuffer(a1.encode(), np.uint8), acgt[rng.integers(0, 4, L)]])[: L]
    ad2 = np.concatenate([np.frombuffer(a2.encode(), np.uint8), acgt[rng.integers(0, 4, L)]])[: L]
    rel = col - ins[:, None]  # position inside adapter + filler where the read runs past the fragment
    inside = rel < 0
    filler = acgt[rng.integers(0, 4, (n, L))]
    r1 = np.where(inside, frag[:, :L], np.where(rel < len(a1), ad1[np.clip(rel, 0, L - 1)], filler))
    idx = np.clip(ins[:, None] - 1 - col, 0, 2 * L - 1)
    r2 = np.where(inside, comp[np.take_along_axis(frag, idx, axis=1)], np.where(rel < len(a2), ad2[np.clip(rel, 0, L - 1)], filler[:, ::-1]))
    for r in (r1, r2):
        err = rng.random((n, L)) < error_rate
        r[err] = acgt[rng.integers(0, 4, int(err.sum()))]
        r[rng.random((n, L)) < n_rate] = 78
        if n_run_rate > 0:
            rows = np.nonzero(rng.random(n) < n_run_rate)[0]
            for i in rows:
                st = int(rng.integers(0, L - 12))
                r[i, st : st + int(rng.integers(7, 12))] = 78
    hi, lo = (ord("F"), ord("#")) if binned_quals else (ord("I"), ord("#"))
    for r, q, ln in ((r1, b.quals1, b.len1), (r2, b.quals2, b.len2)):
        qq = np.full((n, L), hi, np.uint8)
        if binned_quals:
            u = rng.random((n, L))
            qq[u < 0.08] = ord(":")
            qq[u < 0.02] = ord(",")
        if lowq_tail_mean > 0:
            t = rng.geometric(1.0 / (1.0 + lowq_tail_mean), n) - 1
            qq[col >= (L - t)[:, None]] = lo
        q[:n, :L] = qq
        ln[:n] = L
    b.bases1[:n, :L] = r1
    b.bases2[:n, :L] = r2
    return b
