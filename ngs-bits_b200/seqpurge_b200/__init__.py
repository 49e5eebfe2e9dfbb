"""seqpurge_b200 -- ctypes binding of libseqpurge_b200.so (the C ABI in include/seqpurge_b200.h).

Host-side mirror, in Python, of the seam the library replaces in imgag/ngs-bits:
``TrimmingParameters`` (src/SeqPurge/Auxilary.h:100-133) and the job hand-off of
``ThreadCoordinator::analyze`` -> ``AnalysisWorker::run`` (src/SeqPurge/ThreadCoordinator.cpp:92-98,
src/SeqPurge/AnalysisWorker.cpp:79-457).  torch is only plumbing here (device memory, streams, events);
all trimming work happens in the CUDA library, and importing this module fails loudly if the library has
not been built -- there is no Python or CPU fallback.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPG_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libseqpurge_b200.so")  # SPG_LIB: another build of the same library (kernel experiments)

MAXLEN = 1000
F_INSERT, F_ADAPTER, F_Q1, F_Q2, F_N1, F_N2 = 1, 2, 4, 8, 16, 32
PAIR_OK, PAIR_BAD_BASE_R2, PAIR_TOO_LONG, PAIR_BAD_BASE_EC = 0, 1, 2, 3
OPT_FORCE_BYTEWISE, OPT_GRID_CTAS_PER_SM, OPT_MIN_BLOCKS, OPT_TILE_PAIRS, OPT_STAGES, OPT_FULL_LEN, OPT_KERNEL, OPT_SEED_SCAN, OPT_ZERO_COPY_QUALS, OPT_N_LANES, OPT_QUAL_TAILS = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11
QTAIL = 16  # SPG_QTAIL
KERNEL_AUTO, KERNEL_WARP_PER_PAIR, KERNEL_LANE_PER_PAIR = 0, 1, 2

RESULT_DTYPE = np.dtype([("len1", "<u2"), ("len2", "<u2"), ("best_offset", "<i2"), ("flags", "u1"), ("status", "u1")])
assert RESULT_DTYPE.itemsize == 8

DEFAULT_A1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA"  # src/SeqPurge/main.cpp:25
DEFAULT_A2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT"  # src/SeqPurge/main.cpp:26


class SeqPurgeError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). seqpurge_b200 has no CPU fallback."
    )
_lib = C.CDLL(LIB_PATH)


class _Params(C.Structure):
    _fields_ = [
        ("a1", C.c_char_p), ("a1_len", C.c_int32), ("a2", C.c_char_p), ("a2_len", C.c_int32),
        ("adapter_overlap", C.c_int32), ("match_perc", C.c_double), ("mep", C.c_double),
        ("qcut", C.c_int32), ("qwin", C.c_int32), ("qoff", C.c_int32), ("ncut", C.c_int32), ("ec", C.c_int32), ("qc", C.c_int32),
    ]


class _QcStats(C.Structure):
    _fields_ = [
        ("reads_forward", C.c_int64), ("reads_reverse", C.c_int64), ("bases_sequenced", C.c_int64), ("read_q20", C.c_int64),
        ("base_q20", C.c_int64), ("base_q30", C.c_int64), ("errors", C.c_int64),
        ("read_lengths", C.c_int64 * MAXLEN), ("pileup", (C.c_int64 * 5) * MAXLEN),
        ("qsum_forward", C.c_int64 * MAXLEN), ("qsum_reverse", C.c_int64 * MAXLEN),
        ("base_qualities", C.c_int64 * 100), ("read_qualities", C.c_int64 * 100), ("qscore_dist_forward", C.c_int64 * 60), ("qscore_dist_reverse", C.c_int64 * 60),
    ]


def qc_stats_to_dict(st):
    """ctypes spg_qc_stats / spo_qc_stats -> plain python (ints and numpy arrays)."""
    d = {k: int(getattr(st, k)) for k in ("reads_forward", "reads_reverse", "bases_sequenced", "read_q20", "base_q20", "base_q30", "errors")}
    d["read_lengths"] = np.array(st.read_lengths, dtype=np.int64)
    d["pileup"] = np.array([list(row) for row in st.pileup], dtype=np.int64)
    d["qsum_forward"] = np.array(st.qsum_forward, dtype=np.int64)
    d["qsum_reverse"] = np.array(st.qsum_reverse, dtype=np.int64)
    for k in ("base_qualities", "read_qualities", "qscore_dist_forward", "qscore_dist_reverse"):  # the accumulators behind the qcML plots
        d[k] = np.array(getattr(st, k), dtype=np.int64)
    return d


class _SlotView(C.Structure):
    _fields_ = [
        ("bases1", C.c_void_p), ("quals1", C.c_void_p), ("bases2", C.c_void_p), ("quals2", C.c_void_p),
        ("len1", C.c_void_p), ("len2", C.c_void_p), ("stride", C.c_int32), ("max_pairs", C.c_int32),
    ]


class _EcStats(C.Structure):
    _fields_ = [("mismatch_r1", C.c_int64 * MAXLEN), ("mismatch_r2", C.c_int64 * MAXLEN), ("errors_per_read", C.c_int64 * MAXLEN)]


class _SynthConfig(C.Structure):
    _fields_ = [
        ("read_len", C.c_int32), ("insert_mean", C.c_float), ("insert_sd", C.c_float), ("insert_min", C.c_int32),
        ("insert_max", C.c_int32), ("error_rate", C.c_float), ("n_rate", C.c_float), ("lowq_tail_mean", C.c_float),
        ("n_run_rate", C.c_float), ("binned_quals", C.c_int32), ("seed", C.c_uint64), ("a1", C.c_char_p), ("a2", C.c_char_p),
    ]


_lib.spg_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Params), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int]
_lib.spg_create.restype = C.c_int
_lib.spg_slot_buffers.argtypes = [C.c_void_p, C.c_int, C.POINTER(_SlotView)]
_lib.spg_slot_buffers.restype = C.c_int
_lib.spg_slot_qtails.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
_lib.spg_slot_qtails.restype = C.c_int
_lib.spg_submit.argtypes = [C.c_void_p, C.c_int, C.c_int]
_lib.spg_submit.restype = C.c_int
_lib.spg_wait.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
_lib.spg_wait.restype = C.c_int
_lib.spg_trim_device.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
_lib.spg_trim_device.restype = C.c_int
_lib.spg_ec_stats_get.argtypes = [C.c_void_p, C.POINTER(_EcStats)]
_lib.spg_ec_stats_get.restype = C.c_int
_lib.spg_qc_device.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_int64, C.c_void_p]
_lib.spg_qc_device.restype = C.c_int
_lib.spg_qc_stats_get.argtypes = [C.c_void_p, C.POINTER(_QcStats)]
_lib.spg_qc_stats_get.restype = C.c_int
_lib.spg_last_error.argtypes = [C.c_void_p]
_lib.spg_last_error.restype = C.c_char_p
_lib.spg_destroy.argtypes = [C.c_void_p]
_lib.spg_destroy.restype = None
_lib.spg_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
_lib.spg_set_option.restype = C.c_int
_lib.spg_get_option.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
_lib.spg_get_option.restype = C.c_int
_lib.spg_last_kernel.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
_lib.spg_last_kernel.restype = C.c_int
_lib.spg_launch_count.argtypes = [C.c_void_p]
_lib.spg_launch_count.restype = C.c_int64
_lib.spg_synth_device.argtypes = [C.c_int, C.POINTER(_SynthConfig), C.c_int64, C.c_int64] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p]
_lib.spg_synth_device.restype = C.c_int


@dataclass
class TrimmingParameters:
    """Fields of the reference's TrimmingParameters that the analysis step reads (Auxilary.h:100-133), with the
    command-line defaults of src/SeqPurge/main.cpp:25-43."""

    a1: str = DEFAULT_A1
    a2: str = DEFAULT_A2
    adapter_overlap: int = 10
    match_perc: float = 80.0
    mep: float = 0.000001
    qcut: int = 15
    qwin: int = 5
    qoff: int = 33
    ncut: int = 7
    ec: bool = False
    qc: int = 0  # 1: also accumulate the read statistics; 2: with the character checks of FastqEntry::validate (ReadQC)

    def _c(self):
        a1, a2 = self.a1.encode(), self.a2.encode()
        p = _Params(a1, len(a1), a2, len(a2), self.adapter_overlap, self.match_perc, self.mep, self.qcut, self.qwin, self.qoff, self.ncut, int(self.ec), int(self.qc))
        p._keep = (a1, a2)
        return p


@dataclass
class SynthConfig:
    """Synthetic read-pair stream."""

    read_len: int = 150
    insert_mean: float = 250.0
    insert_sd: float = 80.0
    insert_min: int = 1
    insert_max: int = 2000
    error_rate: float = 0.001
    n_rate: float = 1e-4
    lowq_tail_mean: float = 3.0
    n_run_rate: float = 0.0
    binned_quals: bool = False
    seed: int = 0x5E9B200

    def _c(self):
        return _SynthConfig(self.read_len, self.insert_mean, self.insert_sd, self.insert_min, self.insert_max, self.error_rate, self.n_rate,
                            self.lowq_tail_mean, self.n_run_rate, int(self.binned_quals), self.seed, None, None)


def _np_view(ptr, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    buf = (C.c_uint8 * n).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class Slot:
    """Pinned host SoA of one job (one AnalysisJob of the reference's pool)."""

    def __init__(self, view, qtail1=None, qtail2=None):
        self.stride, self.max_pairs = view.stride, view.max_pairs
        if qtail1:  # the optional quality tails (spg_slot_qtails, OPT_QUAL_TAILS)
            self.qtail1 = _np_view(qtail1, (view.max_pairs, QTAIL), np.uint8)
            self.qtail2 = _np_view(qtail2, (view.max_pairs, QTAIL), np.uint8)
        shp = (view.max_pairs, view.stride)
        self.bases1 = _np_view(view.bases1, shp, np.uint8)
        self.quals1 = _np_view(view.quals1, shp, np.uint8)
        self.bases2 = _np_view(view.bases2, shp, np.uint8)
        self.quals2 = _np_view(view.quals2, shp, np.uint8)
        self.len1 = _np_view(view.len1, (view.max_pairs,), np.uint16)
        self.len2 = _np_view(view.len2, (view.max_pairs,), np.uint16)

    def fill_qtails(self, n):
        """Writes the last QTAIL qualities of the first n reads of both planes into qtail1 / qtail2 (what a stager does while it copies a
        read's quality string; shorter reads right-aligned). Vectorised for tests and bench.py."""
        cols = np.arange(QTAIL, dtype=np.int64)[None, :]
        for q, ln, qt in ((self.quals1, self.len1, self.qtail1), (self.quals2, self.len2, self.qtail2)):
            idx = ln[:n].astype(np.int64)[:, None] - QTAIL + cols  # position of tail byte j in its row (negative: in front of the read)
            ok = idx >= 0
            rows = np.arange(n, dtype=np.int64)[:, None]
            qt[:n] = np.where(ok, q[:n][rows, np.where(ok, idx, 0)], 0)


class _FqConfig(C.Structure):
    _fields_ = [("n_slots", C.c_int32), ("max_pairs", C.c_int32), ("max_len", C.c_int32), ("text_cap", C.c_int64), ("min_len", C.c_int32), ("singles", C.c_int32),
                ("stats_only", C.c_int32), ("single_end", C.c_int32), ("validate", C.c_int32), ("fixed_trim", C.c_int32),
                ("trim_start", C.c_int32), ("trim_end", C.c_int32), ("trim_len", C.c_int32), ("trim_max_len", C.c_int32)]


class _FqInput(C.Structure):
    _fields_ = [("text1", C.c_void_p), ("text2", C.c_void_p), ("cap", C.c_int64)]


class _FqOutput(C.Structure):
    _fields_ = [("n_pairs", C.c_int32), ("records1", C.c_int32), ("records2", C.c_int32), ("consumed1", C.c_int64), ("consumed2", C.c_int64),
                ("out", C.c_void_p * 4), ("out_bytes", C.c_int64 * 4), ("results", C.c_void_p), ("len1", C.c_void_p), ("len2", C.c_void_p),
                ("frame_status", C.c_void_p), ("error_pair", C.c_int32), ("max_len", C.c_int32), ("invalid_chars", C.c_int32), ("stats", C.c_void_p)]


class _FqStats(C.Structure):
    _fields_ = [("reads_trimmed_insert", C.c_int64), ("reads_trimmed_adapter", C.c_int64), ("reads_trimmed_q", C.c_int64), ("reads_trimmed_n", C.c_int64),
                ("reads_removed", C.c_int64), ("bases_remaining", C.c_int64 * MAXLEN), ("trimmed_bases_by_length", C.c_int64 * MAXLEN)]


_lib.spg_fq_open.argtypes = [C.c_void_p, C.POINTER(_FqConfig), C.POINTER(C.c_void_p)]
_lib.spg_fq_open.restype = C.c_int
_lib.spg_fq_buffers.argtypes = [C.c_void_p, C.c_int, C.POINTER(_FqInput)]
_lib.spg_fq_buffers.restype = C.c_int
_lib.spg_fq_submit.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int]
_lib.spg_fq_submit.restype = C.c_int
_lib.spg_fq_wait.argtypes = [C.c_void_p, C.c_int, C.POINTER(_FqOutput)]
_lib.spg_fq_wait.restype = C.c_int
_lib.spg_fq_consensus_get.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
_lib.spg_fq_consensus_get.restype = C.c_int
_lib.spg_fq_close.argtypes = [C.c_void_p]
_lib.spg_fq_close.restype = None

FQ_OK, FQ_HEADER_MISMATCH, FQ_LENGTH_MISMATCH, FQ_TOO_LONG = 0, 1, 2, 3


class Engine:
    """One spg_ctx.  ``devices``: CUDA device ids; slot s runs on devices[s % len(devices)]."""

    def __init__(self, params=None, devices=(0,), n_slots=0, max_pairs=0, max_len=0):
        self.params = params or TrimmingParameters()
        self._p = self.params._c()
        self._h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        rc = _lib.spg_create(C.byref(self._h), C.byref(self._p), devs, len(devices), n_slots, max_pairs, max_len)
        if rc != 0:
            raise SeqPurgeError(f"spg_create failed ({rc}): {_lib.spg_last_error(None).decode()}")
        self.n_slots, self.max_pairs, self.max_len = n_slots, max_pairs, max_len
        self._slots = {}
        self._pending = {}

    def _check(self, rc, what):
        if rc != 0:
            raise SeqPurgeError(f"{what} failed ({rc}): {_lib.spg_last_error(self._h).decode()}")

    def slot(self, i):
        if i not in self._slots:
            v = _SlotView()
            self._check(_lib.spg_slot_buffers(self._h, i, C.byref(v)), "spg_slot_buffers")
            t1, t2 = C.c_void_p(), C.c_void_p()
            self._check(_lib.spg_slot_qtails(self._h, i, C.byref(t1), C.byref(t2)), "spg_slot_qtails")
            self._slots[i] = Slot(v, t1.value, t2.value)
        return self._slots[i]

    def submit(self, slot, n_pairs):
        self._check(_lib.spg_submit(self._h, slot, n_pairs), "spg_submit")
        self._pending[slot] = n_pairs

    def wait(self, slot):
        """Returns a structured array (RESULT_DTYPE) viewing the slot's pinned result records."""
        res = C.c_void_p()
        self._check(_lib.spg_wait(self._h, slot, C.byref(res)), "spg_wait")
        n = self._pending.pop(slot)
        if n == 0:
            return np.zeros(0, RESULT_DTYPE)
        return _np_view(res.value, (n,), RESULT_DTYPE)

    def trim_device(self, bases1, quals1, bases2, quals2, len1, len2, results, n_pairs=None, device_index=0, stream=None):
        """Device-resident batch (torch CUDA tensors): rows uint8 [n, stride], lens int16/uint16 [>= round_up(n, 8)],
        results uint8 [n, 8] (or int64 [n]). Queued on `stream` (default: torch's current stream)."""
        import torch

        n = bases1.shape[0] if n_pairs is None else n_pairs
        stride = bases1.stride(0)
        if stream is None:
            stream = torch.cuda.current_stream(bases1.device).cuda_stream
        rc = _lib.spg_trim_device(self._h, device_index, bases1.data_ptr(), quals1.data_ptr(), bases2.data_ptr(), quals2.data_ptr(), len1.data_ptr(),
                                  len2.data_ptr(), stride, n, results.data_ptr(), stream)
        self._check(rc, "spg_trim_device")

    def ec_stats(self):
        st = _EcStats()
        self._check(_lib.spg_ec_stats_get(self._h, C.byref(st)), "spg_ec_stats_get")
        return {k: np.array(getattr(st, k), dtype=np.int64) for k in ("mismatch_r1", "mismatch_r2", "errors_per_read")}

    def qc_device(self, bases1, quals1, bases2, quals2, len1, len2, n_pairs=None, device_index=0, stream=None):
        """Raw-read statistics (-qc) of a device-resident batch (torch CUDA tensors as in trim_device), added to the context's accumulators."""
        import torch

        n = bases1.shape[0] if n_pairs is None else n_pairs
        if stream is None:
            stream = torch.cuda.current_stream(bases1.device).cuda_stream
        rc = _lib.spg_qc_device(self._h, device_index, bases1.data_ptr(), quals1.data_ptr(), bases2.data_ptr(), quals2.data_ptr(), len1.data_ptr(),
                                len2.data_ptr(), bases1.stride(0), n, stream)
        self._check(rc, "spg_qc_device")

    def qc_stats(self):
        """Accumulated -qc statistics of all devices as a dict (scalars and numpy arrays)."""
        st = _QcStats()
        self._check(_lib.spg_qc_stats_get(self._h, C.byref(st)), "spg_qc_stats_get")
        return qc_stats_to_dict(st)

    def set_option(self, option, value):
        self._check(_lib.spg_set_option(self._h, option, value), "spg_set_option")

    def get_option(self, option):
        v = C.c_int(0)
        self._check(_lib.spg_get_option(self._h, option, C.byref(v)), "spg_get_option")
        return v.value

    @property
    def zero_copy_quals(self):
        """True if submitted slots leave their quality planes in pinned host memory when the lane-per-pair kernel runs."""
        return bool(self.get_option(OPT_ZERO_COPY_QUALS))

    @property
    def launch_count(self):
        return int(_lib.spg_launch_count(self._h))

    @property
    def last_kernel(self):
        """Name of the trimming kernel instantiation the last launch ran (spg_last_kernel)."""
        buf = C.create_string_buffer(128)
        _lib.spg_last_kernel(self._h, buf, 128)
        return buf.value.decode()

    def close(self):
        if self._h:
            _lib.spg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FastqChunk:
    """What spg_fq_wait returns for one chunk (copies of the pinned buffers, so the slot can be reused)."""

    def __init__(self, o):
        n = o.n_pairs
        self.n_pairs, self.records, self.consumed = n, (o.records1, o.records2), (o.consumed1, o.consumed2)
        self.out = [C.string_at(o.out[k], o.out_bytes[k]) if o.out_bytes[k] else b"" for k in range(4)]
        self.results = _np_view(o.results, (n,), RESULT_DTYPE).copy() if n and o.results else np.zeros(0, RESULT_DTYPE)
        self.len1 = _np_view(o.len1, (n,), np.uint16).copy() if n else np.zeros(0, np.uint16)
        self.len2 = _np_view(o.len2, (n,), np.uint16).copy() if n else np.zeros(0, np.uint16)
        self.frame_status = _np_view(o.frame_status, (n,), np.uint8).copy() if n else np.zeros(0, np.uint8)
        self.error_pair, self.max_len, self.invalid_chars = o.error_pair, o.max_len, o.invalid_chars
        self.stats = None  # summary counters of the chunk (spg_fq_stats) as a dict
        if o.stats:
            st = _FqStats.from_address(o.stats)
            self.stats = {k: int(getattr(st, k)) for k in ("reads_trimmed_insert", "reads_trimmed_adapter", "reads_trimmed_q", "reads_trimmed_n", "reads_removed")}
            self.stats["bases_remaining"] = np.array(st.bases_remaining, dtype=np.int64)
            self.stats["trimmed_bases_by_length"] = np.array(st.trimmed_bases_by_length, dtype=np.int64)


class FastqStream:
    """FASTQ text in, FASTQ text out on the device (spg_fq_*): framing, trimming, routing by min_len, record layout."""

    def __init__(self, engine, n_slots=2, max_pairs=65536, max_len=160, text_cap=32 << 20, min_len=30, singles=False, stats_only=False, single_end=False,
                 validate=False, fixed_trim=None):
        """stats_only / single_end / validate: the ReadQC form of the stream (framing and read statistics only, see spg_fq_config)."""
        self.engine = engine
        self._h = C.c_void_p()
        ft = fixed_trim or (0, 0, 0, 0)  # (start, end, len, max_len) of the FastqTrim form of the stream
        cfg = _FqConfig(n_slots, max_pairs, max_len, text_cap, min_len, int(singles), int(stats_only), int(single_end), int(validate), int(fixed_trim is not None), *ft)
        engine._check(_lib.spg_fq_open(engine._h, C.byref(cfg), C.byref(self._h)), "spg_fq_open")
        self.text_cap, self.n_slots = text_cap, n_slots

    def buffers(self, slot):
        v = _FqInput()
        self.engine._check(_lib.spg_fq_buffers(self._h, slot, C.byref(v)), "spg_fq_buffers")
        return _np_view(v.text1, (v.cap,), np.uint8), _np_view(v.text2, (v.cap,), np.uint8)

    def submit(self, slot, text1, text2, final1=True, final2=True):
        b1, b2 = self.buffers(slot)
        b1[: len(text1)] = np.frombuffer(text1, np.uint8)
        b2[: len(text2)] = np.frombuffer(text2, np.uint8)
        self.engine._check(_lib.spg_fq_submit(self._h, slot, len(text1), len(text2), int(final1), int(final2)), "spg_fq_submit")

    def wait(self, slot):
        o = _FqOutput()
        self.engine._check(_lib.spg_fq_wait(self._h, slot, C.byref(o)), "spg_fq_wait")
        return FastqChunk(o)

    def consensus(self):
        counts = np.zeros((2, 40, 5), np.int64)
        unknown = C.c_int32(0)
        self.engine._check(_lib.spg_fq_consensus_get(self._h, counts.ctypes.data, C.byref(unknown)), "spg_fq_consensus_get")
        return counts, bool(unknown.value)

    def close(self):
        if self._h and self.engine._h:  # spg_destroy closes the streams that are still attached
            _lib.spg_fq_close(self._h)
        self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def synth_device(cfg, first_pair, n_pairs, bases1, quals1, bases2, quals2, len1, len2, device_id=0, stream=None):
    """Fill torch CUDA tensors (rows uint8 [n, stride]) with pairs [first_pair, first_pair+n_pairs) of the stream `cfg`."""
    import torch

    if stream is None:
        stream = torch.cuda.current_stream(bases1.device).cuda_stream
    c = cfg._c()
    rc = _lib.spg_synth_device(device_id, C.byref(c), first_pair, n_pairs, bases1.data_ptr(), quals1.data_ptr(), bases2.data_ptr(), quals2.data_ptr(),
                               len1.data_ptr(), len2.data_ptr(), bases1.stride(0), stream)
    if rc != 0:
        raise SeqPurgeError(f"spg_synth_device failed ({rc})")


def results_from_tensor(t):
    """torch uint8 [n, 8] (CPU) -> structured array."""
    return t.cpu().numpy().reshape(-1).view(RESULT_DTYPE)
