"""GPU parity tests (run with -m gpu on the B200 box): the CUDA engine, called through the C ABI, against the CPU oracle on
the same inputs -- bit-exact result records (integer/index work).

Inputs: the reference's own SeqPurge fixtures (tests/golden, flags of src/tools-TEST/SeqPurge_Test.cpp:100-208), seeded random
batches covering the edge cases of SURVEY.md appendix B (unequal and zero lengths, N runs, low-quality tails, overlaps > 170 bases,
bytes outside ACGTN, reads of up to 999 bases), and device-generated synthetic batches of the BASELINE configs.
"""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sp():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import __graft_entry__ as g

    g.build()
    import seqpurge_b200

    return seqpurge_b200


def gpu_trim(sp, batch, force_bytewise=False, n_slots=1, chunk=None, full_len=None, kernel=None, expect_kernel=None, seed_scan=None, qual_tails=False, **params):
    """Run a Batch through spg_submit/spg_wait (pinned slot, H2D, kernel, D2H). Returns records (+ edited batch, ec stats with ec)."""
    p = sp.TrimmingParameters(**params)
    chunk = chunk or batch.n
    eng = sp.Engine(p, devices=(0,), n_slots=n_slots, max_pairs=chunk, max_len=min(batch.stride, 999))
    if force_bytewise:
        eng.set_option(sp.OPT_FORCE_BYTEWISE, 1)
    if full_len is not None:  # kernel variant compiled for this read length (0: the general kernel)
        eng.set_option(sp.OPT_FULL_LEN, full_len)
    if kernel is not None:  # thread layout of the read-length variants (sp.KERNEL_WARP_PER_PAIR / sp.KERNEL_LANE_PER_PAIR)
        eng.set_option(sp.OPT_KERNEL, kernel)
    if seed_scan is not None:  # lane-per-pair kernel: exact-block filter in front of the adapter scans on / off
        eng.set_option(sp.OPT_SEED_SCAN, seed_scan)
    if qual_tails:  # the stager also fills the slots' quality tails (spg_slot_qtails): they are shipped instead of the quality planes
        eng.set_option(sp.OPT_QUAL_TAILS, 1)
    out = np.zeros(batch.n, sp.RESULT_DTYPE)
    edited = batch.copy() if params.get("ec") else None
    starts = list(range(0, batch.n, chunk))
    inflight = []
    for k, st in enumerate(starts):
        slot = k % n_slots
        if len(inflight) == n_slots:  # retire in submission order
            s0, st0, n0 = inflight.pop(0)
            out[st0 : st0 + n0] = eng.wait(s0)
            if edited is not None:
                _copy_back(eng.slot(s0), edited, st0, n0)
        n = min(chunk, batch.n - st)
        s = eng.slot(slot)
        assert s.stride == batch.stride
        for name in ("bases1", "quals1", "bases2", "quals2"):
            getattr(s, name)[:n] = getattr(batch, name)[st : st + n]
        s.len1[:n] = batch.len1[st : st + n]
        s.len2[:n] = batch.len2[st : st + n]
        if qual_tails:
            s.fill_qtails(n)
        eng.submit(slot, n)
        inflight.append((slot, st, n))
    for s0, st0, n0 in inflight:
        out[st0 : st0 + n0] = eng.wait(s0)
        if edited is not None:
            _copy_back(eng.slot(s0), edited, st0, n0)
    ec = eng.ec_stats() if params.get("ec") else None
    if expect_kernel is not None:
        assert expect_kernel in eng.last_kernel, eng.last_kernel
    eng.close()
    return out, edited, ec


def _copy_back(slot, edited, st, n):
    for name in ("bases1", "quals1", "bases2", "quals2"):
        getattr(edited, name)[st : st + n] = getattr(slot, name)[:n]


def assert_same(got, want, batch=None):
    a, b = got.view(np.uint64), want.view(np.uint64)
    if not np.array_equal(a, b):
        bad = np.nonzero(a != b)[0]
        i = int(bad[0])
        msg = f"{len(bad)} of {len(a)} records differ; first at pair {i}: gpu={got[i]} oracle={want[i]}"
        if batch is not None:
            msg += f"\nR1={batch.bases1[i, : batch.len1[i]].tobytes()}\nR2={batch.bases2[i, : batch.len2[i]].tobytes()}"
        raise AssertionError(msg)


GOLDEN_CASES = [
    ("test_01", 1, 2, dict(ncut=0, qcut=0)),
    ("test_02", 3, 4, dict(ncut=0, qcut=0)),
    ("test_03", 5, 6, dict(ncut=0, qcut=0)),
    ("test_04", 7, 8, dict(a1="CTGTCTCTTATACACATCT", a2="CTGTCTCTTATACACATCT", ncut=0, qcut=0)),
    ("test_05", 1, 2, dict(qcut=15, ncut=0)),
    ("test_06", 1, 2, dict(ncut=7, qcut=0)),
    ("test_07", 1, 2, dict(qcut=25)),
    ("test_08", 9, 10, dict()),
    ("test_09", 11, 12, dict()),
]


@pytest.mark.parametrize("case", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
@pytest.mark.parametrize("bytewise", [False, True], ids=["planes", "bytewise"])
def test_reference_fixtures(sp, case, bytewise):
    """Inputs and flags of the reference's tool tests; the oracle reproduces the reference's golden outputs on them
    (tests/test_oracle_golden.py), so equality with the oracle here is equality with the reference."""
    _, i1, i2, params = case
    batch = H.golden_batch(i1, i2)
    want, _ = H.oracle_trim(batch, **params)
    got, _, _ = gpu_trim(sp, batch, force_bytewise=bytewise, **params)
    assert_same(got, want, batch)


@pytest.mark.parametrize("bytewise", [False, True], ids=["planes", "bytewise"])
def test_reference_fixture_error_correction(sp, bytewise):
    """test_10 (-ec): edited bases/qualities, result records and the error histograms all match."""
    params = dict(ncut=0, qcut=0, ec=True)
    batch = H.golden_batch(1, 2)
    ref = batch.copy()
    want, want_ec = H.oracle_trim(ref, **params)
    got, edited, got_ec = gpu_trim(sp, batch, force_bytewise=bytewise, **params)
    assert_same(got, want, batch)
    for name in ("bases1", "quals1", "bases2", "quals2"):
        assert np.array_equal(getattr(edited, name)[: batch.n], getattr(ref, name)[: batch.n]), name
    for k in want_ec:
        assert np.array_equal(got_ec[k], want_ec[k]), k


def test_multi_slot_in_order(sp):
    """Several slots in flight, uneven last chunk: records come back in submission order (reference -threads 1 order)."""
    batch = H.golden_batch(5, 6)
    want, _ = H.oracle_trim(batch)
    got, _, _ = gpu_trim(sp, batch, n_slots=3, chunk=1000)
    assert_same(got, want, batch)


@pytest.mark.parametrize("L,seed", [(150, 1), (151, 2), (100, 3), (36, 4), (250, 5), (300, 6)])
def test_random_batches(sp, L, seed):
    batch = H.random_batch(3000, L, seed, error_rate=0.02, n_rate=0.002, lowq_tail=8.0)
    want, _ = H.oracle_trim(batch)
    got, _, _ = gpu_trim(sp, batch)
    assert_same(got, want, batch)
    assert (want["flags"] & 1).sum() > 100 and (want["flags"] & 2).sum() > 10  # both trimming modes exercised


@pytest.mark.parametrize("L,stride,n", [(150, 150, 1003), (101, 102, 77), (250, 250, 2001), (36, 36, 9)])
def test_tight_even_strides_and_ragged_last_tile(sp, L, stride, n):
    """Rows packed at an even stride that is not a multiple of 16, pair counts that are not a multiple of the tile size
    (the last TMA tile is ragged), with -ec so that the edited rows travel back as well."""
    batch = H.random_batch(n, L, 100 + L, error_rate=0.03, n_rate=0.002, lowq_tail=6.0, stride=stride)
    ref = batch.copy()
    want, want_ec = H.oracle_trim(ref, ec=True)
    got, edited, got_ec = gpu_trim(sp, batch, ec=True)
    assert_same(got, want, batch)
    for name in ("bases1", "quals1", "bases2", "quals2"):
        assert np.array_equal(getattr(edited, name)[:n], getattr(ref, name)[:n]), name
    for k in want_ec:
        assert np.array_equal(got_ec[k], want_ec[k]), k


def test_random_ragged_lengths_and_n_runs(sp):
    """len1 != len2, zero-length reads, injected N runs, -ncut/-qcut on (appendix B)."""
    batch = H.random_batch(4000, 150, 11, ragged=True, n_runs=0.2, n_rate=0.01, lowq_tail=20.0)
    want, _ = H.oracle_trim(batch, qcut=20, ncut=7)
    got, _, _ = gpu_trim(sp, batch, qcut=20, ncut=7)
    assert_same(got, want, batch)
    assert (want["flags"] & 0x30).any() and (want["flags"] & 0x0C).any()
    got2, _, _ = gpu_trim(sp, batch, force_bytewise=True, qcut=20, ncut=7)
    assert_same(got2, want, batch)


def test_high_overlap_halving_path(sp):
    """2x250 with inserts below the read length: overlaps of more than 170 compared bases take matchProbability's halving path."""
    batch = H.random_batch(2000, 250, 21, insert_mean=150, insert_sd=40, error_rate=0.02)
    want, _ = H.oracle_trim(batch)
    got, _, _ = gpu_trim(sp, batch)
    assert_same(got, want, batch)
    assert ((want["flags"] & 1) != 0).mean() > 0.9


def test_long_reads_up_to_maxlen(sp):
    """Reads of 400..999 bases run through the byte-wise path (MAXLEN-1 is the longest the reference accepts)."""
    batch = H.random_batch(300, 999, 31, insert_mean=700, insert_sd=300, ragged=True, stride=1000)
    want, _ = H.oracle_trim(batch)
    got, _, _ = gpu_trim(sp, batch)
    assert_same(got, want, batch)


def test_bytes_outside_acgtn(sp):
    """Read 1 may hold any byte (compared as a plain byte, AnalysisWorker.cpp:155-168); read 2 must be ACGTN or the reference throws
    (Sequence.cpp:46-71) -> status 1."""
    batch = H.random_batch(600, 150, 41, error_rate=0.01)
    rng = np.random.default_rng(5)
    for i in range(0, 600, 3):
        j = int(rng.integers(0, batch.len1[i]))
        batch.bases1[i, j] = rng.choice(np.frombuffer(b"acgtnXR.-", np.uint8))
    for i in range(1, 600, 50):
        j = int(rng.integers(0, batch.len2[i]))
        batch.bases2[i, j] = ord("a")
    batch.quals1[7, :20] = 200  # quality bytes >= 0x80 are negative chars in the reference (FastqFileStream.h:23-26)
    want, _ = H.oracle_trim(batch)
    assert (want["status"] == 1).sum() == 12
    got, _, _ = gpu_trim(sp, batch)
    assert_same(got, want, batch)


def test_custom_parameters(sp):
    """Non-default -match_perc / -mep / -qwin / -qoff / adapters with N."""
    batch = H.random_batch(2500, 120, 51, error_rate=0.05, lowq_tail=15.0, a1="AGATCGGAAGAGCNCACGTCTGAACTCC", a2="AGATCGGAAGAGCGTCGTNTAGGGAAAG")
    params = dict(a1="AGATCGGAAGAGCNCACGTCTGAACTCC", a2="AGATCGGAAGAGCGTCGTNTAGGGAAAG", match_perc=70.0, mep=1e-4, qcut=22, qwin=9, qoff=30, ncut=3)
    want, _ = H.oracle_trim(batch, **params)
    got, _, _ = gpu_trim(sp, batch, **params)
    assert_same(got, want, batch)


def test_empty_and_tiny_batches(sp):
    eng = sp.Engine(sp.TrimmingParameters(), devices=(0,), n_slots=1, max_pairs=16, max_len=150)
    eng.submit(0, 0)
    assert len(eng.wait(0)) == 0
    eng.close()
    batch = H.random_batch(1, 150, 61)
    want, _ = H.oracle_trim(batch)
    got, _, _ = gpu_trim(sp, batch)
    assert_same(got, want, batch)


def test_slot_protocol_errors(sp):
    eng = sp.Engine(sp.TrimmingParameters(), devices=(0,), n_slots=1, max_pairs=16, max_len=150)
    with pytest.raises(sp.SeqPurgeError, match="not submitted"):
        eng.wait(0)
    eng.submit(0, 4)
    with pytest.raises(sp.SeqPurgeError, match="not waited"):
        eng.submit(0, 4)
    eng.wait(0)
    with pytest.raises(sp.SeqPurgeError):
        eng.submit(0, 17)
    eng.close()
    with pytest.raises(sp.SeqPurgeError, match="15"):
        sp.Engine(sp.TrimmingParameters(a1="ACGT"), devices=(0,), n_slots=1, max_pairs=16, max_len=150)


SYNTH = {
    "C2_2x150": (dict(read_len=150, insert_mean=250, insert_sd=80, error_rate=0.001, lowq_tail_mean=3.0), dict()),
    "C3_2x250_high_overlap": (dict(read_len=250, insert_mean=150, insert_sd=40, insert_max=249, error_rate=0.001, lowq_tail_mean=3.0), dict()),
    "C4_2x150_errors_lowq": (dict(read_len=150, insert_mean=250, insert_sd=80, error_rate=0.02, lowq_tail_mean=20.0, n_run_rate=0.005), dict(qcut=15, ncut=7)),
    "C5_novaseq_like": (dict(read_len=150, insert_mean=350, insert_sd=100, error_rate=0.002, lowq_tail_mean=2.0, binned_quals=True), dict()),
}


@pytest.mark.parametrize("name", list(SYNTH), ids=list(SYNTH))
def test_device_synthetic_configs(sp, name):
    """BASELINE configs 2-5: batches generated on the device (first slice and a far slice of the stream), trimmed device-resident
    (spg_trim_device on torch tensors), copied back and compared with the oracle on the same bytes."""
    import torch

    cfg_kw, params = SYNTH[name]
    cfg = sp.SynthConfig(**cfg_kw)
    n = 20000
    stride = (cfg.read_len + 15) // 16 * 16
    dev = torch.device("cuda:0")
    eng = sp.Engine(sp.TrimmingParameters(**params), devices=(0,))
    for first in (0, 99_000_000):
        t = {k: torch.empty((n, stride), dtype=torch.uint8, device=dev) for k in ("bases1", "quals1", "bases2", "quals2")}
        l1 = torch.empty(n, dtype=torch.int16, device=dev)
        l2 = torch.empty(n, dtype=torch.int16, device=dev)
        res = torch.empty((n, 8), dtype=torch.uint8, device=dev)
        sp.synth_device(cfg, first, n, t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2)
        eng.set_option(sp.OPT_FULL_LEN, 0)  # the general kernel ...
        eng.trim_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, res)
        torch.cuda.synchronize()
        general = sp.results_from_tensor(res).copy()
        eng.set_option(sp.OPT_FULL_LEN, cfg.read_len)  # ... and the variant compiled for this read length
        eng.trim_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, res)
        torch.cuda.synchronize()
        got = sp.results_from_tensor(res)
        assert_same(general, got)
        batch = H.Batch(n, stride)
        for k in t:
            getattr(batch, k)[:n] = t[k].cpu().numpy()
        batch.len1[:n] = l1.cpu().numpy().view(np.uint16)
        batch.len2[:n] = l2.cpu().numpy().view(np.uint16)
        want, _ = H.oracle_trim(batch, threads=8, **params)
        assert_same(got, want, batch)
        assert (want["status"] == 0).all()
    eng.close()


def test_full_size_config2_properties(sp):
    """BASELINE config 2 at its full size (100 M synthetic 2x150 pairs, 10 resident batches of 10 M): properties that do not need the
    oracle on every pair -- the bit-plane path and the byte-wise path (two independent device implementations of the specification)
    agree on a checksum of checksums, the kernel variant compiled for 150-base reads and the general kernel agree on every record, every record satisfies the invariants of the trimming rules -- plus
    the oracle on random 50 k-pair slices of every batch."""
    import torch

    total = int(os.environ.get("SPG_FULL_SIZE_PAIRS", "100000000"))
    nb = 10
    n = total // nb // 8 * 8
    L, stride = 150, 150
    dev = torch.device("cuda:0")
    free, _ = torch.cuda.mem_get_info()
    if free < nb * n * (4 * stride + 4) + 4 * n * 8 + (2 << 30):
        pytest.skip("not enough free device memory for the full-size config")
    cfg = sp.SynthConfig(read_len=L)
    eng = sp.Engine(sp.TrimmingParameters(), devices=(0,))
    rng = np.random.default_rng(7)
    res = torch.empty((n, 8), dtype=torch.uint8, device=dev)
    res2 = torch.empty((n, 8), dtype=torch.uint8, device=dev)
    sums_planes, sums_bytes = [], []
    for b in range(nb):
        t = {k: torch.empty((n, stride), dtype=torch.uint8, device=dev) for k in ("bases1", "quals1", "bases2", "quals2")}
        l1 = torch.empty(n, dtype=torch.int16, device=dev)
        l2 = torch.empty(n, dtype=torch.int16, device=dev)
        sp.synth_device(cfg, b * n, n, t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2)
        eng.set_option(sp.OPT_FORCE_BYTEWISE, 0)
        eng.set_option(sp.OPT_FULL_LEN, L)  # kernel variant compiled for 150-base reads (what bench.py runs)
        eng.trim_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, res)
        eng.set_option(sp.OPT_FULL_LEN, 0)  # the general kernel
        eng.trim_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, res2)
        torch.cuda.synchronize()
        assert torch.equal(res, res2), "the full-length variant and the general kernel differ"
        w = res.view(torch.int64).view(-1)
        sums_planes.append(int((w * (torch.arange(n, device=dev) % 1000003 + 1)).sum().item()))
        # invariants of the trimming rules on every record
        len1 = (w & 0xFFFF)
        len2 = ((w >> 16) & 0xFFFF)
        off = ((w >> 32) & 0xFFFF)
        flags = ((w >> 48) & 0xFF)
        status = ((w >> 56) & 0xFF)
        assert int(status.max().item()) == 0
        assert int(len1.max().item()) <= L and int(len2.max().item()) <= L
        ins = (flags & 1) != 0
        assert bool(((off == 0xFFFF) == ~ins).all()), "best_offset set iff insert flag"
        assert bool((len1[ins] <= L - off[ins]).all()) and bool((len2[ins] <= L - off[ins]).all())
        assert bool(((flags & 3) != 3).all()), "insert and adapter-only trimming are exclusive"
        frac = float(ins.float().mean().item())
        assert 0.05 < frac < 0.2  # inserts ~ N(250,80): about 10 % are shorter than the read
        # the byte-wise device path on the same batch
        eng.set_option(sp.OPT_FORCE_BYTEWISE, 1)
        eng.trim_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, res2)
        torch.cuda.synchronize()
        w2 = res2.view(torch.int64).view(-1)
        sums_bytes.append(int((w2 * (torch.arange(n, device=dev) % 1000003 + 1)).sum().item()))
        # the oracle on a random slice
        m = 50_000
        st = int(rng.integers(0, n - m)) // 8 * 8
        batch = H.Batch(m, stride)
        for k in t:
            getattr(batch, k)[:m] = t[k][st : st + m].cpu().numpy()
        batch.len1[:m] = l1[st : st + m].cpu().numpy().view(np.uint16)
        batch.len2[:m] = l2[st : st + m].cpu().numpy().view(np.uint16)
        want, _ = H.oracle_trim(batch, threads=8)
        assert_same(sp.results_from_tensor(res[st : st + m]), want, batch)
        del t, l1, l2
    eng.close()
    assert sums_planes == sums_bytes, "plane path and byte-wise path disagree"


@pytest.mark.parametrize("L", [31, 32, 33, 159, 160, 161, 255, 256, 257, 319, 320, 321, 400])
def test_plane_width_boundaries(sp, L):
    """Read lengths around the word / kernel-variant boundaries (NW = 5, 8, 10 plane words, byte-wise beyond 320), ragged lengths."""
    n = 600 if L <= 320 else 200
    batch = H.random_batch(n, L, 700 + L, error_rate=0.02, n_rate=0.003, lowq_tail=6.0, ragged=(L % 2 == 0), stride=(L + 1) // 2 * 2 if L >= 16 else 16)
    want, _ = H.oracle_trim(batch)
    got, _, _ = gpu_trim(sp, batch)
    assert_same(got, want, batch)


@pytest.mark.parametrize("params", [
    dict(qwin=40, qcut=20), dict(qwin=1, qcut=30), dict(qwin=16, qcut=10), dict(qwin=17, qcut=25),
    dict(match_perc=50.0, mep=1e-3), dict(match_perc=100.0), dict(mep=1e-12), dict(ncut=1), dict(qcut=0, ncut=0), dict(qoff=64, qcut=5),
    dict(mep=1.0, match_perc=0.0), dict(mep=2.0, match_perc=30.0),
], ids=lambda p: ",".join(f"{k}={v}" for k, v in p.items()))
def test_parameter_corners(sp, params):
    """Window sizes on both sides of the fast paths (<= 8 paired, <= 32 scan, > 32 general), permissive and strict match filters
    (many / no survivors of the pre-filter), single-N trimming, everything off."""
    batch = H.random_batch(1500, 150, 900, error_rate=0.04, n_rate=0.004, lowq_tail=25.0, n_runs=0.05)
    want, _ = H.oracle_trim(batch, **params)
    got, _, _ = gpu_trim(sp, batch, **params)
    assert_same(got, want, batch)
    got2, _, _ = gpu_trim(sp, batch, force_bytewise=True, **params)
    assert_same(got2, want, batch)


FULL_LENGTHS = [150, 151, 100, 101, 125, 126, 75, 76, 250, 251, 200, 201, 300, 301]


def _full_batch(L, seed, n=1536):
    """Mostly full-length pairs (the fast path of the kernel variant compiled for L), with ragged pairs, N bases and bytes outside
    ACGTN mixed in (they must leave the fast path) and inserts from far shorter to far longer than the read."""
    b = H.random_batch(n, L, seed=seed, insert_mean=0.9 * L, insert_sd=0.6 * L, error_rate=0.02, n_rate=0.0003, lowq_tail=6.0, n_runs=0.01)
    r = H.random_batch(n // 8, L, seed=seed + 1, ragged=True, stride=b.stride)
    idx = np.random.default_rng(seed).choice(n, n // 8, replace=False)
    for j, i in enumerate(idx):
        for k in ("bases1", "quals1", "bases2", "quals2", "len1", "len2"):
            getattr(b, k)[i] = getattr(r, k)[j]
    b.bases1[5, 7] = ord("X")  # read 1: compared as a plain byte
    b.bases2[9, 3] = ord("R")  # read 2: the reference throws
    return b


@pytest.mark.parametrize("L", FULL_LENGTHS)
def test_full_length_variant(sp, L):
    """The kernel variants compiled for one read length (pairs of two full-length reads take a path with compile-time masks) against
    the oracle and against the general kernel."""
    batch = _full_batch(L, seed=4000 + L)
    assert int(((batch.len1[: batch.n] == L) & (batch.len2[: batch.n] == L)).sum()) > batch.n // 2
    want, _ = H.oracle_trim(batch)
    got, _, _ = gpu_trim(sp, batch, full_len=L, expect_kernel=f"trim_lanes_kernel<NW={5 if L <= 160 else 8 if L <= 256 else 10},FULL={L},")
    assert_same(got, want, batch)
    warp, _, _ = gpu_trim(sp, batch, full_len=L, kernel=sp.KERNEL_WARP_PER_PAIR, expect_kernel=f"trim_kernel<NW={5 if L <= 160 else 8 if L <= 256 else 10},CW=8,MINB=3,FULL={L}>")
    assert_same(warp, want, batch)
    general, _, _ = gpu_trim(sp, batch, full_len=0, expect_kernel="FULL=0>")
    assert_same(general, want, batch)


@pytest.mark.parametrize("params", [
    dict(a1="CTGTCTCTTATACACATCT", a2="CTGTCTCTTATACACATCT"),  # a_size 19
    dict(a1="AGATCGGAAGAGCNCACGTCTGAAC", a2="AGATCGGAAGAGCGTCNTGTAGGGA"),  # N inside the adapters
    dict(match_perc=70.0, mep=1e-4, qcut=20, qwin=7, ncut=3),
    dict(adapter_overlap=6, qcut=0, ncut=0),
], ids=["short_adapters", "adapter_n", "loose", "overlap6"])
def test_full_length_variant_parameters(sp, params):
    for L in (150, 251):
        kw = {k: params[k] for k in ("a1", "a2") if k in params}
        batch = H.random_batch(1024, L, seed=77 + L, insert_mean=0.9 * L, insert_sd=0.6 * L, error_rate=0.03, n_rate=0.0002, **kw)
        want, _ = H.oracle_trim(batch, **params)
        got, _, _ = gpu_trim(sp, batch, full_len=L, **params)
        assert_same(got, want, batch)
        warp, _, _ = gpu_trim(sp, batch, full_len=L, kernel=sp.KERNEL_WARP_PER_PAIR, **params)
        assert_same(warp, want, batch)


LANE_CASES = {
    # (read length, row stride, batch keywords, trimming parameters)
    "plain_150_tight": (150, 150, dict(insert_mean=160, insert_sd=90, error_rate=0.01, n_rate=0.0, lowq_tail=4.0), dict()),
    "plain_150_wide_rows": (150, 160, dict(insert_mean=160, insert_sd=90, error_rate=0.03, n_rate=0.0, lowq_tail=4.0), dict()),
    "long_lowq_tails": (150, 150, dict(insert_mean=200, insert_sd=90, error_rate=0.02, n_rate=0.0, lowq_tail=40.0), dict(qcut=20)),
    "windows_1_to_8": (126, 126, dict(insert_mean=100, insert_sd=60, error_rate=0.02, n_rate=0.0, lowq_tail=12.0), dict(qwin=8, qcut=25)),
    "window_1": (101, 102, dict(insert_mean=100, insert_sd=60, error_rate=0.02, n_rate=0.0, lowq_tail=12.0), dict(qwin=1, qcut=30)),
    "no_quality_trimming": (75, 76, dict(insert_mean=60, insert_sd=40, error_rate=0.02, n_rate=0.0005), dict(qcut=0, ncut=0)),
    "all_overlap_250": (250, 250, dict(insert_mean=150, insert_sd=60, error_rate=0.02, n_rate=0.0), dict()),
    "short_inserts_300": (300, 300, dict(insert_mean=40, insert_sd=60, error_rate=0.01, n_rate=0.0001), dict()),
    "loose_filter": (150, 150, dict(insert_mean=160, insert_sd=90, error_rate=0.08, n_rate=0.0), dict(match_perc=60.0, mep=1e-3)),
    "adapter_only_hits": (150, 150, dict(insert_mean=140, insert_sd=25, error_rate=0.12, n_rate=0.0, lowq_tail=2.0), dict()),
    "short_adapters_19": (100, 100, dict(insert_mean=90, insert_sd=25, error_rate=0.1, n_rate=0.0, a1="CTGTCTCTTATACACATCT", a2="CTGTCTCTTATACACATCT"),
                          dict(a1="CTGTCTCTTATACACATCT", a2="CTGTCTCTTATACACATCT")),
    "many_n_runs": (150, 150, dict(insert_mean=140, insert_sd=60, error_rate=0.02, n_rate=0.004, n_runs=0.3, lowq_tail=5.0), dict(ncut=7)),
    "single_n_cut": (126, 126, dict(insert_mean=120, insert_sd=60, error_rate=0.02, n_rate=0.003, n_runs=0.1), dict(ncut=1)),
    "n_250_ncut3": (250, 250, dict(insert_mean=170, insert_sd=60, error_rate=0.03, n_rate=0.002, n_runs=0.2, lowq_tail=10.0), dict(ncut=3, qcut=20)),
    "n_300_no_ncut": (300, 300, dict(insert_mean=250, insert_sd=120, error_rate=0.02, n_rate=0.002, n_runs=0.1), dict(ncut=0)),
    "n_and_ragged_mix": (151, 152, dict(insert_mean=160, insert_sd=90, error_rate=0.02, n_rate=0.002, n_runs=0.05, lowq_tail=10.0), dict()),
}


@pytest.mark.parametrize("name", list(LANE_CASES), ids=list(LANE_CASES))
def test_lane_per_pair_kernel(sp, name):
    """The lane-per-pair layout of the read-length variants (spg_lanes.cuh) against the oracle and against the warp-per-pair kernel:
    tight and wide row strides (rows that start in the middle of a word), pair counts that leave the last 32-pair tile ragged,
    trimming points far from the 3' end (several blocks of the per-lane quality search), all window sizes the layout serves."""
    L, stride, kw, params = LANE_CASES[name]
    n = 32 * 37 + 13
    batch = H.random_batch(n, L, seed=5000 + L + len(name), stride=stride, **kw)
    if name == "n_and_ragged_mix":
        r = H.random_batch(200, L, seed=9, ragged=True, stride=stride)
        for j in range(200):
            for k in ("bases1", "quals1", "bases2", "quals2", "len1", "len2"):
                getattr(batch, k)[5 * j + 1] = getattr(r, k)[j]
    want, _ = H.oracle_trim(batch, **params)
    got, _, _ = gpu_trim(sp, batch, full_len=L, kernel=sp.KERNEL_LANE_PER_PAIR, expect_kernel="trim_lanes_kernel", **params)
    assert_same(got, want, batch)
    every, _, _ = gpu_trim(sp, batch, full_len=L, kernel=sp.KERNEL_LANE_PER_PAIR, seed_scan=0, **params)  # adapter scans without the filter
    assert_same(every, want, batch)
    warp, _, _ = gpu_trim(sp, batch, full_len=L, kernel=sp.KERNEL_WARP_PER_PAIR, expect_kernel="trim_kernel", **params)
    assert_same(warp, want, batch)
    tails, _, _ = gpu_trim(sp, batch, full_len=L, kernel=sp.KERNEL_LANE_PER_PAIR, qual_tails=True, **params)  # quality tails shipped with the bases
    assert_same(tails, want, batch)


def test_quality_tails_in_the_slot(sp):
    """SPG_OPT_QUAL_TAILS: the last 16 qualities of every read travel with the bases, the quality rows stay in the pinned slot. Reads whose
    trimming point lies within the tail are decided from it, all others (cut by the adapter steps, trimmed deeper, bytes >= 0x80 in the
    tail, shorter than the row) from the row -- the records must not depend on the option. Slots filled to less than half take the
    separate copies of spg_submit, the last tile is ragged, rows start in the middle of a word."""
    L = 150
    batch = H.random_batch(32 * 20 + 5, L, seed=77, stride=L, insert_mean=170, insert_sd=80, error_rate=0.02, n_rate=0.001, lowq_tail=9.0)
    rng = np.random.default_rng(3)
    for i in rng.choice(batch.n, 60, replace=False):  # qualities the byte arithmetic of the fast search does not cover
        batch.quals1[i, L - 1 - int(rng.integers(0, 16))] = 0x80 + int(rng.integers(0, 100))
    for i in rng.choice(batch.n, 60, replace=False):  # tails that fail completely: the search goes on in the row
        batch.quals2[i, L - 30 :] = 33 + 2
    r = H.random_batch(100, L, seed=10, ragged=True, stride=L)
    for j in range(100):
        for k in ("bases1", "quals1", "bases2", "quals2", "len1", "len2"):
            getattr(batch, k)[6 * j + 2] = getattr(r, k)[j]
    for params in (dict(), dict(qcut=25), dict(qcut=2), dict(qwin=7, qcut=20), dict(qcut=0)):
        want, _ = H.oracle_trim(batch, **params)
        for chunk, n_slots in ((None, 1), (300, 2)):
            got, _, _ = gpu_trim(sp, batch, full_len=L, qual_tails=True, chunk=chunk, n_slots=n_slots, **params)
            assert_same(got, want, batch)
    # garbage in the tails of reads that are NOT decided from them must not matter: pairs with an insert hit are cut before quality trimming
    want, _ = H.oracle_trim(batch)
    p = sp.TrimmingParameters()
    eng = sp.Engine(p, devices=(0,), n_slots=1, max_pairs=batch.n, max_len=L)
    eng.set_option(sp.OPT_QUAL_TAILS, 1)
    s = eng.slot(0)
    for name in ("bases1", "quals1", "bases2", "quals2"):
        getattr(s, name)[: batch.n] = getattr(batch, name)[: batch.n]
    s.len1[: batch.n] = batch.len1[: batch.n]
    s.len2[: batch.n] = batch.len2[: batch.n]
    s.fill_qtails(batch.n)
    hit = (want["flags"] & 1) != 0
    assert hit.sum() > 50
    s.qtail1[: batch.n][hit] = 33 + 40
    s.qtail2[: batch.n][hit] = 33
    eng.submit(0, batch.n)
    assert_same(eng.wait(0).copy(), want, batch)
    eng.close()


def test_lane_per_pair_low_complexity_reads(sp):
    """Homopolymer and short-period reads: every insert offset survives the pre-filter, far more than a lane can queue -- such pairs
    must take the general path and still agree with the oracle."""
    L = 150
    batch = H.random_batch(32 * 6, L, seed=12, stride=L, n_rate=0.0)
    for i in range(0, batch.n, 3):
        batch.bases1[i, :L] = ord("A")
        batch.bases2[i, :L] = ord("T")
    for i in range(1, batch.n, 3):
        batch.bases1[i, :L] = np.frombuffer((b"AC" * L)[:L], np.uint8)
        batch.bases2[i, :L] = np.frombuffer((b"GT" * L)[:L], np.uint8)
    want, _ = H.oracle_trim(batch)
    got, _, _ = gpu_trim(sp, batch, full_len=L, kernel=sp.KERNEL_LANE_PER_PAIR, expect_kernel="trim_lanes_kernel")
    assert_same(got, want, batch)


@pytest.mark.parametrize("stages,ctas", [(2, 1), (3, 2), (4, 0)])
def test_lane_per_pair_ring_many_rounds(sp, stages, ctas):
    """Many tiles per CTA (the ring wraps dozens of times, warps claim tiles further ahead than the ring is deep): the lane-per-pair
    kernel against the warp-per-pair kernel on every record and against the oracle on a slice."""
    import torch

    dev = torch.device("cuda:0")
    n, L = 400_000, 150
    cfg = sp.SynthConfig(read_len=L, insert_mean=200, insert_sd=80, error_rate=0.005, lowq_tail_mean=6.0)
    t = {k: torch.empty((n, L), dtype=torch.uint8, device=dev) for k in ("bases1", "quals1", "bases2", "quals2")}
    l1 = torch.empty(n, dtype=torch.int16, device=dev)
    l2 = torch.empty(n, dtype=torch.int16, device=dev)
    sp.synth_device(cfg, 7_000_000, n, t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2)
    eng = sp.Engine(sp.TrimmingParameters(), devices=(0,))
    eng.set_option(sp.OPT_FULL_LEN, L)
    res = {}
    for layout in (sp.KERNEL_WARP_PER_PAIR, sp.KERNEL_LANE_PER_PAIR):
        eng.set_option(sp.OPT_KERNEL, layout)
        if layout == sp.KERNEL_LANE_PER_PAIR:
            eng.set_option(sp.OPT_STAGES, stages)
            eng.set_option(sp.OPT_GRID_CTAS_PER_SM, ctas)
        r = torch.empty((n, 8), dtype=torch.uint8, device=dev)
        for _ in range(2):
            eng.trim_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, r)
        torch.cuda.synchronize()
        res[layout] = r
    assert "trim_lanes_kernel" in eng.last_kernel
    assert torch.equal(res[sp.KERNEL_WARP_PER_PAIR], res[sp.KERNEL_LANE_PER_PAIR])
    m = 40_000
    batch = H.Batch(m, L)
    for k in t:
        getattr(batch, k)[:m] = t[k][n - m :].cpu().numpy()
    batch.len1[:m] = l1[n - m :].cpu().numpy().view(np.uint16)
    batch.len2[:m] = l2[n - m :].cpu().numpy().view(np.uint16)
    want, _ = H.oracle_trim(batch, threads=8)
    assert_same(sp.results_from_tensor(res[sp.KERNEL_LANE_PER_PAIR][n - m :]).copy(), want, batch)
    eng.close()


def test_engine_reuses_kernels_across_row_strides(sp):
    """One context, device-resident launches of the same kernel instantiation with growing row strides (a stream reopened for longer
    reads): the shared-memory attribute and the cached occupancy must follow the launch geometry."""
    import torch

    dev = torch.device("cuda:0")
    eng = sp.Engine(sp.TrimmingParameters(), devices=(0,))
    for layout in (sp.KERNEL_WARP_PER_PAIR, sp.KERNEL_LANE_PER_PAIR):
        eng.set_option(sp.OPT_KERNEL, layout)
        for L, stride in ((180, 192), (220, 224), (250, 256), (150, 150), (150, 160)):
            batch = H.random_batch(500, L, seed=L + stride, stride=stride)
            want, _ = H.oracle_trim(batch)
            t = {k: torch.from_numpy(getattr(batch, k)).to(dev) for k in ("bases1", "quals1", "bases2", "quals2")}
            l1 = torch.from_numpy(batch.len1.view(np.int16)).to(dev)
            l2 = torch.from_numpy(batch.len2.view(np.int16)).to(dev)
            res = torch.empty((batch.n, 8), dtype=torch.uint8, device=dev)
            eng.set_option(sp.OPT_FULL_LEN, L)
            eng.trim_device(t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2, res, n_pairs=batch.n)
            torch.cuda.synchronize()
            assert_same(sp.results_from_tensor(res).copy(), want, batch)
    eng.close()


def test_full_length_variant_error_correction(sp):
    params = dict(ec=True)
    batch = H.random_batch(1024, 150, seed=991, insert_mean=120, insert_sd=40, error_rate=0.03, n_rate=0.0)
    ref = batch.copy()
    want, want_ec = H.oracle_trim(ref, **params)
    got, edited, got_ec = gpu_trim(sp, batch, full_len=150, **params)
    assert_same(got, want, batch)
    for name in ("bases1", "quals1", "bases2", "quals2"):
        assert np.array_equal(getattr(edited, name)[: batch.n], getattr(ref, name)[: batch.n]), name
    for k in want_ec:
        assert np.array_equal(got_ec[k], want_ec[k]), k
