"""GPU parity of the FASTQ stream: text in -> framing, trimming, routing, record layout, consensus counters
on the device (spg_fq_* of the C ABI) -> text out, against
  * the reference's golden output files (decompressed content, as the reference's COMPARE_FILES does), and
  * the oracle CLI, whose reader/writer restate FastqFileStream::readEntry / OutputWorker::run / FastqOutfileStream::write,
on the reference fixtures, on awkward texts (CRLF, no final newline, blank lines, truncated last record) and with chunk
boundaries in every position (small text buffers, few pairs per chunk, carry-over by the caller).
"""
import gzip
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from test_gpu_parity import sp  # noqa: F401
from test_oracle_golden import CASES, COMMON

pytestmark = pytest.mark.gpu
G = H.GOLDEN

FLAG_TO_PARAM = {"-a1": ("a1", str), "-a2": ("a2", str), "-qcut": ("qcut", int), "-ncut": ("ncut", int), "-qwin": ("qwin", int), "-qoff": ("qoff", int),
                 "-match_perc": ("match_perc", float), "-mep": ("mep", float)}


def parse_flags(flags):
    """SeqPurge command line flags of a test case -> (engine parameters, min_len, singles)."""
    params, min_len, singles = {}, 30, False
    it = iter(flags)
    for f in it:
        if f in FLAG_TO_PARAM:
            k, conv = FLAG_TO_PARAM[f]
            params[k] = conv(next(it))
        elif f == "-min_len":
            min_len = int(next(it))
        elif f == "-out3":
            next(it)
            singles = True
        elif f == "-ec":
            params["ec"] = True
        else:
            raise AssertionError(f)
    return params, min_len, singles


def gz(path):
    with gzip.open(path, "rb") as f:
        return f.read()


def run_stream(sp, text1, text2, params, min_len=30, singles=False, text_cap=8 << 20, max_pairs=8192, max_len=160, n_slots=2, feed=None):  # noqa: F811
    """Drives spg_fq_* like the command line does: fills the chunks from the two texts (`feed` bytes at a time at most), carries the
    unconsumed tail over, keeps n_slots chunks in flight and retires them in order. Returns (outs[4], chunks)."""
    eng = sp.Engine(sp.TrimmingParameters(**params), devices=(0,))
    fq = sp.FastqStream(eng, n_slots=n_slots, max_pairs=max_pairs, max_len=max_len, text_cap=text_cap, min_len=min_len, singles=singles)
    texts, pos, carry = [text1, text2], [0, 0], [b"", b""]
    outs, chunks = [[], [], [], []], []
    feed = feed or text_cap
    # one chunk at a time in flight per dependency: the carry of chunk k is needed to build chunk k+1, so the stream is driven
    # synchronously here (the command line overlaps the inflate of the next bytes instead)
    slot = 0
    while True:
        chunk, final = [], []
        for f in range(2):
            room = text_cap - len(carry[f])
            take = min(room, feed, len(texts[f]) - pos[f])
            chunk.append(carry[f] + texts[f][pos[f] : pos[f] + take])
            pos[f] += take
            final.append(pos[f] == len(texts[f]))
        fq.submit(slot, chunk[0], chunk[1], final[0], final[1])
        c = fq.wait(slot)
        slot = (slot + 1) % n_slots
        chunks.append(c)
        for k in range(4):
            outs[k].append(c.out[k])
        carry = [chunk[0][c.consumed[0] :], chunk[1][c.consumed[1] :]]
        if c.error_pair >= 0:
            break
        if all(final) and (not carry[0] or not carry[1]):
            break
        assert c.n_pairs > 0 or any(len(chunk[f]) < text_cap and not final[f] for f in range(2)), "no progress: a record does not fit the text buffer"
    counts, unknown = fq.consensus()
    fq.close()
    eng.close()
    return [b"".join(o) for o in outs], chunks, carry, (counts, unknown)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reference_goldens_one_chunk(sp, case):  # noqa: F811
    name, i1, i2, o1, o2, flags = case
    params, min_len, singles = parse_flags(flags)
    t1, t2 = gz(f"{G}/SeqPurge_in{i1}.fastq.gz"), gz(f"{G}/SeqPurge_in{i2}.fastq.gz")
    longest = max(len(x) for x in t1.split(b"\n")[1::4] + t2.split(b"\n")[1::4])
    outs, chunks, carry, _ = run_stream(sp, t1, t2, params, min_len, singles, max_pairs=65536, max_len=max(longest, 100))
    assert chunks[0].max_len == longest
    assert len(chunks) == 1 and carry == [b"", b""] and chunks[0].error_pair == -1
    assert outs[0] == gz(f"{G}/SeqPurge_out{o1}.fastq.gz")
    assert outs[1] == gz(f"{G}/SeqPurge_out{o2}.fastq.gz")
    if name == "test_07":
        assert outs[2] == gz(f"{G}/SeqPurge_out15_R1.fastq.gz")
        assert outs[3] == gz(f"{G}/SeqPurge_out15_R2.fastq.gz")
    else:
        assert outs[2] == b"" and outs[3] == b""


@pytest.mark.parametrize("text_cap,max_pairs,feed", [(4096, 8192, None), (1 << 20, 7, None), (2048, 3, 700), (1 << 16, 100, 5000)])
@pytest.mark.parametrize("case", [CASES[0], CASES[6], CASES[9]], ids=["test_01", "test_07", "test_10_ec"])
def test_chunk_boundaries_everywhere(sp, case, text_cap, max_pairs, feed):  # noqa: F811
    name, i1, i2, o1, o2, flags = case
    params, min_len, singles = parse_flags(flags)
    outs, chunks, carry, _ = run_stream(sp, gz(f"{G}/SeqPurge_in{i1}.fastq.gz"), gz(f"{G}/SeqPurge_in{i2}.fastq.gz"), params, min_len, singles,
                                        text_cap=text_cap, max_pairs=max_pairs, feed=feed, n_slots=3)
    assert len(chunks) > 3 and carry == [b"", b""]
    assert outs[0] == gz(f"{G}/SeqPurge_out{o1}.fastq.gz")
    assert outs[1] == gz(f"{G}/SeqPurge_out{o2}.fastq.gz")
    if name == "test_07":
        assert outs[2] == gz(f"{G}/SeqPurge_out15_R1.fastq.gz")
        assert outs[3] == gz(f"{G}/SeqPurge_out15_R2.fastq.gz")


def oracle_cli(oracle_build, tmp_path, text1, text2, flags):
    """The oracle's command line on two plain-text FASTQ files -> decompressed outputs (out1, out2, out3_R1, out3_R2), summary text."""
    (tmp_path / "a_R1.fastq").write_bytes(text1)
    (tmp_path / "a_R2.fastq").write_bytes(text2)
    cmd = [os.path.join(oracle_build, "seqpurge_oracle"), "-in1", str(tmp_path / "a_R1.fastq"), "-in2", str(tmp_path / "a_R2.fastq"), "-out1", str(tmp_path / "o1.gz"),
           "-out2", str(tmp_path / "o2.gz"), "-out3", str(tmp_path / "o3"), "-summary", str(tmp_path / "summary.txt")] + flags
    subprocess.run(cmd, check=True)
    return [gz(tmp_path / "o1.gz"), gz(tmp_path / "o2.gz"), gz(tmp_path / "o3_R1.fastq.gz"), gz(tmp_path / "o3_R2.fastq.gz")], (tmp_path / "summary.txt").read_text()


def awkward_texts():
    t1 = gz(f"{G}/SeqPurge_in1.fastq.gz")[:60000]
    t2 = gz(f"{G}/SeqPurge_in2.fastq.gz")
    n = t1.count(b"\n") // 4
    t1 = b"\n".join(t1.split(b"\n")[: 4 * n]) + b"\n"
    t2 = b"\n".join(t2.split(b"\n")[: 4 * n]) + b"\n"
    return {
        "plain": (t1, t2),
        "crlf": (t1.replace(b"\n", b"\r\n"), t2.replace(b"\n", b"\r\r\n")),
        "no_final_newline": (t1[:-1], t2[:-1]),
        "no_final_newline_crlf": (t1.replace(b"\n", b"\r\n")[:-2], t2[:-1] + b"\r"),
        "blank_lines_at_the_end": (t1 + b"\n", t2 + b"\n"),
        "truncated_last_record": (t1[: t1.rstrip(b"\n").rfind(b"\n") + 1], t2[: t2.rstrip(b"\n").rfind(b"\n") + 1]),  # quality line of the last record missing
        "empty": (b"", b""),
    }


@pytest.mark.parametrize("kind", list(awkward_texts()))
@pytest.mark.parametrize("text_cap,max_pairs", [(8 << 20, 8192), (3000, 5)])
def test_awkward_texts_match_the_oracle_reader(sp, oracle_build, tmp_path, kind, text_cap, max_pairs):  # noqa: F811
    t1, t2 = awkward_texts()[kind]
    flags = ["-min_len", "15", "-qcut", "20"]
    params, min_len, _ = parse_flags(flags)
    if kind == "truncated_last_record":
        # |bases| != |qualities| in the last record: the reference's stream does not validate, the host layer here raises
        outs, chunks, carry, _ = run_stream(sp, t1, t2, params, min_len, True, text_cap=text_cap, max_pairs=max_pairs)
        last = chunks[-1]
        assert last.error_pair == last.n_pairs - 1 and last.frame_status[last.error_pair] == sp.FQ_LENGTH_MISMATCH
        return
    want, _ = oracle_cli(oracle_build, tmp_path, t1, t2, flags)
    outs, chunks, carry, _ = run_stream(sp, t1, t2, params, min_len, True, text_cap=text_cap, max_pairs=max_pairs)
    assert all(c.error_pair == -1 for c in chunks) and carry == [b"", b""]
    for k in range(4):
        assert outs[k] == want[k], (kind, k)


def test_header_mismatch_and_unequal_record_counts(sp):  # noqa: F811
    t1, t2 = awkward_texts()["plain"]
    lines = t2.split(b"\n")
    lines[4 * 17] = lines[4 * 17].replace(b":", b";", 1)
    outs, chunks, _, _ = run_stream(sp, t1, b"\n".join(lines), dict(), 30, False)
    assert chunks[0].error_pair == 17 and chunks[0].frame_status[17] == sp.FQ_HEADER_MISMATCH and (chunks[0].frame_status[:17] == 0).all()
    # "/1" and "/2" suffixes are tolerated, other suffixes are not (AnalysisWorker.cpp:113-117)
    a = b"@r/1 x\nACGT\n+\nIIII\n" * 3
    b = b"@r/2 y\nACGT\n+\nIIII\n" * 3
    _, chunks, _, _ = run_stream(sp, a, b, dict(), 0, False)
    assert chunks[0].error_pair == -1 and chunks[0].n_pairs == 3
    _, chunks, _, _ = run_stream(sp, a, b.replace(b"/2", b"/3"), dict(), 0, False)
    assert chunks[0].error_pair == 0
    # one file has more entries: all pairs of the shorter file are processed, the rest stays unconsumed for the caller to report
    _, chunks, carry, _ = run_stream(sp, t1, t2 + b"@extra\nACGT\n+\nIIII\n", dict(), 30, False)
    assert chunks[-1].records[1] == chunks[-1].records[0] + 1 and carry[0] == b"" and carry[1] == b"@extra\nACGT\n+\nIIII\n"


def test_read_longer_than_the_rows_is_reported(sp):  # noqa: F811
    r = b"@x\n" + b"ACGT" * 60 + b"\n+\n" + b"I" * 240 + b"\n"
    _, chunks, _, _ = run_stream(sp, r * 4, r * 4, dict(), 30, False, max_len=160)
    assert chunks[0].error_pair == 0 and chunks[0].frame_status[0] == sp.FQ_TOO_LONG and chunks[0].max_len == 240
    outs, chunks, _, _ = run_stream(sp, r * 4, r * 4, dict(), 30, False, max_len=240)
    assert chunks[0].error_pair == -1 and chunks[0].n_pairs == 4


def test_consensus_counters_and_records(sp):  # noqa: F811
    """Adapter consensus (AnalysisWorker.cpp:279-290) against a direct restatement over the oracle's result records."""
    t1, t2 = gz(f"{G}/SeqPurge_in1.fastq.gz"), gz(f"{G}/SeqPurge_in2.fastq.gz")
    params = dict(ncut=0, qcut=0)
    outs, chunks, _, (counts, unknown) = run_stream(sp, t1, t2, params, 15, False, max_pairs=300)
    batch = H.golden_batch(1, 2)
    want, _ = H.oracle_trim(batch, **params)
    got = np.concatenate([c.results for c in chunks])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    assert np.array_equal(np.concatenate([c.len1 for c in chunks]), batch.len1[: batch.n])
    exp = np.zeros((2, 40, 5), np.int64)
    idx = {65: 0, 67: 1, 71: 2, 84: 3, 78: 4}
    for p in np.nonzero(want["flags"] & 1)[0]:
        o1, o2, off = int(batch.len1[p]), int(batch.len2[p]), int(want["best_offset"][p])
        nl = o2 - off
        for i in range(40):
            if nl + i < o1:
                exp[0, i, idx[int(batch.bases1[p, nl + i])]] += 1
            if i < off:
                exp[1, i, idx[int(batch.bases2[p, nl + i])]] += 1
    assert not unknown and np.array_equal(counts, exp)


@pytest.mark.parametrize("singles,min_len", [(False, 30), (True, 60), (True, 0)])
def test_summary_counters_on_the_device(sp, singles, min_len):  # noqa: F811
    """spg_fq_stats (the counters OutputWorker::run adds to TrimmingStatistics, OutputWorker.cpp:59-77) of every chunk against a
    direct restatement over the chunk's result records and untrimmed lengths."""
    t1, t2 = gz(f"{G}/SeqPurge_in1.fastq.gz"), gz(f"{G}/SeqPurge_in2.fastq.gz")
    params = dict(qcut=20, ncut=3)
    outs, chunks, _, _ = run_stream(sp, t1, t2, params, min_len, singles, max_pairs=700)
    assert sum(c.n_pairs for c in chunks) == 2502
    for c in chunks:
        st, r = c.stats, c.results
        assert st is not None
        assert st["reads_trimmed_insert"] == 2 * int(((r["flags"] & 1) != 0).sum())
        assert st["reads_trimmed_adapter"] == 2 * int(((r["flags"] & 2) != 0).sum())
        assert st["reads_trimmed_q"] == int(((r["flags"] & 4) != 0).sum() + ((r["flags"] & 8) != 0).sum())
        assert st["reads_trimmed_n"] == int(((r["flags"] & 16) != 0).sum() + ((r["flags"] & 32) != 0).sum())
        ok1, ok2 = r["len1"] >= min_len, r["len2"] >= min_len
        removed = np.where(ok1 & ok2, 0, np.where((ok1 | ok2) & singles, 1, 2))
        assert st["reads_removed"] == int(removed.sum())
        rem = np.bincount(np.concatenate([r["len1"], r["len2"]]).astype(np.int64), minlength=1000)
        assert np.array_equal(st["bases_remaining"], rem)
        trimmed = np.zeros(1000, np.int64)
        np.add.at(trimmed, c.len1.astype(np.int64), c.len1.astype(np.int64) - r["len1"])
        np.add.at(trimmed, c.len2.astype(np.int64), c.len2.astype(np.int64) - r["len2"])
        assert np.array_equal(st["trimmed_bases_by_length"], trimmed)


def test_synthetic_20k_pairs_against_the_oracle_cli(sp, oracle_build, tmp_path):  # noqa: F811
    import torch

    cfg = sp.SynthConfig(read_len=150, insert_mean=250, insert_sd=80, error_rate=0.01, lowq_tail_mean=6.0, n_run_rate=0.002)
    n, stride = 20000, 150
    dev = torch.device("cuda:0")
    t = {k: torch.empty((n, stride), dtype=torch.uint8, device=dev) for k in ("bases1", "quals1", "bases2", "quals2")}
    l1 = torch.empty(n, dtype=torch.int16, device=dev)
    l2 = torch.empty(n, dtype=torch.int16, device=dev)
    sp.synth_device(cfg, 12345, n, t["bases1"], t["quals1"], t["bases2"], t["quals2"], l1, l2)
    torch.cuda.synchronize()
    rows = {k: v.cpu().numpy() for k, v in t.items()}
    L1, L2 = l1.cpu().numpy(), l2.cpu().numpy()
    recs = [[], []]
    for i in range(n):
        h = b"@SIM:1:B200:1:%d:%d" % (i // 1000, i)
        recs[0].append(h + b" 1:N:0:ACGT\n" + rows["bases1"][i, : L1[i]].tobytes() + b"\n+\n" + rows["quals1"][i, : L1[i]].tobytes() + b"\n")
        recs[1].append(h + b" 2:N:0:ACGT\n" + rows["bases2"][i, : L2[i]].tobytes() + b"\n+\n" + rows["quals2"][i, : L2[i]].tobytes() + b"\n")
    t1, t2 = b"".join(recs[0]), b"".join(recs[1])
    flags = ["-qcut", "15", "-ncut", "7", "-min_len", "30"]
    want, summary = oracle_cli(oracle_build, tmp_path, t1, t2, flags)
    params, min_len, _ = parse_flags(flags)
    outs, chunks, carry, (counts, unknown) = run_stream(sp, t1, t2, params, min_len, True, text_cap=1 << 20, max_pairs=4096)
    assert len(chunks) > 5 and carry == [b"", b""]
    for k in range(4):
        assert outs[k] == want[k], k
    # the consensus sequences the summary prints are the per-position majority bases of the counters (TrimmingStatistics::writeStatistics)
    assert not unknown and counts.sum() > 0
