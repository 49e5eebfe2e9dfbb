// StreamPipeline.cpp -- the command line's default pipeline: the host only inflates and deflates, everything in between runs on the
// GPUs: FASTQ text -> spg_fq_* (framing, rows, [-qc], trimming, routing, record layout, consensus) -> FASTQ text.
//
//   2 reader threads   one per input file list: gzread, cut after `block_size` records (4 lines each, counted with memchr), note the
//                      longest sequence line -> queue of text chunks             (InputWorker::run, src/SeqPurge/InputWorker.cpp:16-77)
//   this thread        pairs the chunks of the two lists, copies them into a pinned slot, spg_fq_submit; retires slots in order:
//                      statistics from the 8-byte records, output text to the writers   (AnalysisWorker / OutputWorker seam)
//   writers            one GzipTextWriter per output file (+ a shared deflate pool with -threads N > 1)
#include "StreamPipeline.h"

#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>

#include "GpuAnalysisWorker.h"
#include "GzipTextWriter.h"
#include "QcReport.h"
#include "ChunkReader.h"
#include "TextSource.h"

namespace seqpurge
{

namespace
{

bool endsWith(const std::string& s, const char* suffix)
{
	const size_t n = strlen(suffix);
	return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}
bool isAcgtn(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N'; }
[[noreturn]] void throwBadComplement(const std::string& bases)
{
	for (char c : bases)
		if (!isAcgtn(c)) throw ProgrammingException(std::string("Could not convert base '") + c + "' to complement!"); // Sequence.cpp:68
	throw ProgrammingException("Could not convert base to complement!");
}

// the exception the reference's worker raises for this pair (same checks, same order as GpuAnalysisWorker::start / wait)
[[noreturn]] void throwPairError(const FastqEntry& e1, const FastqEntry& e2, int frame_status, int result_status)
{
	auto token = [](const std::string& h) {
		const size_t p = h.find(' ');
		return p == std::string::npos ? h : h.substr(0, p);
	};
	std::string t1 = token(e1.header), t2 = token(e2.header);
	if (endsWith(t1, "/1") && endsWith(t2, "/2"))
	{
		t1.resize(t1.size() - 2);
		t2.resize(t2.size() - 2);
	}
	if (t1 != t2) throw ArgumentException("Headers of reads do not match:\n" + t1 + "\n" + t2);
	if (e1.qualities.size() != e1.bases.size() || e2.qualities.size() != e2.bases.size())
		throw FileParseException("Differing length of bases and qualities string in sequence '" + e1.header + "'.");
	if (!std::all_of(e2.bases.begin(), e2.bases.end(), isAcgtn)) throwBadComplement(e2.bases);
	if (std::max(e1.bases.size(), e2.bases.size()) >= (size_t)MAXLEN)
		throw ArgumentException("Read length unsupported! A maximum read length of " + std::to_string(MAXLEN) + " is supported!");
	if (result_status == SPG_PAIR_BAD_BASE_EC) throwBadComplement(e1.bases);
	throw ProgrammingException("pair rejected by the device (frame status " + std::to_string(frame_status) + ", result status " + std::to_string(result_status) + ")");
}

} // namespace

void runStreamPipeline(const TrimmingParameters& params, std::ostream& summary, TrimmingStatistics& stats, ErrorCorrectionStatistics& ec_stats, spg_qc_stats* qc_stats)
{
	const int pairs = params.block_size;
	const bool singles = !params.out3.empty();
	// SPG_TIMING=1: where this thread spent its time (stderr)
	const bool timing = getenv("SPG_TIMING") != nullptr;
	using clk = std::chrono::steady_clock;
	double t_init = 0, t_pop = 0, t_copy = 0, t_wait = 0, t_stats = 0, t_write = 0, t_close = 0;
	auto since = [](clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); };
	clk::time_point tp = clk::now();
	const int n_slots = std::max(2, std::min(params.block_prefetch, 4)) * (int)params.gpus.size(); // chunks in flight: the GPU part of a chunk is far shorter than its inflate
	(void)summary;

	// The engine comes up first (CUDA context, decision tables: 0.3-1.5 s depending on the box). Started next to a pool of inflating threads it
	// took 1.2-3.0 s on one box (profiles/cli_throughput_r2.log): context creation maps memory all the time and competes with the page
	// faults of the readers for the address-space lock; the overlap it would buy is the inflate of one chunk.
	spg_params ep = toEngineParams(params);
	spg_ctx* engine = nullptr;
	if (spg_create(&engine, &ep, params.gpus.data(), (int)params.gpus.size(), 0, 0, 0) != SPG_OK)
		throw Exception(std::string("Could not initialize the CUDA trimming engine: ") + spg_last_error(nullptr));
	struct EngineGuard
	{
		spg_ctx* e;
		~EngineGuard() { spg_destroy(e); }
	} engine_guard{engine};
	t_init += since(tp);

	std::unique_ptr<WorkerPool> pool;
	if (params.threads > 1) pool.reset(new WorkerPool(params.threads));
	std::unique_ptr<GzipTextWriter> writers[4];
	writers[0].reset(new GzipTextWriter(params.out1, params.compression_level, pool.get(), params.bgzf));
	writers[1].reset(new GzipTextWriter(params.out2, params.compression_level, pool.get(), params.bgzf));
	if (singles)
	{
		writers[2].reset(new GzipTextWriter(params.out3 + "_R1.fastq.gz", params.compression_level, pool.get(), params.bgzf));
		writers[3].reset(new GzipTextWriter(params.out3 + "_R2.fastq.gz", params.compression_level, pool.get(), params.bgzf));
	}

	ChunkQueue q1(4), q2(4);
	std::thread reader1([&]() { readerLoop(params.files_in1, pairs, q1, pool.get()); });
	std::thread reader2([&]() { readerLoop(params.files_in2, pairs, q2, pool.get()); });
	struct ReaderGuard
	{
		ChunkQueue &a, &b;
		std::thread &t1, &t2;
		~ReaderGuard()
		{
			a.abort();
			b.abort();
			if (t1.joinable()) t1.join();
			if (t2.joinable()) t2.join();
		}
	} reader_guard{q1, q2, reader1, reader2};


	spg_fq* fq = nullptr;
	int fq_max_len = 0;
	int64_t fq_text_cap = 0;
	int64_t acons[2][40][5];
	memset(acons, 0, sizeof(acons));
	bool acons_unknown = false;
	auto closeStream = [&]() {
		if (!fq) return;
		int64_t c[2][40][5];
		int32_t unknown = 0;
		if (spg_fq_consensus_get(fq, c, &unknown) != SPG_OK) throw Exception(spg_last_error(engine));
		for (int k = 0; k < 2 * 40 * 5; ++k) (&acons[0][0][0])[k] += (&c[0][0][0])[k];
		if (unknown) acons_unknown = true;
		spg_fq_close(fq);
		fq = nullptr;
	};

	struct InFlight
	{
		int slot;
		std::unique_ptr<TextChunk> a, b;
	};
	// Chunks in flight, in submission order. This thread feeds (pairs chunks, copies them into a pinned slot, submits); a second thread
	// retires them IN ORDER (waits for the device, adds the counters, hands the output text to the writers), so that the two host-side
	// copies of a chunk overlap instead of adding up. `in_flight` counts submitted-and-not-yet-retired chunks (= slots in use).
	std::deque<InFlight> in_flight;
	std::mutex fl_mu;
	std::condition_variable fl_cv;
	bool fl_done = false;             // no more chunks will be submitted
	std::exception_ptr retire_error;  // first exception of the retire thread
	int next_slot = 0;

	auto retire_one = [&](InFlight& f) {
		spg_fq_output o;
		clk::time_point t0 = clk::now();
		if (spg_fq_wait(fq, f.slot, &o) != SPG_OK) throw Exception(spg_last_error(engine));
		t_wait += since(t0);
		t0 = clk::now();
		if (o.error_pair >= 0)
		{
			const int p = o.error_pair;
			throwPairError(entryAt(*f.a, p), entryAt(*f.b, p), o.frame_status[p], o.results[p].status);
		}
		if (o.n_pairs != f.a->records || (size_t)o.consumed1 != f.a->data.size() || (size_t)o.consumed2 != f.b->data.size())
			throw ProgrammingException("the device framed the chunk differently from the reader");
		// statistics (OutputWorker.cpp:59-77): reduced on the device from the result records (spg_fq_stats)
		if (!o.stats) throw ProgrammingException("the stream did not return summary counters");
		const spg_fq_stats& st = *o.stats;
		for (int len = 0; len < SPG_MAXLEN; ++len)
		{
			if (st.bases_remaining[len]) stats.bases_remaining[len] += (double)st.bases_remaining[len];
			if (len > 0 && st.trimmed_bases_by_length[len]) stats.bases_perc_trim_sum += (double)st.trimmed_bases_by_length[len] / len;
		}
		stats.read_num += 2LL * o.n_pairs;
		stats.reads_trimmed_insert += (double)st.reads_trimmed_insert;
		stats.reads_trimmed_adapter += (double)st.reads_trimmed_adapter;
		stats.reads_trimmed_q += (double)st.reads_trimmed_q;
		stats.reads_trimmed_n += (double)st.reads_trimmed_n;
		stats.reads_removed += (double)st.reads_removed;
		t_stats += since(t0);
		t0 = clk::now();
		// the two (or four) output texts are copied into the writers' pieces side by side
		std::thread second;
		if (writers[1] && o.out_bytes[1] > 0) second = std::thread([&]() { writers[1]->write(o.out[1], (size_t)o.out_bytes[1]); });
		try
		{
			for (int k = 0; k < 4; ++k)
				if (k != 1 && writers[k] && o.out_bytes[k] > 0) writers[k]->write(o.out[k], (size_t)o.out_bytes[k]);
		}
		catch (...)
		{
			if (second.joinable()) second.join();
			throw;
		}
		if (second.joinable()) second.join();
		t_write += since(t0);
	};
	std::thread retirer([&]() {
		try
		{
			for (;;)
			{
				std::unique_lock<std::mutex> l(fl_mu);
				fl_cv.wait(l, [&] { return !in_flight.empty() || fl_done; });
				if (in_flight.empty()) break;
				InFlight& f = in_flight.front(); // stays in the queue (its slot stays taken) until it is retired
				l.unlock();
				retire_one(f);
				l.lock();
				in_flight.pop_front();
				fl_cv.notify_all();
			}
		}
		catch (...)
		{
			std::lock_guard<std::mutex> l(fl_mu);
			retire_error = std::current_exception();
			fl_cv.notify_all();
		}
	});
	struct RetireGuard
	{
		std::thread& t;
		std::mutex& mu;
		std::condition_variable& cv;
		bool& done;
		~RetireGuard()
		{
			{
				std::lock_guard<std::mutex> l(mu);
				done = true;
			}
			cv.notify_all();
			if (t.joinable()) t.join();
		}
	} retire_guard{retirer, fl_mu, fl_cv, fl_done};
	// blocks until at most `keep` chunks are in flight; rethrows what the retire thread ran into
	auto drain_to = [&](size_t keep) {
		std::unique_lock<std::mutex> l(fl_mu);
		fl_cv.wait(l, [&] { return in_flight.size() <= keep || retire_error; });
		if (retire_error) std::rethrow_exception(retire_error);
	};

	auto moreEntries = [&](size_t fi, bool first_has_more) {
		const std::string& a = params.files_in1[std::min(fi, params.files_in1.size() - 1)];
		const std::string& b = params.files_in2[std::min(fi, params.files_in2.size() - 1)];
		// InputWorker.cpp:37-44
		if (first_has_more) throw FileParseException("File " + a + " has more entries than " + b + "!");
		throw FileParseException("File " + b + " has more entries than " + a + "!");
	};

	for (;;)
	{
		clk::time_point t0 = clk::now();
		std::unique_ptr<TextChunk> a = q1.pop(), b = q2.pop();
		t_pop += since(t0);
		if (!a || !b)
		{
			if (q1.failure()) std::rethrow_exception(q1.failure());
			if (q2.failure()) std::rethrow_exception(q2.failure());
			if (a || b) throw ProgrammingException("input file lists ended at different chunks");
			break;
		}
		if (a->file_index != b->file_index) throw ProgrammingException("readers out of step");
		if (a->records != b->records)
		{
			drain_to(0); // the pairs in front of the mismatch are processed like in the reference
			moreEntries(a->file_index, a->records > b->records);
		}
		// one file ended on this chunk, the other did not: whatever follows in the other file is a surplus entry
		while (a->file_end != b->file_end)
		{
			std::unique_ptr<TextChunk>& open = a->file_end ? b : a;
			ChunkQueue& q = a->file_end ? q2 : q1;
			std::unique_ptr<TextChunk> nxt = q.pop();
			if (!nxt) throw ProgrammingException("reader ended inside a file");
			if (nxt->records > 0)
			{
				drain_to(0);
				moreEntries(a->file_index, !a->file_end);
			}
			open->file_end = nxt->file_end;
		}
		if (a->records == 0) continue;

		const int need_len = std::min(std::max(a->max_read_len, b->max_read_len), MAXLEN - 1);
		const int64_t need_text = (int64_t)std::max(a->data.size(), b->data.size());
		if (!fq || need_len > fq_max_len || need_text > fq_text_cap)
		{
			drain_to(0);
			closeStream();
			spg_fq_config cfg;
			memset(&cfg, 0, sizeof(cfg));
			cfg.n_slots = n_slots;
			cfg.max_pairs = pairs;
			cfg.max_len = std::min(MAXLEN - 1, std::max(std::max((need_len + 15) / 16 * 16, fq_max_len), 160));
			cfg.text_cap = std::max<int64_t>(std::max<int64_t>(need_text + need_text / 4, fq_text_cap), 1 << 20);
			cfg.min_len = std::max(params.min_len, 0);
			cfg.singles = singles ? 1 : 0;
			t0 = clk::now();
			if (spg_fq_open(engine, &cfg, &fq) != SPG_OK) throw Exception(std::string("Could not open the FASTQ stream on the device: ") + spg_last_error(engine));
			t_init += since(t0);
			fq_max_len = cfg.max_len;
			fq_text_cap = cfg.text_cap;
			next_slot = 0;
			// the reads of a run share one length: the kernel variant compiled for it is picked from the first chunk on (a hint only,
			// results do not depend on it)
			spg_set_option(engine, SPG_OPT_FULL_LEN, need_len);
		}
		drain_to((size_t)n_slots - 1); // a free slot (slots are used round robin and retired in order)
		const int slot = next_slot;
		next_slot = (next_slot + 1) % n_slots;
		spg_fq_input in;
		if (spg_fq_buffers(fq, slot, &in) != SPG_OK) throw Exception(spg_last_error(engine));
		t0 = clk::now();
		{
			std::thread second([&]() { memcpy(in.text2, b->data.data(), b->data.size()); });
			memcpy(in.text1, a->data.data(), a->data.size());
			second.join();
		}
		t_copy += since(t0);
		if (spg_fq_submit(fq, slot, (int64_t)a->data.size(), (int64_t)b->data.size(), 1, 1) != SPG_OK) throw Exception(spg_last_error(engine));
		{
			std::lock_guard<std::mutex> l(fl_mu);
			in_flight.push_back(InFlight{slot, std::move(a), std::move(b)});
		}
		fl_cv.notify_all();
	}
	drain_to(0);
	{
		std::lock_guard<std::mutex> l(fl_mu);
		fl_done = true;
	}
	fl_cv.notify_all();
	retirer.join();
	closeStream();
	if (acons_unknown) throw ArgumentException("Unknown base in the adapter consensus window!"); // Pileup::inc
	for (int i = 0; i < 40; ++i)
	{
		BaseCounts* dst[2] = {&stats.acons1[(size_t)i], &stats.acons2[(size_t)i]};
		for (int rd = 0; rd < 2; ++rd)
		{
			dst[rd]->a += acons[rd][i][0];
			dst[rd]->c += acons[rd][i][1];
			dst[rd]->g += acons[rd][i][2];
			dst[rd]->t += acons[rd][i][3];
			dst[rd]->n += acons[rd][i][4];
		}
	}
	if (params.ec)
	{
		spg_ec_stats s;
		if (spg_ec_stats_get(engine, &s) != SPG_OK) throw Exception(spg_last_error(engine));
		for (int i = 0; i < MAXLEN; ++i)
		{
			ec_stats.mismatch_r1[(size_t)i] += s.mismatch_r1[i];
			ec_stats.mismatch_r2[(size_t)i] += s.mismatch_r2[i];
			ec_stats.errors_per_read[(size_t)i] += s.errors_per_read[i];
		}
	}
	if (qc_stats && !params.qc.empty())
	{
		if (spg_qc_stats_get(engine, qc_stats) != SPG_OK) throw Exception(spg_last_error(engine));
	}
	tp = clk::now();
	for (int k = 0; k < 4; ++k)
		if (writers[k]) writers[k]->close();
	t_close += since(tp);
	if (timing)
		fprintf(stderr, "[spg timing] init %.3f s, waiting for input %.3f s, copy to pinned %.3f s, waiting for the GPU %.3f s, statistics %.3f s, handing to writers %.3f s, closing writers %.3f s\n",
		        t_init, t_pop, t_copy, t_wait, t_stats, t_write, t_close);
}

} // namespace seqpurge
