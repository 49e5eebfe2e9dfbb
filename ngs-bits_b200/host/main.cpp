// seqpurge_b200 -- command line with the flags of the reference's SeqPurge (src/SeqPurge/main.cpp:17-54, doc/tools/SeqPurge.md).
//
// Default pipeline (StreamPipeline.cpp): the host inflates the inputs and deflates the outputs; FASTQ framing, trimming, routing by
// -min_len, record layout, adapter consensus and -qc statistics run on the GPUs behind spg_fq_* of the C ABI.
// -host_framing selects the reference's block pipeline instead (load -> analyze -> write, src/SeqPurge/ThreadCoordinator.cpp:83-106;
// output routing and statistics of src/SeqPurge/OutputWorker.cpp:36-77, src/SeqPurge/FastqWriter.cpp:17-38) with records parsed and
// formatted on the host and only the analysis step handed to the CUDA engine through GpuAnalysisWorker -- the binding INTEGRATION.md
// describes for the reference itself.
//
// Blocks are dealt to the GPUs round robin (slot s -> device s % n) and retired in submission order, so the output equals the
// reference's `-threads 1` output whatever the number of GPUs. -threads N (N > 1) deflates the output with N threads (same content,
// other .gz bytes). -qc writes the read statistics as qcML (values only, no plots). Not supported here: -debug, -progress.
#include <algorithm>
#include <cerrno>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <exception>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>

#include "FastqFileStream.h"
#include "GpuAnalysisWorker.h"
#include "QcReport.h"
#include "StreamPipeline.h"

using namespace seqpurge;

namespace
{

struct InputStreams
{
	size_t current_index = 0;
	std::unique_ptr<FastqFileStream> istream1, istream2;
};

// InputWorker::run (src/SeqPurge/InputWorker.cpp:16-77). Returns true once the end of the data has been reached.
bool loadJob(AnalysisJob& job, InputStreams& streams, const TrimmingParameters& params)
{
	job.clear();
	bool end_of_data_reached = false;
	int pairs_read = 0;
	while (pairs_read < params.block_size && !end_of_data_reached)
	{
		if (streams.istream1->atEnd() && streams.istream2->atEnd())
		{
			++streams.current_index;
			if (streams.current_index >= params.files_in1.size()) end_of_data_reached = true;
			else
			{
				streams.istream1.reset(new FastqFileStream(params.files_in1[streams.current_index]));
				streams.istream2.reset(new FastqFileStream(params.files_in2[streams.current_index]));
			}
		}
		else if (streams.istream1->atEnd()) throw FileParseException("File " + streams.istream2->filename() + " has more entries than " + streams.istream1->filename() + "!");
		else if (streams.istream2->atEnd()) throw FileParseException("File " + streams.istream1->filename() + " has more entries than " + streams.istream2->filename() + "!");
		if (!end_of_data_reached)
		{
			streams.istream1->readEntry(job.r1[(size_t)pairs_read]);
			streams.istream2->readEntry(job.r2[(size_t)pairs_read]);
			++pairs_read;
		}
	}
	if (pairs_read > 0)
	{
		job.status = TO_BE_ANALYZED;
		job.read_count = pairs_read;
	}
	else job.read_count = 0;
	return end_of_data_reached;
}

struct OutputStreams
{
	std::unique_ptr<FastqOutfileStream> ostream1, ostream2, ostream3, ostream4;
};

// OutputWorker::run + FastqWriter::run for the forward reads (src/SeqPurge/OutputWorker.cpp:19-103, src/SeqPurge/FastqWriter.cpp:17-38);
// the reverse reads of complete pairs are written by the second writer thread, like the reference's second FastqWriter
void writeJob(AnalysisJob& job, OutputStreams& streams, const TrimmingParameters& params, TrimmingStatistics& stats)
{
	int reads_removed = 0;
	const size_t min_len = (size_t)std::max(params.min_len, 0);
	for (int r = 0; r < job.read_count; ++r)
	{
		const FastqEntry& e1 = job.r1[(size_t)r];
		const FastqEntry& e2 = job.r2[(size_t)r];
		if (e1.bases.size() >= min_len && e2.bases.size() >= min_len)
		{
			streams.ostream1->write(e1, e1.bases.size());
		}
		else if (streams.ostream3 && e1.bases.size() >= min_len)
		{
			reads_removed += 1;
			streams.ostream3->write(e1, e1.bases.size());
		}
		else if (streams.ostream4 && e2.bases.size() >= min_len)
		{
			reads_removed += 1;
			streams.ostream4->write(e2, e2.bases.size());
		}
		else reads_removed += 2;
	}
	stats.read_num += 2LL * job.read_count;
	stats.reads_trimmed_insert += job.reads_trimmed_insert;
	stats.reads_trimmed_adapter += job.reads_trimmed_adapter;
	stats.reads_trimmed_n += job.reads_trimmed_n;
	stats.reads_trimmed_q += job.reads_trimmed_q;
	stats.reads_removed += reads_removed;
	for (int r = 0; r < job.read_count; ++r)
	{
		const size_t l1 = job.r1[(size_t)r].bases.size(), l2 = job.r2[(size_t)r].bases.size();
		stats.bases_remaining[l1] += 1;
		stats.bases_remaining[l2] += 1;
		const int o1 = job.length_r1_orig[(size_t)r], o2 = job.length_r2_orig[(size_t)r];
		if (o1 > 0) stats.bases_perc_trim_sum += (double)(o1 - (int)l1) / o1;
		if (o2 > 0) stats.bases_perc_trim_sum += (double)(o2 - (int)l2) / o2;
	}
	job.status = DONE;
}

int maxReadLength(const AnalysisJob& job)
{
	size_t m = 0;
	for (int r = 0; r < job.read_count; ++r) m = std::max(m, std::max(job.r1[(size_t)r].bases.size(), job.r2[(size_t)r].bases.size()));
	return (int)m;
}

std::vector<int> parseIntList(const std::string& s)
{
	std::vector<int> v;
	std::stringstream ss(s);
	std::string tok;
	while (std::getline(ss, tok, ','))
	{
		char* end = nullptr;
		const long x = strtol(tok.c_str(), &end, 10);
		if (tok.empty() || *end != '\0' || x < 0 || x > 1023) throw CommandLineParsingException("'" + tok + "' in the device list '" + s + "' is not a CUDA device index!");
		if (std::find(v.begin(), v.end(), (int)x) != v.end()) throw CommandLineParsingException("Device " + tok + " is listed twice in '" + s + "'!");
		v.push_back((int)x);
	}
	return v;
}

void usage()
{
	std::cout << "seqpurge_b200: removes adapter sequences from paired-end sequencing data (SeqPurge on B200 GPUs).\n"
	             "Mandatory: -in1 <files> -in2 <files> -out1 <file> -out2 <file>\n"
	             "Optional (defaults of SeqPurge): -a1 -a2 -match_perc 80 -mep 0.000001 -qcut 15 -qwin 5 -qoff 33 -ncut 7 -min_len 30 -threads 1\n"
	             "          -out3 <prefix> -summary <file> -qc <file.qcML> -block_size 10000 -block_prefetch 32 -ec -compression_level 1\n"
	             "New: -gpus 0[,1,...]  CUDA devices the blocks are dealt to (default 0)\n"
	             "     -threads N       N > 1: the output files are deflated by N threads (same content, different .gz bytes)\n"
	             "     -bgzf            write the outputs as BGZF (blocked gzip, readable by any gzip reader, inflatable in parallel); BGZF inputs\n"
	             "                      are inflated by the -threads pool\n"
	             "     -host_framing    parse and format FASTQ records on the host (the reference's block pipeline) instead of on the device\n";
}

} // namespace

int main(int argc, char** argv)
{
	TrimmingParameters params;
	try
	{
		for (int i = 1; i < argc; ++i)
		{
			const std::string f = argv[i];
			auto next = [&]() -> std::string {
				if (i + 1 >= argc) throw CommandLineParsingException("Parameter '" + f + "' needs a value!");
				return argv[++i];
			};
			// numbers like ToolBase parses them (QString::toInt / toDouble with an ok flag): anything that is not a number is an error
			auto nextInt = [&]() -> int {
				const std::string v = next();
				char* end = nullptr;
				errno = 0;
				const long x = strtol(v.c_str(), &end, 10);
				if (v.empty() || *end != '\0' || errno != 0 || x < INT_MIN || x > INT_MAX)
					throw CommandLineParsingException("Value '" + v + "' of parameter '" + f + "' is not an integer!");
				return (int)x;
			};
			auto nextDouble = [&]() -> double {
				const std::string v = next();
				char* end = nullptr;
				errno = 0;
				const double x = strtod(v.c_str(), &end);
				if (v.empty() || *end != '\0' || errno != 0) throw CommandLineParsingException("Value '" + v + "' of parameter '" + f + "' is not a number!");
				return x;
			};
			auto trimmed = [](std::string v) { // QByteArray::trimmed(), src/SeqPurge/main.cpp:67-69
				const char* ws = " \t\n\v\f\r";
				const size_t a = v.find_first_not_of(ws);
				if (a == std::string::npos) return std::string();
				return v.substr(a, v.find_last_not_of(ws) - a + 1);
			};
			auto list = [&](std::vector<std::string>& dst) {
				while (i + 1 < argc && argv[i + 1][0] != '-') dst.push_back(argv[++i]);
			};
			if (f == "--help" || f == "-help")
			{
				usage();
				return 0;
			}
			else if (f == "-in1") list(params.files_in1);
			else if (f == "-in2") list(params.files_in2);
			else if (f == "-out1") params.out1 = next();
			else if (f == "-out2") params.out2 = next();
			else if (f == "-out3") params.out3 = next();
			else if (f == "-summary") params.summary = next();
			else if (f == "-a1") params.a1 = trimmed(next());
			else if (f == "-a2") params.a2 = trimmed(next());
			else if (f == "-match_perc") params.match_perc = nextDouble();
			else if (f == "-mep") params.mep = nextDouble();
			else if (f == "-qcut") params.qcut = nextInt();
			else if (f == "-qwin") params.qwin = nextInt();
			else if (f == "-qoff") params.qoff = nextInt();
			else if (f == "-ncut") params.ncut = nextInt();
			else if (f == "-min_len") params.min_len = nextInt();
			else if (f == "-threads") params.threads = nextInt();
			else if (f == "-block_size") params.block_size = nextInt();
			else if (f == "-block_prefetch") params.block_prefetch = nextInt();
			else if (f == "-compression_level") params.compression_level = nextInt();
			else if (f == "-ec") params.ec = true;
			else if (f == "-gpus") params.gpus = parseIntList(next());
			else if (f == "-qc") params.qc = next();
			else if (f == "-host_framing") params.host_framing = true;
			else if (f == "-bgzf") params.bgzf = true;
			else if (f == "-debug" || f == "-progress") throw CommandLineParsingException("Parameter '" + f + "' is not supported by seqpurge_b200.");
			else throw CommandLineParsingException("Unknown parameter '" + f + "'!");
		}
		if (params.files_in1.empty() || params.files_in2.empty() || params.out1.empty() || params.out2.empty())
			throw CommandLineParsingException("Mandatory parameters: -in1 -in2 -out1 -out2 (see --help)");
		if (params.files_in1.size() != params.files_in2.size()) throw CommandLineParsingException("Input file lists 'in1' and 'in2' differ in counts!");
		if (params.a1.size() < 15) throw CommandLineParsingException("Forward adapter " + params.a1 + " too short!");
		if (params.a2.size() < 15) throw CommandLineParsingException("Reverse adapter " + params.a2 + " too short!");
		params.a_size = (int)std::min<size_t>(20, std::min(params.a1.size(), params.a2.size()));
		if (params.block_size < 1 || params.block_prefetch < 1 || params.gpus.empty()) throw CommandLineParsingException("block_size, block_prefetch and gpus must be positive!");

		const auto t_start = std::chrono::steady_clock::now();
		if (!params.host_framing) // default: FASTQ framing and output assembly on the device (StreamPipeline.cpp)
		{
			std::ofstream summary_file;
			if (!params.summary.empty())
			{
				summary_file.open(params.summary);
				if (!summary_file) throw FileAccessException("Could not open file '" + params.summary + "' for writing!");
			}
			std::ostream& summary = params.summary.empty() ? std::cout : summary_file;
			TrimmingStatistics stats;
			ErrorCorrectionStatistics ec_stats;
			std::unique_ptr<spg_qc_stats> qc_stats(new spg_qc_stats());
			memset(qc_stats.get(), 0, sizeof(spg_qc_stats));
			runStreamPipeline(params, summary, stats, ec_stats, qc_stats.get());
			stats.writeStatistics(summary, params);
			if (!params.qc.empty())
			{
				std::vector<std::string> sources = params.files_in1;
				sources.insert(sources.end(), params.files_in2.begin(), params.files_in2.end());
				storeQcML(params.qc, *qc_stats, sources, "");
			}
			if (params.ec) ec_stats.writeStatistics(summary);
			const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
			char buf[64];
			snprintf(buf, sizeof(buf), "%.3fs", secs);
			summary << "overall runtime: " << buf << "\n";
			return 0;
		}
		InputStreams in;
		in.istream1.reset(new FastqFileStream(params.files_in1[0]));
		in.istream2.reset(new FastqFileStream(params.files_in2[0]));
		OutputStreams out;
		out.ostream1.reset(new FastqOutfileStream(params.out1, params.compression_level));
		out.ostream2.reset(new FastqOutfileStream(params.out2, params.compression_level));
		if (!params.out3.empty())
		{
			out.ostream3.reset(new FastqOutfileStream(params.out3 + "_R1.fastq.gz", params.compression_level));
			out.ostream4.reset(new FastqOutfileStream(params.out3 + "_R2.fastq.gz", params.compression_level));
		}
		std::ofstream summary_file;
		if (!params.summary.empty())
		{
			summary_file.open(params.summary);
			if (!summary_file) throw FileAccessException("Could not open file '" + params.summary + "' for writing!");
		}
		std::ostream& summary = params.summary.empty() ? std::cout : summary_file;

		TrimmingStatistics stats;
		ErrorCorrectionStatistics ec_stats;
		const int n_jobs = params.block_prefetch;
		std::vector<AnalysisJob> job_pool;
		job_pool.reserve((size_t)n_jobs);
		for (int i = 0; i < n_jobs; ++i) job_pool.emplace_back(i, params.block_size);
		std::vector<std::unique_ptr<GpuAnalysisWorker>> workers((size_t)n_jobs);

		spg_ctx* engine = nullptr;
		int engine_max_len = 0;
		spg_params ep = toEngineParams(params);
		std::unique_ptr<spg_qc_stats> qc_stats(new spg_qc_stats());
		memset(qc_stats.get(), 0, sizeof(spg_qc_stats));
		auto accumulateEc = [&]() { // takes the -ec histograms and the -qc statistics out of an engine before it goes away
			if (!engine) return;
			if (!params.qc.empty())
			{
				std::unique_ptr<spg_qc_stats> q(new spg_qc_stats());
				if (spg_qc_stats_get(engine, q.get()) != SPG_OK) throw Exception(spg_last_error(engine));
				qcAccumulate(*qc_stats, *q);
			}
			if (!params.ec) return;
			spg_ec_stats s;
			if (spg_ec_stats_get(engine, &s) != SPG_OK) throw Exception(spg_last_error(engine));
			for (int i = 0; i < MAXLEN; ++i)
			{
				ec_stats.mismatch_r1[(size_t)i] += s.mismatch_r1[i];
				ec_stats.mismatch_r2[(size_t)i] += s.mismatch_r2[i];
				ec_stats.errors_per_read[(size_t)i] += s.errors_per_read[i];
			}
		};
		auto createEngine = [&](int max_len) {
			if (engine)
			{
				accumulateEc();
				spg_destroy(engine);
				engine = nullptr;
			}
			if (spg_create(&engine, &ep, params.gpus.data(), (int)params.gpus.size(), n_jobs, params.block_size, max_len) != SPG_OK)
				throw Exception(std::string("Could not initialize the CUDA trimming engine: ") + spg_last_error(nullptr));
			spg_set_option(engine, SPG_OPT_QUAL_TAILS, 1); // GpuAnalysisWorker::start writes the slots' quality tails
			engine_max_len = max_len;
		};

		// ---- pipeline: reader thread -> this thread (GPU hand-off, in submission order) -> two writer threads ------------------------------
		// Same stages as the reference (1 reader, analysis, 1 writer with one thread per output stream; ThreadCoordinator.cpp:40-42,
		// OutputWorker.cpp:24-32), but jobs retire strictly in input order.
		enum { FREE, LOADED, IN_FLIGHT, ANALYZED };
		std::mutex mu;
		std::condition_variable cv;
		std::vector<int> state((size_t)n_jobs, FREE);
		std::vector<int> writers_left((size_t)n_jobs, 0);
		std::deque<int> loaded, to_write1, to_write2;
		bool reader_done = false, analysis_done = false, abort_all = false;
		std::exception_ptr failure;
		auto fail = [&](std::exception_ptr e) {
			std::lock_guard<std::mutex> g(mu);
			if (!failure) failure = e;
			abort_all = true;
			cv.notify_all();
		};

		std::thread reader([&]() {
			try
			{
				int j = 0;
				bool end = false;
				while (!end)
				{
					{
						std::unique_lock<std::mutex> l(mu);
						cv.wait(l, [&] { return state[(size_t)j] == FREE || abort_all; });
						if (abort_all) return;
					}
					end = loadJob(job_pool[(size_t)j], in, params);
					if (job_pool[(size_t)j].read_count > 0)
					{
						std::lock_guard<std::mutex> g(mu);
						state[(size_t)j] = LOADED;
						loaded.push_back(j);
						cv.notify_all();
						j = (j + 1) % n_jobs;
					}
				}
				std::lock_guard<std::mutex> g(mu);
				reader_done = true;
				cv.notify_all();
			}
			catch (...)
			{
				fail(std::current_exception());
			}
		});

		auto writer = [&](std::deque<int>& queue, bool first) {
			try
			{
				for (;;)
				{
					int j;
					{
						std::unique_lock<std::mutex> l(mu);
						cv.wait(l, [&] { return !queue.empty() || analysis_done || abort_all; });
						if (abort_all) return;
						if (queue.empty()) return; // analysis_done
						j = queue.front();
						queue.pop_front();
					}
					AnalysisJob& job = job_pool[(size_t)j];
					if (first) writeJob(job, out, params, stats); // out1, singletons, statistics
					else
					{
						const size_t min_len = (size_t)std::max(params.min_len, 0);
						for (int r = 0; r < job.read_count; ++r) // FastqWriter::run for the reverse reads
						{
							const FastqEntry& e1 = job.r1[(size_t)r];
							const FastqEntry& e2 = job.r2[(size_t)r];
							if (e1.bases.size() >= min_len && e2.bases.size() >= min_len) out.ostream2->write(e2, e2.bases.size());
						}
					}
					std::lock_guard<std::mutex> g(mu);
					if (--writers_left[(size_t)j] == 0)
					{
						state[(size_t)j] = FREE;
						cv.notify_all();
					}
				}
			}
			catch (...)
			{
				fail(std::current_exception());
			}
		};
		std::thread writer1(writer, std::ref(to_write1), true);
		std::thread writer2(writer, std::ref(to_write2), false);

		std::deque<int> in_flight; // job indices in submission order
		auto retire = [&]() {
			const int j = in_flight.front();
			in_flight.pop_front();
			workers[(size_t)j]->wait();
			std::lock_guard<std::mutex> g(mu);
			state[(size_t)j] = ANALYZED;
			writers_left[(size_t)j] = 2;
			to_write1.push_back(j);
			to_write2.push_back(j);
			cv.notify_all();
		};
		try
		{
			for (;;)
			{
				int j = -1;
				{
					std::unique_lock<std::mutex> l(mu);
					// start a loaded job if there is one; otherwise retire the oldest job in flight; otherwise wait for the reader
					cv.wait(l, [&] { return !loaded.empty() || !in_flight.empty() || reader_done || abort_all; });
					if (abort_all) break;
					if (!loaded.empty())
					{
						j = loaded.front();
						loaded.pop_front();
					}
					else if (in_flight.empty() && reader_done) break;
				}
				if (j < 0)
				{
					retire();
					continue;
				}
				AnalysisJob& job = job_pool[(size_t)j];
				const int need = std::min(maxReadLength(job), MAXLEN - 1);
				if (!engine || need > engine_max_len)
				{
					while (!in_flight.empty()) retire(); // drain before the slots are re-created for longer reads
					createEngine(std::min(MAXLEN - 1, std::max((need + 15) / 16 * 16, 160)));
					spg_set_option(engine, SPG_OPT_FULL_LEN, need); // the run's read length: picks the kernel variant compiled for it (a hint only)
				}
				workers[(size_t)j].reset(new GpuAnalysisWorker(job, params, stats, ec_stats, engine, j));
				workers[(size_t)j]->start();
				{
					std::lock_guard<std::mutex> g(mu);
					state[(size_t)j] = IN_FLIGHT;
				}
				in_flight.push_back(j);
			}
			while (!in_flight.empty() && !abort_all) retire();
		}
		catch (...)
		{
			fail(std::current_exception());
		}
		{
			std::lock_guard<std::mutex> g(mu);
			analysis_done = true;
			cv.notify_all();
		}
		reader.join();
		writer1.join();
		writer2.join();
		if (failure)
		{
			if (engine) spg_destroy(engine);
			std::rethrow_exception(failure);
		}
		accumulateEc();
		if (engine) spg_destroy(engine);

		out.ostream1->close();
		out.ostream2->close();
		if (out.ostream3) out.ostream3->close();
		if (out.ostream4) out.ostream4->close();

		stats.writeStatistics(summary, params);
		if (!params.qc.empty()) // ThreadCoordinator.cpp:137-141
		{
			std::vector<std::string> sources = params.files_in1;
			sources.insert(sources.end(), params.files_in2.begin(), params.files_in2.end());
			storeQcML(params.qc, *qc_stats, sources, ""); // the reference passes an empty parameter string, too
		}
		if (params.ec) ec_stats.writeStatistics(summary);
		const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
		char buf[64];
		snprintf(buf, sizeof(buf), "%.3fs", secs);
		summary << "overall runtime: " << buf << "\n";
		return 0;
	}
	catch (const std::exception& e)
	{
		std::cerr << "seqpurge_b200: " << e.what() << std::endl;
		return 1;
	}
}
