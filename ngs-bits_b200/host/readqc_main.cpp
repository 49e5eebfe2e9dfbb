// readqc_b200 -- the ReadQC tool of imgag/ngs-bits (src/ReadQC/main.cpp) on the B200 engine: QC metrics of unprocessed reads.
//
// Same flags as the reference for what is supported: -in1 <files> [-in2 <files>] [-out <qcML or txt>] [-txt] [-gpus 0,1] [-threads N].
// The host only inflates (BGZF inputs in parallel with -threads); line framing, record checks (FastqEntry::validate,
// src/cppNGS/FastqFileStream.cpp:3-48) and the statistics of StatisticsReads::update (src/cppNGS/StatisticsReads.cpp:26-81) run on
// the device: spg_fq_* with stats_only (+ single_end without -in2), the reduction is spg::qc_kernel. The report holds the eight quality
// parameters of StatisticsReads::getResult; the plots are not rendered (QcReport.h). Not offered: -out1/-out2 (a copy of the
// input), -long_read (reads of 1000 bases and more).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iostream>
#include <memory>
#include <thread>

#include "../../include/seqpurge_b200.h"
#include "ChunkReader.h"
#include "GzipTextWriter.h"
#include "QcReport.h"
#include "SeqPurgeTypes.h"

using namespace seqpurge;

namespace
{

// the message FastqEntry::validate raises for this entry (FastqFileStream.cpp:3-48, short reads), or "" if it is valid
std::string validationError(const FastqEntry& e)
{
	const std::string message = "Invalid Fastq file entry: ";
	if (e.header.empty() || e.header[0] != '@') return message + "First header line does not start with '@': '" + e.header + "'.";
	if (e.header2.empty() || e.header2[0] != '+') return message + "Second header line does not start with '+': '" + e.header2 + "'.";
	if (e.bases.size() != e.qualities.size())
		return message + "Differing length of bases (" + std::to_string(e.bases.size()) + ") and qualities string (" + std::to_string(e.bases.size()) + ") in sequence '" + e.header + "'.";
	for (char c : e.bases)
		if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') return message + "Invalid base '" + c + "' encountered in sequence '" + e.header + "'.";
	for (char c : e.qualities)
		if ((int)c < 33 || (int)c > 74) return message + "Invalid quality character '" + c + "' with value '" + std::to_string((int)c) + "' encountered in sequence '" + e.header + "'.";
	return "";
}

// first invalid entry of a chunk, in reading order (the device only reports that there is one)
[[noreturn]] void throwFirstInvalid(const TextChunk* a, const TextChunk* b)
{
	for (const TextChunk* c : {a, b})
	{
		if (!c) continue;
		for (int i = 0; i < c->records; ++i)
		{
			const FastqEntry e = entryAt(*c, i);
			if (e.bases.size() >= (size_t)MAXLEN) throw ArgumentException("Read length unsupported! A maximum read length of " + std::to_string(MAXLEN) + " is supported (no -long_read)!");
			const std::string m = validationError(e);
			if (!m.empty()) throw FileParseException(m);
		}
	}
	throw ProgrammingException("the device rejected a chunk that the host finds valid");
}

struct Options
{
	std::vector<std::string> in1, in2;
	std::string out;
	bool txt = false;
	std::vector<int> gpus{0};
	int threads = 1;
	int block_size = 32768;
};

} // namespace

int main(int argc, char** argv)
{
	try
	{
		Options o;
		std::vector<std::string> args(argv + 1, argv + argc);
		for (size_t i = 0; i < args.size(); ++i)
		{
			const std::string f = args[i];
			auto list = [&](std::vector<std::string>& dst) {
				while (i + 1 < args.size() && args[i + 1][0] != '-') dst.push_back(args[++i]);
			};
			auto next = [&]() -> std::string {
				if (i + 1 >= args.size()) throw CommandLineParsingException("Parameter '" + f + "' needs a value!");
				return args[++i];
			};
			if (f == "--help" || f == "-h")
			{
				std::cout << "readqc_b200: QC metrics on unprocessed NGS reads (ReadQC of ngs-bits on the B200 engine).\n"
				             "  -in1 <files>   forward input FASTQ file(s), plain / gzip / BGZF\n"
				             "  -in2 <files>   reverse input FASTQ file(s) for paired-end mode (same number of files and reads as -in1)\n"
				             "  -out <file>    output qcML file (STDOUT if unset)\n"
				             "  -txt           TXT format instead of qcML\n"
				             "  -gpus 0,1      devices;  -threads N  inflate threads for BGZF inputs\n";
				return 0;
			}
			else if (f == "-in1") list(o.in1);
			else if (f == "-in2") list(o.in2);
			else if (f == "-out") o.out = next();
			else if (f == "-txt") o.txt = true;
			else if (f == "-threads") o.threads = atoi(next().c_str());
			else if (f == "-block_size") o.block_size = atoi(next().c_str());
			else if (f == "-gpus")
			{
				o.gpus.clear();
				std::string v = next();
				size_t p = 0;
				while (p <= v.size())
				{
					size_t q = v.find(',', p);
					if (q == std::string::npos) q = v.size();
					o.gpus.push_back(atoi(v.substr(p, q - p).c_str()));
					p = q + 1;
				}
			}
			else if (f == "-out1" || f == "-out2" || f == "-long_read" || f == "-compression_level")
				throw CommandLineParsingException("Parameter '" + f + "' is not supported by readqc_b200.");
			else throw CommandLineParsingException("Unknown parameter '" + f + "'!");
		}
		if (o.in1.empty()) throw CommandLineParsingException("Mandatory parameter: -in1 (see --help)");
		if (!o.in2.empty() && o.in1.size() != o.in2.size()) throw CommandLineParsingException("Input file lists 'in1' and 'in2' differ in counts!");
		if (o.block_size < 1 || o.gpus.empty()) throw CommandLineParsingException("block_size and gpus must be positive!");
		const bool paired = !o.in2.empty();

		std::unique_ptr<WorkerPool> pool;
		if (o.threads > 1) pool.reset(new WorkerPool(o.threads));
		ChunkQueue q1(4), q2(4);
		std::thread reader1([&]() { readerLoop(o.in1, o.block_size, q1, pool.get()); });
		std::thread reader2;
		if (paired) reader2 = std::thread([&]() { readerLoop(o.in2, o.block_size, q2, pool.get()); });
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

		// the engine: only the statistics part is used; qc = 2 adds the character checks of FastqEntry::validate
		spg_params ep;
		memset(&ep, 0, sizeof(ep));
		const std::string a1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA", a2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT";
		ep.a1 = a1.c_str();
		ep.a1_len = (int)a1.size();
		ep.a2 = a2.c_str();
		ep.a2_len = (int)a2.size();
		ep.adapter_overlap = 10;
		ep.match_perc = 80.0;
		ep.mep = 1e-6;
		ep.qcut = 15;
		ep.qwin = 5;
		ep.qoff = 33;
		ep.ncut = 7;
		ep.qc = 2;
		spg_ctx* engine = nullptr;
		if (spg_create(&engine, &ep, o.gpus.data(), (int)o.gpus.size(), 0, 0, 0) != SPG_OK)
			throw Exception(std::string("Could not initialize the CUDA engine: ") + spg_last_error(nullptr));
		struct EngineGuard
		{
			spg_ctx* e;
			~EngineGuard() { spg_destroy(e); }
		} engine_guard{engine};

		const int n_slots = 3 * (int)o.gpus.size();
		spg_fq* fq = nullptr;
		int fq_max_len = 0;
		int64_t fq_text_cap = 0;
		struct InFlight
		{
			int slot;
			std::unique_ptr<TextChunk> a, b;
		};
		std::deque<InFlight> in_flight;
		int next_slot = 0;
		auto retire = [&]() {
			InFlight f = std::move(in_flight.front());
			in_flight.pop_front();
			spg_fq_output out;
			if (spg_fq_wait(fq, f.slot, &out) != SPG_OK) throw Exception(spg_last_error(engine));
			if (out.error_pair >= 0 || out.invalid_chars) throwFirstInvalid(f.a.get(), f.b.get());
			if (out.n_pairs != f.a->records || (size_t)out.consumed1 != f.a->data.size() || (f.b && (size_t)out.consumed2 != f.b->data.size()))
				throw ProgrammingException("the device framed the chunk differently from the reader");
		};

		std::vector<std::string> infiles;
		size_t files_seen = 0;
		for (;;)
		{
			std::unique_ptr<TextChunk> a = q1.pop(), b = paired ? q2.pop() : nullptr;
			if (!a || (paired && !b))
			{
				if (q1.failure()) std::rethrow_exception(q1.failure());
				if (paired && q2.failure()) std::rethrow_exception(q2.failure());
				if (a || b) throw ProgrammingException("input file lists ended at different chunks");
				break;
			}
			while (files_seen <= a->file_index) // source files in the order the reference lists them (main.cpp:73,99)
			{
				infiles.push_back(o.in1[files_seen]);
				if (paired) infiles.push_back(o.in2[files_seen]);
				++files_seen;
			}
			if (paired)
			{
				// the reference reads the two files one after the other and compares the entry counts at the end (main.cpp:93-97)
				while (a->file_end != b->file_end)
				{
					std::unique_ptr<TextChunk>& open = a->file_end ? b : a;
					std::unique_ptr<TextChunk> nxt = (a->file_end ? q2 : q1).pop();
					if (!nxt) throw ProgrammingException("reader ended inside a file");
					if (nxt->records > 0 || a->records != b->records)
						throw ArgumentException("Differing number of reads in file '" + o.in1[a->file_index] + "' and '" + o.in2[a->file_index] + "'!");
					open->file_end = nxt->file_end;
				}
				if (a->records != b->records) throw ArgumentException("Differing number of reads in file '" + o.in1[a->file_index] + "' and '" + o.in2[a->file_index] + "'!");
			}
			if (a->records == 0) continue;
			const int need_len = std::min(std::max(a->max_read_len, b ? b->max_read_len : 0), MAXLEN - 1);
			const int64_t need_text = (int64_t)std::max(a->data.size(), b ? b->data.size() : (size_t)0);
			if (!fq || need_len > fq_max_len || need_text > fq_text_cap)
			{
				while (!in_flight.empty()) retire();
				if (fq) spg_fq_close(fq);
				fq = nullptr;
				spg_fq_config cfg;
				memset(&cfg, 0, sizeof(cfg));
				cfg.n_slots = n_slots;
				cfg.max_pairs = o.block_size;
				cfg.max_len = std::min(MAXLEN - 1, std::max(std::max((need_len + 15) / 16 * 16, fq_max_len), 160));
				cfg.text_cap = std::max<int64_t>(std::max<int64_t>(need_text + need_text / 4, fq_text_cap), 1 << 20);
				cfg.stats_only = 1;
				cfg.single_end = paired ? 0 : 1;
				cfg.validate = 1;
				if (spg_fq_open(engine, &cfg, &fq) != SPG_OK) throw Exception(std::string("Could not open the FASTQ stream on the device: ") + spg_last_error(engine));
				fq_max_len = cfg.max_len;
				fq_text_cap = cfg.text_cap;
				next_slot = 0;
			}
			if ((int)in_flight.size() == n_slots) retire();
			const int slot = next_slot;
			next_slot = (next_slot + 1) % n_slots;
			spg_fq_input in;
			if (spg_fq_buffers(fq, slot, &in) != SPG_OK) throw Exception(spg_last_error(engine));
			memcpy(in.text1, a->data.data(), a->data.size());
			if (b) memcpy(in.text2, b->data.data(), b->data.size());
			if (spg_fq_submit(fq, slot, (int64_t)a->data.size(), b ? (int64_t)b->data.size() : 0, 1, 1) != SPG_OK) throw Exception(spg_last_error(engine));
			in_flight.push_back(InFlight{slot, std::move(a), std::move(b)});
		}
		while (!in_flight.empty()) retire();
		if (fq) spg_fq_close(fq);

		std::unique_ptr<spg_qc_stats> stats(new spg_qc_stats);
		memset(stats.get(), 0, sizeof(spg_qc_stats));
		if (spg_qc_stats_get(engine, stats.get()) != SPG_OK) throw Exception(spg_last_error(engine));
		if (stats->errors != 0) throw ProgrammingException("the device counted an invalid character that the record checks did not report");
		if (o.txt)
		{
			std::ofstream file;
			if (!o.out.empty())
			{
				file.open(o.out);
				if (!file) throw FileAccessException("Could not open file '" + o.out + "' for writing!");
			}
			std::ostream& os = o.out.empty() ? std::cout : file;
			for (const auto& kv : qcMetrics(*stats)) os << kv.first << ": " << kv.second << "\n"; // QCCollection::appendToStringList
		}
		else
		{
			storeQcML(o.out.empty() ? "/dev/stdout" : o.out, *stats, infiles, "", "readqc_b200"); // the reference passes an empty parameter string, too (main.cpp:113)
		}
		return 0;
	}
	catch (const Exception& e)
	{
		std::cerr << e.what() << std::endl;
		return 1;
	}
	catch (const std::exception& e) // std::bad_alloc, std::system_error of a thread that could not be started, ...
	{
		std::cerr << "Error: " << e.what() << std::endl;
		return 1;
	}
}
