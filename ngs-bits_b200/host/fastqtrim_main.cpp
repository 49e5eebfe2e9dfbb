// fastqtrim_b200 -- the FastqTrim tool of imgag/ngs-bits (src/FastqTrim/main.cpp) on the B200 engine: trims start/end bases from all
// reads of one FASTQ file.
//
// Flags of the reference: -in <file> -out <file> [-start N] [-end N] [-len N] [-max_len N] [-compression_level L]; new: -gpus, -threads N
// (parallel deflate, parallel inflate of BGZF input), -bgzf. The host inflates and deflates; line framing, the trimming rule of
// src/FastqTrim/main.cpp:47-77 and the record layout of FastqOutfileStream::write run on the device (spg_fq_* with single_end +
// fixed_trim). With -threads 1 the .gz bytes are those of the reference's writer (same zlib call sequence, GzipTextWriter.h).
// Not offered: -long_read (reads of 1000 bases and more). A record whose bases and qualities differ in length is an error here (the
// reference cuts the two strings independently).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <iostream>
#include <memory>
#include <thread>

#include "../../include/seqpurge_b200.h"
#include "ChunkReader.h"
#include "GzipTextWriter.h"
#include "SeqPurgeTypes.h"

using namespace seqpurge;

int main(int argc, char** argv)
{
	try
	{
		std::string in, out;
		int start = 0, end = 0, len = 0, max_len = 0, level = 1, threads = 1, block_size = 32768;
		bool bgzf = false;
		std::vector<int> gpus{0};
		std::vector<std::string> args(argv + 1, argv + argc);
		for (size_t i = 0; i < args.size(); ++i)
		{
			const std::string f = args[i];
			auto next = [&]() -> std::string {
				if (i + 1 >= args.size()) throw CommandLineParsingException("Parameter '" + f + "' needs a value!");
				return args[++i];
			};
			if (f == "--help" || f == "-h")
			{
				std::cout << "fastqtrim_b200: trims start/end bases from all reads in a FASTQ file (FastqTrim of ngs-bits on the B200 engine).\n"
				             "  -in <file> -out <file>  input / output FASTQ (input plain, gzip or BGZF; output gzip)\n"
				             "  -start N  -end N        bases to trim from the start / end of every read (reads that would be empty are removed)\n"
				             "  -len N                  restrict the read length to N afterwards\n"
				             "  -max_len N              only trim reads shorter than N\n"
				             "  -compression_level L    1 (default) .. 9;  -threads N  parallel deflate / BGZF inflate;  -bgzf  BGZF output;  -gpus 0,1\n";
				return 0;
			}
			else if (f == "-in") in = next();
			else if (f == "-out") out = next();
			else if (f == "-start") start = atoi(next().c_str());
			else if (f == "-end") end = atoi(next().c_str());
			else if (f == "-len") len = atoi(next().c_str());
			else if (f == "-max_len") max_len = atoi(next().c_str());
			else if (f == "-compression_level") level = atoi(next().c_str());
			else if (f == "-threads") threads = atoi(next().c_str());
			else if (f == "-block_size") block_size = atoi(next().c_str());
			else if (f == "-bgzf") bgzf = true;
			else if (f == "-gpus")
			{
				gpus.clear();
				std::string v = next();
				size_t p = 0;
				while (p <= v.size())
				{
					size_t q = v.find(',', p);
					if (q == std::string::npos) q = v.size();
					gpus.push_back(atoi(v.substr(p, q - p).c_str()));
					p = q + 1;
				}
			}
			else if (f == "-long_read") throw CommandLineParsingException("Parameter '-long_read' is not supported by fastqtrim_b200.");
			else throw CommandLineParsingException("Unknown parameter '" + f + "'!");
		}
		if (in.empty() || out.empty()) throw CommandLineParsingException("Mandatory parameters: -in -out (see --help)");
		if (start < 0 || end < 0 || len < 0 || max_len < 0 || block_size < 1 || gpus.empty()) throw CommandLineParsingException("start, end, len, max_len must not be negative!");

		std::unique_ptr<WorkerPool> pool;
		if (threads > 1) pool.reset(new WorkerPool(threads));
		GzipTextWriter writer(out, level, pool.get(), bgzf);
		ChunkQueue q(4);
		std::vector<std::string> files{in};
		std::thread reader([&]() { readerLoop(files, block_size, q, pool.get()); });
		struct ReaderGuard
		{
			ChunkQueue& a;
			std::thread& t;
			~ReaderGuard()
			{
				a.abort();
				if (t.joinable()) t.join();
			}
		} reader_guard{q, reader};

		spg_params ep; // the engine needs trimming parameters to come up; this tool does not use them
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
		spg_ctx* engine = nullptr;
		if (spg_create(&engine, &ep, gpus.data(), (int)gpus.size(), 0, 0, 0) != SPG_OK) throw Exception(std::string("Could not initialize the CUDA engine: ") + spg_last_error(nullptr));
		struct EngineGuard
		{
			spg_ctx* e;
			~EngineGuard() { spg_destroy(e); }
		} engine_guard{engine};

		const int n_slots = 3 * (int)gpus.size();
		spg_fq* fq = nullptr;
		int fq_max_len = 0;
		int64_t fq_text_cap = 0;
		struct InFlight
		{
			int slot;
			std::unique_ptr<TextChunk> a;
		};
		std::deque<InFlight> in_flight;
		int next_slot = 0;
		auto retire = [&]() {
			InFlight f = std::move(in_flight.front());
			in_flight.pop_front();
			spg_fq_output o;
			if (spg_fq_wait(fq, f.slot, &o) != SPG_OK) throw Exception(spg_last_error(engine));
			if (o.error_pair >= 0)
			{
				const FastqEntry e = entryAt(*f.a, o.error_pair);
				if (e.bases.size() >= (size_t)MAXLEN) throw ArgumentException("Read length unsupported! A maximum read length of " + std::to_string(MAXLEN) + " is supported (no -long_read)!");
				throw FileParseException("Differing length of bases and qualities string in sequence '" + e.header + "'.");
			}
			if (o.n_pairs != f.a->records || (size_t)o.consumed1 != f.a->data.size()) throw ProgrammingException("the device framed the chunk differently from the reader");
			if (o.out_bytes[0] > 0) writer.write(o.out[0], (size_t)o.out_bytes[0]);
		};
		for (;;)
		{
			std::unique_ptr<TextChunk> a = q.pop();
			if (!a)
			{
				if (q.failure()) std::rethrow_exception(q.failure());
				break;
			}
			if (a->records == 0) continue;
			const int need_len = std::min(a->max_read_len, MAXLEN - 1);
			const int64_t need_text = (int64_t)a->data.size();
			if (!fq || need_len > fq_max_len || need_text > fq_text_cap)
			{
				while (!in_flight.empty()) retire();
				if (fq) spg_fq_close(fq);
				fq = nullptr;
				spg_fq_config cfg;
				memset(&cfg, 0, sizeof(cfg));
				cfg.n_slots = n_slots;
				cfg.max_pairs = block_size;
				cfg.max_len = std::min(MAXLEN - 1, std::max(std::max((need_len + 15) / 16 * 16, fq_max_len), 160));
				cfg.text_cap = std::max<int64_t>(std::max<int64_t>(need_text + need_text / 4, fq_text_cap), 1 << 20);
				cfg.single_end = 1;
				cfg.fixed_trim = 1;
				cfg.trim_start = std::min(start, MAXLEN - 1);
				cfg.trim_end = end;
				cfg.trim_len = len;
				cfg.trim_max_len = max_len;
				if (spg_fq_open(engine, &cfg, &fq) != SPG_OK) throw Exception(std::string("Could not open the FASTQ stream on the device: ") + spg_last_error(engine));
				fq_max_len = cfg.max_len;
				fq_text_cap = cfg.text_cap;
				next_slot = 0;
			}
			if ((int)in_flight.size() == n_slots) retire();
			const int slot = next_slot;
			next_slot = (next_slot + 1) % n_slots;
			spg_fq_input buf;
			if (spg_fq_buffers(fq, slot, &buf) != SPG_OK) throw Exception(spg_last_error(engine));
			memcpy(buf.text1, a->data.data(), a->data.size());
			if (spg_fq_submit(fq, slot, (int64_t)a->data.size(), 0, 1, 1) != SPG_OK) throw Exception(spg_last_error(engine));
			in_flight.push_back(InFlight{slot, std::move(a)});
		}
		while (!in_flight.empty()) retire();
		if (fq) spg_fq_close(fq);
		writer.close();
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
