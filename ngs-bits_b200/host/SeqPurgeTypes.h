// SeqPurgeTypes.h -- Qt-free host-side mirror of the types at the seam the CUDA engine replaces.
//
// Same names and meaning as the reference (imgag/ngs-bits) so that the worker below reads like the reference's:
//   FastqEntry                  src/cppNGS/FastqFileStream.h:11-36
//   AnalysisJob, AnalysisStatus src/SeqPurge/Auxilary.h:15-66
//   TrimmingParameters          src/SeqPurge/Auxilary.h:100-133 (defaults: src/SeqPurge/main.cpp:25-43)
//   TrimmingStatistics          src/SeqPurge/Auxilary.h:136-221
//   ErrorCorrectionStatistics   src/SeqPurge/Auxilary.h:224-271
// QByteArray -> std::string, QVector -> std::vector, THROW(XException, msg) -> throw XException(msg).
#pragma once
#include <cstdint>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace seqpurge
{

// QString::number(v, 'f', 2) as the reference prints percentages and qcML values: two decimals, an exact tie goes up, NaN is "nan"
std::string fixed2(double v);

const int MAXLEN = 1000; // src/SeqPurge/Auxilary.h:12

// exception kinds of cppCORE (src/cppCORE/Exceptions.h:15-174); the tool prints the message and exits with 1
struct Exception : public std::runtime_error
{
	explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
struct ArgumentException : public Exception { using Exception::Exception; };
struct ProgrammingException : public Exception { using Exception::Exception; };
struct FileParseException : public Exception { using Exception::Exception; };
struct FileAccessException : public Exception { using Exception::Exception; };
struct CommandLineParsingException : public Exception { using Exception::Exception; };

struct FastqEntry
{
	std::string header;
	std::string bases;
	std::string header2;
	std::string qualities;
	void clear()
	{
		header.clear();
		bases.clear();
		header2.clear();
		qualities.clear();
	}
};

enum AnalysisStatus
{
	TO_BE_ANALYZED,
	TO_BE_WRITTEN,
	DONE
};

struct AnalysisJob
{
	AnalysisJob(int i, int block_size) : index(i), r1(block_size), r2(block_size), length_r1_orig(block_size), length_r2_orig(block_size) { clear(); }
	int index;
	std::vector<FastqEntry> r1;
	std::vector<FastqEntry> r2;
	int read_count;
	AnalysisStatus status;
	std::vector<int> length_r1_orig;
	std::vector<int> length_r2_orig;
	int reads_trimmed_insert;
	int reads_trimmed_adapter;
	int reads_trimmed_q;
	int reads_trimmed_n;
	void clear()
	{
		read_count = -1;
		status = DONE;
		reads_trimmed_insert = reads_trimmed_adapter = reads_trimmed_q = reads_trimmed_n = 0;
	}
};

struct TrimmingParameters
{
	std::vector<std::string> files_in1, files_in2;
	std::string out1, out2, out3, summary;
	int adapter_overlap = 10;
	std::string a1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA";
	std::string a2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT";
	int a_size = 20;
	double match_perc = 80.0;
	double mep = 0.000001;
	int min_len = 30;
	int block_prefetch = 32;
	int block_size = 10000;
	int threads = 1;
	int progress = -1;
	int qcut = 15;
	int qwin = 5;
	int qoff = 33;
	int ncut = 7;
	bool ec = false;
	bool debug = false;
	int compression_level = 1; // Z_BEST_SPEED
	std::string qc;
	std::vector<int> gpus = {0}; // new, behaviour-neutral: CUDA devices the blocks are dealt to round robin
	bool bgzf = false;           // new: outputs are written as BGZF (blocked gzip); stream pipeline only
	bool host_framing = false;   // new: FASTQ records are parsed/formatted on the host (block pipeline of the reference) instead of on the device
};

// counts of one pileup column, the part of cppNGS Pileup the consensus adapter needs (src/cppNGS/Pileup.cpp:17-32)
struct BaseCounts
{
	long long a = 0, c = 0, g = 0, t = 0, n = 0;
	void inc(char base); // throws ArgumentException on an unknown base like Pileup::inc
	long long depth() const { return a + c + g + t; }
	long long max() const;
};

struct TrimmingStatistics
{
	TrimmingStatistics() : bases_remaining(MAXLEN, 0.0), acons1(40), acons2(40) {}
	long long read_num = 0;
	std::vector<double> bases_remaining;
	std::vector<BaseCounts> acons1, acons2;
	double reads_trimmed_insert = 0, reads_trimmed_adapter = 0, reads_trimmed_q = 0, reads_trimmed_n = 0, reads_removed = 0, bases_perc_trim_sum = 0;
	void writeStatistics(std::ostream& out, const TrimmingParameters& params) const;
};

struct ErrorCorrectionStatistics
{
	ErrorCorrectionStatistics() : mismatch_r1(MAXLEN, 0), mismatch_r2(MAXLEN, 0), errors_per_read(MAXLEN, 0) {}
	std::vector<long long> mismatch_r1, mismatch_r2, errors_per_read;
	void writeStatistics(std::ostream& out) const;
};

} // namespace seqpurge
