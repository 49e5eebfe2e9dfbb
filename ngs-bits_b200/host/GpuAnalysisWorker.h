// GpuAnalysisWorker.h -- the replacement of the reference's AnalysisWorker (src/SeqPurge/AnalysisWorker.{h,cpp}).
//
// Same constructor arguments and the same effect on the job as AnalysisWorker::run(): after wait() the entries of the job are
// trimmed (lengths) and, with -ec, corrected; the job counters, length_r*_orig and the adapter-consensus counters are filled.
// The per-pair computation itself runs in libseqpurge_b200.so (include/seqpurge_b200.h); this class only moves a job into its
// pinned slot (start) and applies the result records (wait). Splitting run() in two lets the coordinator keep one job per slot in
// flight on every GPU and still retire jobs in input order (the order of the reference's `-threads 1` output).
#pragma once
#include "../../include/seqpurge_b200.h"
#include "SeqPurgeTypes.h"

namespace seqpurge
{

class GpuAnalysisWorker
{
public:
	GpuAnalysisWorker(AnalysisJob& job, const TrimmingParameters& params, TrimmingStatistics& stats, ErrorCorrectionStatistics& ecstats, spg_ctx* engine, int slot);

	void start(); // header check, length check, AoS -> pinned SoA slot, spg_submit
	void wait();  // spg_wait, apply records to the job; throws what AnalysisWorker::run would have emitted as error()
	void run()
	{
		start();
		wait();
	}

	// trimmed lengths of read 1 / read 2 of pair r (valid after wait); the entries themselves keep their full strings
	int length1(int r) const { return len1_[(size_t)r]; }
	int length2(int r) const { return len2_[(size_t)r]; }

private:
	AnalysisJob& job_;
	const TrimmingParameters& params_;
	TrimmingStatistics& stats_;
	ErrorCorrectionStatistics& ecstats_;
	spg_ctx* engine_;
	int slot_;
	std::vector<int> len1_, len2_;
};

// fills spg_params from TrimmingParameters (the strings must outlive the call to spg_create only)
spg_params toEngineParams(const TrimmingParameters& p);

} // namespace seqpurge
