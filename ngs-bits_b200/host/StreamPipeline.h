// StreamPipeline.h -- default pipeline of the command line: inflate on the host, everything between the inflated input text and
// the output text on the GPUs (spg_fq_* of the C ABI), deflate on the host. See StreamPipeline.cpp.
#pragma once
#include <ostream>

#include "../../include/seqpurge_b200.h"
#include "SeqPurgeTypes.h"

namespace seqpurge
{

// Runs the whole job described by params (inputs, outputs, trimming parameters, -gpus, -threads = deflate threads).
// Fills the statistics the summary prints; qc_stats (may be null) receives the -qc accumulators.
void runStreamPipeline(const TrimmingParameters& params, std::ostream& summary, TrimmingStatistics& stats, ErrorCorrectionStatistics& ec_stats, spg_qc_stats* qc_stats);

} // namespace seqpurge
