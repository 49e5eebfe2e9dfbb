// QcReport.h -- the -qc report of SeqPurge: read statistics of the untrimmed reads as a qcML file.
//
// The accumulators come from the CUDA engine (spg_qc_stats, filled by spg::qc_kernel); this file turns them into the
// quality parameters of StatisticsReads::getResult (src/cppNGS/StatisticsReads.cpp:140-200, paired-end case) and writes them in the
// layout of QCCollection::storeToQCML (src/cppNGS/QCCollection.cpp:200-262): metaDataParameter lines, eight qualityParameter lines,
// three attachment entries. The plots of the reference (PNG images in <binary> elements) and the embedded XSL stylesheet are not
// produced -- the reference's own test ignores the <binary> lines (src/tools-TEST/SeqPurge_Test.cpp:108-112).
#pragma once
#include <string>
#include <utility>
#include <vector>

#include "../../include/seqpurge_b200.h"

namespace seqpurge
{


// name -> value strings as a qcML file holds them (integers as such, doubles with two decimals)
std::vector<std::pair<std::string, std::string>> qcMetrics(const spg_qc_stats& s);

// adds b into a (statistics of several engines / devices)
void qcAccumulate(spg_qc_stats& a, const spg_qc_stats& b);

void storeQcML(const std::string& filename, const spg_qc_stats& stats, const std::vector<std::string>& source_files, const std::string& parameters,
               const std::string& software = "seqpurge_b200");

} // namespace seqpurge
