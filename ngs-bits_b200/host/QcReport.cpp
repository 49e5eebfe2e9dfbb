#include "QcReport.h"

#include <cstdio>
#include <ctime>
#include <fstream>

#include "SeqPurgeTypes.h"

namespace seqpurge
{

namespace
{
std::string id4(const char* prefix, int i)
{
	char buf[32];
	snprintf(buf, sizeof(buf), "%s%04d", prefix, i);
	return buf;
}
std::string htmlEscaped(const std::string& s)
{
	std::string o;
	for (char c : s)
	{
		if (c == '&') o += "&amp;";
		else if (c == '<') o += "&lt;";
		else if (c == '>') o += "&gt;";
		else if (c == '"') o += "&quot;";
		else o += c;
	}
	return o;
}
std::string fileName(const std::string& path)
{
	size_t p = path.find_last_of('/');
	return p == std::string::npos ? path : path.substr(p + 1);
}

struct ParamInfo
{
	const char* name;
	const char* description;
	const char* accession;
};
// names, descriptions and ontology accessions of the terms (qcML.obo) StatisticsReads::getResult reports
const ParamInfo kValues[8] = {
    {"read count", "Total number of reads (forward and reverse reads of paired-end sequencing count as two reads).", "QC:2000005"},
    {"read length", "Raw read length of a single read before trimming. Comma-separated list of lenghs or length range, if reads have different lengths.", "QC:2000006"},
    {"bases sequenced (MB)", "Bases sequenced in total (in megabases).", "QC:2000049"},
    {"Q20 read percentage", "The percentage of reads with a mean base quality score greater than Q20.", "QC:2000007"},
    {"Q20 base percentage", "The percentage of bases with a minimum quality score of Q20.", "QC:2000148"},
    {"Q30 base percentage", "The percentage of bases with a minimum quality score of Q30.", "QC:2000008"},
    {"no base call percentage", "The percentage of bases without base call (N).", "QC:2000009"},
    {"gc content percentage", "The percentage of bases that are called to be G or C.", "QC:2000010"},
};
const ParamInfo kPlots[3] = {
    {"base distribution plot", "Base distribution plot per cycle.", "QC:2000011"},
    {"Q score plot", "Mean Q score per cycle for forward/reverse reads.", "QC:2000012"},
    {"read Q score distribution", "Distrubition of the mean forward/reverse Q score for each read.", "QC:2000138"},
};
} // namespace

std::vector<std::pair<std::string, std::string>> qcMetrics(const spg_qc_stats& s)
{
	if (s.errors != 0) throw ArgumentException("Unknown base or base quality outside 0..99 in the input reads (read QC)!"); // Pileup::inc / StatisticsReads.cpp:59
	const long long total_reads = s.reads_forward + s.reads_reverse;
	long long c_base_n = 0, c_base_gc = 0, bases_total = 0;
	for (int i = 0; i < SPG_MAXLEN; ++i)
	{
		c_base_n += s.pileup[i][4];
		c_base_gc += s.pileup[i][1] + s.pileup[i][2];
		for (int k = 0; k < 5; ++k) bases_total += s.pileup[i][k];
	}
	std::vector<int> keys;
	for (int i = 0; i < SPG_MAXLEN; ++i)
		if (s.read_lengths[i] > 0) keys.push_back(i);
	std::string lengths;
	if (keys.empty()) lengths = "";
	else if (keys.size() < 4)
	{
		lengths = std::to_string(keys[0]);
		for (size_t i = 1; i < keys.size(); ++i) lengths += ", " + std::to_string(keys[i]);
	}
	else lengths = std::to_string(keys.front()) + "-" + std::to_string(keys.back());

	std::vector<std::pair<std::string, std::string>> out;
	out.emplace_back(kValues[0].name, std::to_string(total_reads));
	out.emplace_back(kValues[1].name, lengths);
	out.emplace_back(kValues[2].name, fixed2((double)s.bases_sequenced / 1000000.0));
	out.emplace_back(kValues[3].name, fixed2(100.0 * s.read_q20 / total_reads));
	out.emplace_back(kValues[4].name, fixed2(100.0 * s.base_q20 / bases_total));
	out.emplace_back(kValues[5].name, fixed2(100.0 * s.base_q30 / bases_total));
	out.emplace_back(kValues[6].name, fixed2(100.0 * c_base_n / bases_total));
	out.emplace_back(kValues[7].name, fixed2(100.0 * c_base_gc / (bases_total - c_base_n)));
	return out;
}

void qcAccumulate(spg_qc_stats& a, const spg_qc_stats& b)
{
	a.reads_forward += b.reads_forward;
	a.reads_reverse += b.reads_reverse;
	a.bases_sequenced += b.bases_sequenced;
	a.read_q20 += b.read_q20;
	a.base_q20 += b.base_q20;
	a.base_q30 += b.base_q30;
	a.errors += b.errors;
	for (int i = 0; i < SPG_MAXLEN; ++i)
	{
		a.read_lengths[i] += b.read_lengths[i];
		for (int k = 0; k < 5; ++k) a.pileup[i][k] += b.pileup[i][k];
		a.qsum_forward[i] += b.qsum_forward[i];
		a.qsum_reverse[i] += b.qsum_reverse[i];
	}
	for (int i = 0; i < 100; ++i)
	{
		a.base_qualities[i] += b.base_qualities[i];
		a.read_qualities[i] += b.read_qualities[i];
	}
	for (int i = 0; i < 60; ++i)
	{
		a.qscore_dist_forward[i] += b.qscore_dist_forward[i];
		a.qscore_dist_reverse[i] += b.qscore_dist_reverse[i];
	}
}

void storeQcML(const std::string& filename, const spg_qc_stats& stats, const std::vector<std::string>& source_files, const std::string& parameters, const std::string& software)
{
	const auto values = qcMetrics(stats);
	std::ofstream out(filename);
	if (!out) throw FileAccessException("Could not open file '" + filename + "' for writing!");
	char date[64];
	time_t now = time(nullptr);
	strftime(date, sizeof(date), "%Y-%m-%dT%H:%M:%S", localtime(&now));

	out << "<?xml version=\"1.0\" encoding=\"ISO-8859-1\"?>\n";
	out << "<qcML version=\"0.0.8\" xmlns=\"http://www.prime-xs.eu/ms/qcml\" >\n";
	out << "  <runQuality ID=\"rq0001\">\n";
	out << "    <metaDataParameter ID=\"md0001\" name=\"creation software\" value=\"" << htmlEscaped(software) << "\" cvRef=\"QC\" accession=\"QC:1000002\"/>\n";
	out << "    <metaDataParameter ID=\"md0002\" name=\"creation software parameters\" value=\"" << htmlEscaped(parameters) << "\" cvRef=\"QC\" accession=\"QC:1000003\"/>\n";
	out << "    <metaDataParameter ID=\"md0003\" name=\"creation date\" value=\"" << date << "\" cvRef=\"QC\" accession=\"QC:1000004\"/>\n";
	int idx = 4;
	for (const std::string& sf : source_files)
	{
		out << "    <metaDataParameter ID=\"" << id4("md", idx) << "\" name=\"source file\" value=\"" << htmlEscaped(fileName(sf)) << "\" cvRef=\"QC\" accession=\"QC:1000005\"/>\n";
		++idx;
	}
	for (int i = 0; i < 8; ++i)
	{
		out << "    <qualityParameter ID=\"" << id4("qp", i + 1) << "\" name=\"" << kValues[i].name << "\" description=\"" << htmlEscaped(kValues[i].description) << "\" value=\""
		    << values[(size_t)i].second << "\" cvRef=\"QC\" accession=\"" << kValues[i].accession << "\"/>\n";
	}
	for (int i = 0; i < 3; ++i) // the plots themselves (PNG in a <binary> element) are not rendered here
	{
		out << "    <attachment ID=\"" << id4("qp", 9 + i) << "\" name=\"" << kPlots[i].name << "\" description=\"" << htmlEscaped(kPlots[i].description) << "\" cvRef=\"QC\" accession=\""
		    << kPlots[i].accession << "\">\n";
		out << "    </attachment>\n";
	}
	out << "  </runQuality>\n";
	out << "  <cvList>\n";
	out << "    <cv uri=\"https://raw.githubusercontent.com/imgag/ngs-bits/master/src/cppNGS/Resources/qcML.obo\" ID=\"QC\" fullName=\"QC\" version=\"0.1\"/>\n";
	out << "  </cvList>\n";
	out << "</qcML>\n";
}

} // namespace seqpurge
