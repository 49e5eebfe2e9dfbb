#include "GpuAnalysisWorker.h"

#include <algorithm>
#include <cstring>

namespace seqpurge
{

spg_params toEngineParams(const TrimmingParameters& p)
{
	spg_params e;
	e.a1 = p.a1.data();
	e.a1_len = (int32_t)p.a1.size();
	e.a2 = p.a2.data();
	e.a2_len = (int32_t)p.a2.size();
	e.adapter_overlap = p.adapter_overlap;
	e.match_perc = p.match_perc;
	e.mep = p.mep;
	e.qcut = p.qcut;
	e.qwin = p.qwin;
	e.qoff = p.qoff;
	e.ncut = p.ncut;
	e.ec = p.ec ? 1 : 0;
	e.qc = p.qc.empty() ? 0 : 1;
	return e;
}

GpuAnalysisWorker::GpuAnalysisWorker(AnalysisJob& job, const TrimmingParameters& params, TrimmingStatistics& stats, ErrorCorrectionStatistics& ecstats, spg_ctx* engine, int slot)
    : job_(job), params_(params), stats_(stats), ecstats_(ecstats), engine_(engine), slot_(slot)
{
}

namespace
{
// first space-delimited token of a header (QByteArray::split(' ').at(0))
std::string firstToken(const std::string& h)
{
	size_t p = h.find(' ');
	return p == std::string::npos ? h : h.substr(0, p);
}
bool endsWith(const std::string& s, const char* suffix)
{
	size_t n = strlen(suffix);
	return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}
bool isAcgtn(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'N'; }
void throwBadComplement(const std::string& bases)
{
	for (char c : bases)
		if (!isAcgtn(c)) throw ProgrammingException(std::string("Could not convert base '") + c + "' to complement!"); // Sequence.cpp:68
	throw ProgrammingException("Could not convert base to complement!");
}
} // namespace

void GpuAnalysisWorker::start()
{
	spg_slot_view v;
	if (spg_slot_buffers(engine_, slot_, &v) != SPG_OK) throw Exception(spg_last_error(engine_));
	if (job_.read_count > v.max_pairs) throw ProgrammingException("job is larger than the engine slot");
	// the last SPG_QTAIL qualities of every read go into the slot's tail planes as well (SPG_OPT_QUAL_TAILS, set by the command line): they
	// travel with the bases, the quality rows stay in the pinned slot
	uint8_t* qt1 = nullptr;
	uint8_t* qt2 = nullptr;
	int tails = 0;
	if (spg_get_option(engine_, SPG_OPT_QUAL_TAILS, &tails) != SPG_OK) throw Exception(spg_last_error(engine_));
	if (tails && spg_slot_qtails(engine_, slot_, &qt1, &qt2) != SPG_OK) throw Exception(spg_last_error(engine_));
	auto writeTail = [](uint8_t* tail, const std::string& q) {
		const size_t n = std::min(q.size(), (size_t)SPG_QTAIL);
		memcpy(tail + SPG_QTAIL - n, q.data() + q.size() - n, n);
	};

	for (int r = 0; r < job_.read_count; ++r)
	{
		const FastqEntry& e1 = job_.r1[(size_t)r];
		const FastqEntry& e2 = job_.r2[(size_t)r];

		// check that headers match (AnalysisWorker.cpp:110-120)
		std::string tmp1 = firstToken(e1.header);
		std::string tmp2 = firstToken(e2.header);
		if (endsWith(tmp1, "/1") && endsWith(tmp2, "/2"))
		{
			tmp1.resize(tmp1.size() - 2);
			tmp2.resize(tmp2.size() - 2);
		}
		if (tmp1 != tmp2) throw ArgumentException("Headers of reads do not match:\n" + tmp1 + "\n" + tmp2);

		const size_t len1 = e1.bases.size(), len2 = e2.bases.size();
		if (e1.qualities.size() != len1 || e2.qualities.size() != len2)
			throw FileParseException("Differing length of bases and qualities string in sequence '" + e1.header + "'."); // the reference streams do not validate; see DESIGN.md
		// the reference builds revcomp(read 2) first (throws on a byte outside ACGTN) and checks the length afterwards (AnalysisWorker.cpp:123-134)
		if (std::max(len1, len2) >= (size_t)MAXLEN)
		{
			if (!std::all_of(e2.bases.begin(), e2.bases.end(), isAcgtn)) throwBadComplement(e2.bases);
			throw ArgumentException("Read length unsupported! A maximum read length of " + std::to_string(MAXLEN) + " is supported!");
		}
		if (std::max(len1, len2) > (size_t)v.stride) throw ProgrammingException("read longer than the engine's max_len");

		const size_t off = (size_t)r * (size_t)v.stride;
		memcpy(v.bases1 + off, e1.bases.data(), len1);
		memcpy(v.quals1 + off, e1.qualities.data(), len1);
		memcpy(v.bases2 + off, e2.bases.data(), len2);
		memcpy(v.quals2 + off, e2.qualities.data(), len2);
		v.len1[r] = (uint16_t)len1;
		v.len2[r] = (uint16_t)len2;
		if (tails)
		{
			writeTail(qt1 + (size_t)r * SPG_QTAIL, e1.qualities);
			writeTail(qt2 + (size_t)r * SPG_QTAIL, e2.qualities);
		}
		job_.length_r1_orig[(size_t)r] = (int)len1;
		job_.length_r2_orig[(size_t)r] = (int)len2;
	}
	if (spg_submit(engine_, slot_, job_.read_count) != SPG_OK) throw Exception(spg_last_error(engine_));
}

void GpuAnalysisWorker::wait()
{
	const spg_result* res = nullptr;
	if (spg_wait(engine_, slot_, &res) != SPG_OK) throw Exception(spg_last_error(engine_));
	spg_slot_view v;
	if (spg_slot_buffers(engine_, slot_, &v) != SPG_OK) throw Exception(spg_last_error(engine_));
	(void)ecstats_; // the -ec histograms are accumulated inside the engine (spg_ec_stats_get)

	len1_.assign((size_t)job_.read_count, 0);
	len2_.assign((size_t)job_.read_count, 0);
	for (int r = 0; r < job_.read_count; ++r)
	{
		const spg_result& k = res[r];
		FastqEntry& e1 = job_.r1[(size_t)r];
		FastqEntry& e2 = job_.r2[(size_t)r];
		if (k.status == SPG_PAIR_BAD_BASE_R2) throwBadComplement(e2.bases);
		if (k.status == SPG_PAIR_BAD_BASE_EC) throwBadComplement(e1.bases);
		if (k.status == SPG_PAIR_TOO_LONG) throw ArgumentException("Read length unsupported! A maximum read length of " + std::to_string(MAXLEN) + " is supported!");

		if (k.flags & SPG_F_INSERT)
		{
			// update consensus adapter sequence from the untrimmed, uncorrected reads (AnalysisWorker.cpp:279-290)
			const int len2 = job_.length_r2_orig[(size_t)r];
			const size_t new_length = (size_t)(len2 - k.best_offset);
			for (size_t i = 0; i < 40 && new_length + i < e1.bases.size(); ++i) stats_.acons1[i].inc(e1.bases[new_length + i]);
			for (size_t i = 0; i < 40 && i < (size_t)k.best_offset; ++i) stats_.acons2[i].inc(e2.bases[new_length + i]);
			job_.reads_trimmed_insert += 2;
			if (params_.ec) // corrected bases/qualities come back in the slot (AnalysisWorker.cpp:19-77)
			{
				const size_t off = (size_t)r * (size_t)v.stride;
				memcpy(&e1.bases[0], v.bases1 + off, e1.bases.size());
				memcpy(&e1.qualities[0], v.quals1 + off, e1.qualities.size());
				memcpy(&e2.bases[0], v.bases2 + off, e2.bases.size());
				memcpy(&e2.qualities[0], v.quals2 + off, e2.qualities.size());
			}
		}
		if (k.flags & SPG_F_ADAPTER) job_.reads_trimmed_adapter += 2;
		job_.reads_trimmed_q += ((k.flags & SPG_F_Q1) ? 1 : 0) + ((k.flags & SPG_F_Q2) ? 1 : 0);
		job_.reads_trimmed_n += ((k.flags & SPG_F_N1) ? 1 : 0) + ((k.flags & SPG_F_N2) ? 1 : 0);

		e1.bases.resize(k.len1);
		e1.qualities.resize(k.len1);
		e2.bases.resize(k.len2);
		e2.qualities.resize(k.len2);
		len1_[(size_t)r] = k.len1;
		len2_[(size_t)r] = k.len2;
	}
	job_.status = TO_BE_WRITTEN;
}

} // namespace seqpurge
