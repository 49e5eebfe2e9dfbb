#include "SeqPurgeTypes.h"

#include <algorithm>
#include <cstdio>

namespace seqpurge
{

// QCValue::toString -> QString::number(v, 'f', 2). Qt formats through the double-conversion library, whose fixed notation resolves
// an exact tie upwards (like ECMAScript's toFixed: 0.125 -> "0.13"), where printf resolves it to even ("0.12"). The reference's golden
// ReadQC_out4.qcML holds such a value (125 000 bases = 0.125 MB), so the third decimal of the exact expansion decides here.
std::string fixed2(double v)
{
	if (v != v) return "nan"; // QString::number prints "nan" (glibc would print "-nan" for 0/0)
	if (!(v >= 0.0) || v > 1e15) // negative or huge: not produced by these statistics
	{
		char b[64];
		snprintf(b, sizeof(b), "%.2f", v);
		return b;
	}
	char buf[128];
	snprintf(buf, sizeof(buf), "%.40f", v); // glibc prints the exact binary value; 40 decimals are far beyond the third one
	std::string s(buf);
	const size_t dot = s.find('.');
	const bool up = s[dot + 3] >= '5';
	long long cents = 0; // integer part and two decimals as one number
	for (size_t i = 0; i < dot + 3; ++i)
		if (s[i] != '.') cents = cents * 10 + (s[i] - '0');
	if (up) ++cents;
	char out[64];
	snprintf(out, sizeof(out), "%lld.%02lld", cents / 100, cents % 100);
	return out;
}


void BaseCounts::inc(char base)
{
	switch (base)
	{
		case 'A': case 'a': ++a; break;
		case 'C': case 'c': ++c; break;
		case 'G': case 'g': ++g; break;
		case 'T': case 't': ++t; break;
		case 'N': case 'n': ++n; break;
		case '-': case '~': break;
		default: throw ArgumentException(std::string("Unknown base '") + base + "' in pileup!");
	}
}

long long BaseCounts::max() const { return std::max(std::max(a, c), std::max(g, t)); }

namespace
{
std::string right4(long long v)
{
	char buf[32];
	snprintf(buf, sizeof(buf), "%4lld", v);
	return buf;
}
std::string consensus(const std::vector<BaseCounts>& ac)
{
	std::string seq;
	for (size_t i = 0; i < ac.size(); ++i)
	{
		long long depth = ac[i].depth();
		if (depth < 20) break;
		long long mx = ac[i].max();
		if ((double)mx / depth <= 0.5) seq += 'N';
		else if (ac[i].a == mx) seq += 'A';
		else if (ac[i].c == mx) seq += 'C';
		else if (ac[i].g == mx) seq += 'G';
		else if (ac[i].t == mx) seq += 'T';
		if (i == 39) break;
	}
	return seq;
}
void histogram(std::ostream& out, const std::vector<long long>& v)
{
	int max = (int)v.size() - 1;
	while (max > 0 && v[(size_t)max] == 0) --max;
	for (int i = 1; i <= max; ++i) out << right4(i) << ": " << v[(size_t)i] << "\n";
}
} // namespace

void TrimmingStatistics::writeStatistics(std::ostream& out, const TrimmingParameters& params) const
{
	out << "Reads (forward + reverse): " << read_num << "\n\n";
	out << "Reads trimmed by insert match: " << (long long)reads_trimmed_insert << "\n";
	out << "Reads trimmed by adapter match: " << (long long)reads_trimmed_adapter << "\n";
	out << "Reads trimmed by quality: " << (long long)reads_trimmed_q << "\n";
	out << "Reads trimmed by N stretches: " << (long long)reads_trimmed_n << "\n";
	double reads_trimmed = reads_trimmed_insert + reads_trimmed_adapter;
	out << "Trimmed reads: " << (long long)reads_trimmed << " of " << read_num << " (" << fixed2(100.0 * reads_trimmed / read_num) << "%)\n";
	out << "Removed reads: " << (long long)reads_removed << " of " << read_num << " (" << fixed2(100.0 * reads_removed / read_num) << "%)\n";
	out << "Removed bases: " << fixed2(100.0 * bases_perc_trim_sum / read_num) << "%\n\n";
	out << "Forward adapter sequence (given)    : " << params.a1 << "\n";
	out << "Forward adapter sequence (consensus): " << consensus(acons1) << "\n";
	out << "Reverse adapter sequence (given)    : " << params.a2 << "\n";
	out << "Reverse adapter sequence (consensus): " << consensus(acons2) << "\n\n";
	out << "Read length distribution after trimming:\n";
	int max = (int)bases_remaining.size() - 1;
	while (max > 0 && bases_remaining[(size_t)max] == 0) --max;
	for (int i = 0; i <= max; ++i) out << right4(i) << ": " << (long long)bases_remaining[(size_t)i] << "\n";
}

void ErrorCorrectionStatistics::writeStatistics(std::ostream& out) const
{
	out << "\nRead error per cycle (read 1):\n";
	histogram(out, mismatch_r1);
	out << "\nRead error per cycle (read 2):\n";
	histogram(out, mismatch_r2);
	out << "\nRead error count distribution:\n";
	histogram(out, errors_per_read);
}

} // namespace seqpurge
