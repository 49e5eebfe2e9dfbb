/*
 * seqpurge_oracle.c -- CPU restatement of the SeqPurge per-read-pair trimming path (TEST INFRASTRUCTURE ONLY,
 * see seqpurge_oracle.h for the parity status). Plain C11, byte-at-a-time like the reference; no attempt at speed
 * beyond what the reference itself does (it keeps the reference's early abort of the inner loop, so that the
 * CPU baseline timing is fair).
 *
 * Citations are file:line in the reference checkout (imgag/ngs-bits).
 */
#include "seqpurge_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------------------------------------------------------
 * BasicStatistics::precalculateFactorials / factorial / matchProbability  (src/cppCORE/BasicStatistics.cpp:249-307)
 * ------------------------------------------------------------------------------------------------------------- */

static double g_fact[256];
static int g_fact_count = 0;
static pthread_once_t g_fact_once = PTHREAD_ONCE_INIT;

static int valid_float(double v) /* BasicStatistics::isValidFloat(double): finite */
{
	return isfinite(v);
}

static void fact_init(void) /* BasicStatistics.cpp:249-262: multiply up until the double overflows */
{
	int i = 0;
	double value = 1.0;
	while (valid_float(value))
	{
		g_fact[g_fact_count++] = value;
		++i;
		value *= i;
	}
}

double spo_factorial(int n) /* BasicStatistics.cpp:264-279 */
{
	pthread_once(&g_fact_once, fact_init);
	if (n < 0) return NAN; /* the reference throws; never reached on this path */
	if (g_fact_count < n + 1) return NAN;
	return g_fact[n];
}

double spo_match_probability(double p, int n, int count) /* BasicStatistics.cpp:281-307 */
{
	int mismatches = count - n;
	while (!valid_float(spo_factorial(count))) /* :284-290 halve until count! fits a double */
	{
		n /= 2;
		mismatches /= 2;
		count = n + mismatches;
	}

	double output = 0.0;
	for (int i = n; i <= count; ++i) /* :293-298, same left-to-right evaluation order */
	{
		double q = pow(1.0 - p, (double)(count - i)) * pow(p, (double)i) * spo_factorial(count) / spo_factorial(i) / spo_factorial(count - i);
		output += q;
	}
	return output; /* :301-304 throws if not finite -- cannot happen for p=0.25 */
}

/* ---------------------------------------------------------------------------------------------------------------
 * FastqEntry::quality / trimQuality / trimN  (src/cppNGS/FastqFileStream.h:23-26, FastqFileStream.cpp:52-117)
 * ------------------------------------------------------------------------------------------------------------- */

static inline int quality(const char* quals, int i, int offset)
{
	return (int)(signed char)quals[i] - offset; /* QByteArray holds (signed) char */
}

int spo_trim_quality(const char* quals, int* len, int cutoff, int window, int offset)
{
	int count = *len;
	if (count < window) return 0; /* :56 */

	double sum = 0; /* :59-63 */
	for (int i = count - 1; i > count - window; --i) sum += quality(quals, i, offset);

	for (int i = count - window; i >= 0; --i) /* :66-81 */
	{
		sum += quality(quals, i, offset);
		if (sum / window >= cutoff)
		{
			int count_new = i + window;
			while (count_new > 0 && quality(quals, count_new - 1, offset) < cutoff) --count_new;
			*len = count_new;
			return count - count_new;
		}
		sum -= quality(quals, i + window - 1, offset);
	}

	*len = 0; /* :84-86 no window reaches the cutoff */
	return count;
}

int spo_trim_n(const char* bases, int* len, int num_n)
{
	int count = *len;
	if (count < num_n) return 0; /* :93 */

	int sum = 0; /* :96-100 */
	for (int i = 0; i < num_n - 1; ++i) sum += (bases[i] == 'N');

	for (int i = num_n - 1; i < count; ++i) /* :103-114 */
	{
		sum += (bases[i] == 'N');
		if (sum == num_n)
		{
			int count_new = i - num_n + 1;
			*len = count_new;
			return count - count_new;
		}
		sum -= (bases[i - num_n + 1] == 'N');
	}
	return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Sequence::complement / toReverseComplement  (src/cppNGS/Sequence.cpp:41-112)
 * ------------------------------------------------------------------------------------------------------------- */

static inline int complement_base(char b) /* returns 0 for a byte the reference throws on */
{
	switch (b)
	{
		case 'A': return 'T';
		case 'C': return 'G';
		case 'T': return 'A';
		case 'G': return 'C';
		case 'N': return 'N';
		default: return 0;
	}
}

int spo_revcomp(const char* in, int len, char* out)
{
	for (int i = 0; i < len; ++i)
	{
		int c = complement_base(in[len - 1 - i]);
		if (c == 0) return -1;
		out[i] = (char)c;
	}
	return 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * AnalysisWorker::correctErrors  (src/SeqPurge/AnalysisWorker.cpp:19-77)
 * returns 0, or -1 where Sequence::complement would throw on a read-1 byte
 * ------------------------------------------------------------------------------------------------------------- */

static int correct_errors(const spo_params* p, char* r1, char* q1, int n1, char* r2, char* q2, int n2, spo_ecstats* ec)
{
	int mm_count = 0;
	const int count = n1 < n2 ? n1 : n2; /* :22 */
	for (int i = 0; i < count; ++i)
	{
		const int i2 = count - i - 1;
		if ((int)r1[i] != complement_base(r2[i2])) /* :28 (read 2 is ACGTN here: revcomp succeeded) */
		{
			++mm_count;
			int qa = quality(q1, i, p->qoff);
			int qb = quality(q2, i2, p->qoff);
			if (qa > qb) /* :48-58 */
			{
				int rep = complement_base(r1[i]);
				if (rep == 0) return -1;
				r2[i2] = (char)rep;
				q2[i2] = q1[i];
				if (ec) ++ec->mismatch_r2[i2];
			}
			else if (qa < qb) /* :59-69 */
			{
				r1[i] = (char)complement_base(r2[i2]);
				q1[i] = q2[i2];
				if (ec) ++ec->mismatch_r1[i];
			}
		}
	}
	if (mm_count > 0 && ec) ++ec->errors_per_read[mm_count]; /* :73-76 */
	return 0;
}

/* three-way base comparison used everywhere on this path (AnalysisWorker.cpp:155-168 etc.) */
#define CMP3(b1, b2, matches, mismatches, invalid) \
	do                                               \
	{                                                \
		if ((b1) == 'N' || (b2) == 'N') ++(invalid);  \
		else if ((b1) == (b2)) ++(matches);           \
		else ++(mismatches);                          \
	} while (0)

/* first offset of `seq` (length len) at which the adapter matches (steps 2/3, AnalysisWorker.cpp:307-353, :355-407) */
static int adapter_scan(const spo_params* p, const char* seq, int len, const char* adapter)
{
	for (int offset = 0; offset < len; ++offset)
	{
		int matches = 0, mismatches = 0, invalid = 0;
		for (int i = 0; i < p->a_size; ++i)
		{
			if (offset + i >= len) break;
			char b1 = seq[offset + i];
			char b2 = adapter[i];
			CMP3(b1, b2, matches, mismatches, invalid);
		}
		(void)invalid;
		if (100.0 * matches / (matches + mismatches) < p->match_perc) continue; /* 0/0 -> NaN -> not skipped here */
		double prob = spo_match_probability(0.25, matches, matches + mismatches);
		if (prob > p->mep) continue;
		return offset;
	}
	return -1;
}

/* ---------------------------------------------------------------------------------------------------------------
 * AnalysisWorker::run, body of the per-pair loop  (src/SeqPurge/AnalysisWorker.cpp:122-441)
 * ------------------------------------------------------------------------------------------------------------- */

void spo_process_pair(const spo_params* p, char* r1, char* q1, int len1, char* r2, char* q2, int len2, spo_record* out, spo_ecstats* ec)
{
	char seq2[SPO_MAXLEN];
	memset(out, 0, sizeof(*out));
	out->best_offset = -1;

	/* :123-134.  The reference builds revcomp(R2) first (may throw) and checks the length afterwards. */
	if (len2 < SPO_MAXLEN)
	{
		if (spo_revcomp(r2, len2, seq2) != 0)
		{
			out->status = SPO_E_BASE_R2;
			return;
		}
	}
	else
	{
		for (int i = 0; i < len2; ++i)
		{
			if (complement_base(r2[i]) == 0)
			{
				out->status = SPO_E_BASE_R2;
				return;
			}
		}
	}
	const char* seq1 = r1;
	const int min_length = len1 < len2 ? len1 : len2;
	const int max_length = len1 > len2 ? len1 : len2;
	if (max_length >= SPO_MAXLEN)
	{
		out->status = SPO_E_MAXLEN;
		return;
	}

	int n1 = len1; /* current lengths of the two entries (bases and qualities move together) */
	int n2 = len2;
	unsigned flags = 0;

	/* step 1: trim by insert match (:137-266) */
	int best_offset = -1;
	double best_p = 1.0;
	for (int offset = 1; offset < min_length; ++offset)
	{
		int max_mismatches = (int)(ceil((1.0 - p->match_perc / 100.0) * (min_length - offset))); /* :146 */

		int matches = 0, mismatches = 0, invalid = 0;
		for (int j = offset; j < min_length; ++j)
		{
			char b1 = seq1[j - offset];
			char b2 = seq2[j];
			if (b1 == 'N' || b2 == 'N') ++invalid;
			else if (b1 == b2) ++matches;
			else
			{
				++mismatches;
				if (mismatches > max_mismatches) break; /* :166 */
			}
		}
		(void)invalid;

		if ((matches + mismatches) == 0 || 100.0 * matches / (matches + mismatches) < p->match_perc) continue; /* :170 */

		double prob = spo_match_probability(0.25, matches, matches + mismatches); /* :178-179 */
		if (prob > p->mep) continue;

		/* adapter presence on at least one side (:182-259) */
		int a1_m = 0, a1_mm = 0, a1_inv = 0;
		{
			int pos = len2 - offset; /* seq1.mid(len2-offset, adapter_overlap) */
			int alen = 0;
			if (pos < len1)
			{
				alen = len1 - pos;
				if (alen > p->adapter_overlap) alen = p->adapter_overlap;
			}
			for (int i = 0; i < alen; ++i)
			{
				char b1 = seq1[pos + i];
				char b2 = p->a1[i];
				CMP3(b1, b2, a1_m, a1_mm, a1_inv);
			}
		}
		int a2_m = 0, a2_mm = 0, a2_inv = 0;
		{
			/* seq2.left(offset).toReverseComplement().left(adapter_overlap) == R2[len2-offset .. ) */
			int alen = offset < p->adapter_overlap ? offset : p->adapter_overlap;
			for (int i = 0; i < alen; ++i)
			{
				char b1 = r2[len2 - offset + i];
				char b2 = p->a2[i];
				CMP3(b1, b2, a2_m, a2_mm, a2_inv);
			}
		}
		(void)a1_inv;
		(void)a2_inv;

		if (offset < 10) /* :231-245 */
		{
			int max_mm = 2;
			if (offset < 6) max_mm = 1;
			if (offset < 3) max_mm = 0;
			if (!(a1_mm <= max_mm || a2_mm <= max_mm)) continue;
		}
		else /* :246-259 */
		{
			double p1 = spo_match_probability(0.25, a1_m, a1_m + a1_mm);
			double p2 = spo_match_probability(0.25, a2_m, a2_m + a2_mm);
			if (p1 * p2 > p->mep) continue;
		}

		if (prob < best_p) /* :261-265 */
		{
			best_p = prob;
			best_offset = offset;
		}
	}

	if (best_offset != -1) /* :269-302 */
	{
		int new_length = len2 - best_offset;
		if (new_length < n1) n1 = new_length; /* QByteArray::truncate never extends */
		if (new_length < n2) n2 = new_length;
		flags |= SPO_F_INSERT;
		if (p->ec)
		{
			if (correct_errors(p, r1, q1, n1, r2, q2, n2, ec) != 0)
			{
				memset(out, 0, sizeof(*out));
				out->best_offset = -1;
				out->status = 3;
				return;
			}
		}
	}
	else /* steps 2+3 (:304-427) */
	{
		int offset_forward = adapter_scan(p, seq1, len1, p->a1);
		if (offset_forward != -1) n1 = offset_forward;
		int offset_reverse = adapter_scan(p, r2, len2, p->a2);
		if (offset_reverse != -1) n2 = offset_reverse;

		if (offset_forward != -1 || offset_reverse != -1)
		{
			flags |= SPO_F_ADAPTER;
			if (offset_forward == -1 && offset_reverse < n1) n1 = offset_reverse;
			if (offset_reverse == -1 && offset_forward < n2) n2 = offset_forward;
		}
	}

	/* quality trimming (:430-434) */
	if (p->qcut > 0)
	{
		if (spo_trim_quality(q1, &n1, p->qcut, p->qwin, p->qoff) > 0) flags |= SPO_F_Q1;
		if (spo_trim_quality(q2, &n2, p->qcut, p->qwin, p->qoff) > 0) flags |= SPO_F_Q2;
	}

	/* N trimming (:437-441) */
	if (p->ncut > 0)
	{
		if (spo_trim_n(r1, &n1, p->ncut) > 0) flags |= SPO_F_N1;
		if (spo_trim_n(r2, &n2, p->ncut) > 0) flags |= SPO_F_N2;
	}

	out->len1 = (uint16_t)n1;
	out->len2 = (uint16_t)n2;
	out->best_offset = (int16_t)best_offset;
	out->flags = (uint8_t)flags;
	out->status = SPO_OK;
}

void spo_default_params(spo_params* p) /* src/SeqPurge/main.cpp:25-43, :71 */
{
	static const char* A1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA";
	static const char* A2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT";
	p->a1 = A1;
	p->a1_len = 33;
	p->a2 = A2;
	p->a2_len = 33;
	p->a_size = 20;
	p->adapter_overlap = 10;
	p->match_perc = 80.0;
	p->mep = 0.000001;
	p->qcut = 15;
	p->qwin = 5;
	p->qoff = 33;
	p->ncut = 7;
	p->ec = 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * batch form: blocks of pairs over a pthread pool, like the reference's analysis pool over AnalysisJobs
 * ------------------------------------------------------------------------------------------------------------- */

typedef struct batch_ctx
{
	const spo_params* p;
	uint8_t *b1, *q1, *b2, *q2;
	const uint16_t *len1, *len2;
	int stride;
	int64_t n;
	spo_record* out;
	int64_t next; /* next block start, claimed under the mutex */
	int64_t block;
	pthread_mutex_t mu;
	int want_ec;
} batch_ctx;

typedef struct batch_thread
{
	batch_ctx* ctx;
	spo_ecstats* ec; /* per-thread histogram, NULL if not wanted */
} batch_thread;

static void* batch_worker(void* arg)
{
	batch_thread* t = (batch_thread*)arg;
	batch_ctx* c = t->ctx;
	for (;;)
	{
		pthread_mutex_lock(&c->mu);
		int64_t start = c->next;
		c->next += c->block;
		pthread_mutex_unlock(&c->mu);
		if (start >= c->n) break;
		int64_t end = start + c->block < c->n ? start + c->block : c->n;
		for (int64_t r = start; r < end; ++r)
		{
			size_t off = (size_t)r * (size_t)c->stride;
			spo_process_pair(c->p, (char*)c->b1 + off, (char*)c->q1 + off, c->len1[r], (char*)c->b2 + off, (char*)c->q2 + off, c->len2[r], &c->out[r], t->ec);
		}
	}
	return NULL;
}

void spo_trim_batch(const spo_params* p, uint8_t* bases1, uint8_t* quals1, uint8_t* bases2, uint8_t* quals2, const uint16_t* len1, const uint16_t* len2,
                    int stride, int64_t n, spo_record* out, spo_ecstats* ec, int threads)
{
	spo_factorial(0); /* make sure the cache exists before threads start */
	if (threads < 1) threads = 1;
	batch_ctx c;
	c.p = p;
	c.b1 = bases1;
	c.q1 = quals1;
	c.b2 = bases2;
	c.q2 = quals2;
	c.len1 = len1;
	c.len2 = len2;
	c.stride = stride;
	c.n = n;
	c.out = out;
	c.next = 0;
	c.block = 10000; /* default -block_size of the reference (main.cpp:38) */
	if (c.block * threads > n) c.block = (n + threads - 1) / threads;
	if (c.block < 1) c.block = 1;
	c.want_ec = (ec != NULL);
	pthread_mutex_init(&c.mu, NULL);

	batch_thread* ts = (batch_thread*)calloc((size_t)threads, sizeof(batch_thread));
	pthread_t* ids = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
	for (int i = 0; i < threads; ++i)
	{
		ts[i].ctx = &c;
		ts[i].ec = ec ? (spo_ecstats*)calloc(1, sizeof(spo_ecstats)) : NULL;
	}
	if (threads == 1)
	{
		batch_worker(&ts[0]);
	}
	else
	{
		for (int i = 0; i < threads; ++i) pthread_create(&ids[i], NULL, batch_worker, &ts[i]);
		for (int i = 0; i < threads; ++i) pthread_join(ids[i], NULL);
	}
	if (ec)
	{
		for (int i = 0; i < threads; ++i)
		{
			for (int k = 0; k < SPO_MAXLEN; ++k)
			{
				ec->mismatch_r1[k] += ts[i].ec->mismatch_r1[k];
				ec->mismatch_r2[k] += ts[i].ec->mismatch_r2[k];
				ec->errors_per_read[k] += ts[i].ec->errors_per_read[k];
			}
			free(ts[i].ec);
		}
	}
	pthread_mutex_destroy(&c.mu);
	free(ts);
	free(ids);
}

/* ---------------------------------------------------------------------------------------------------------------
 * StatisticsReads::update(const FastqEntry&, ReadDirection)  (src/cppNGS/StatisticsReads.cpp:26-81), raw reads
 * ------------------------------------------------------------------------------------------------------------- */

static void qc_update(const uint8_t* bases, const uint8_t* quals, int cycles, int reverse, spo_qc_stats* s)
{
	if (reverse) ++s->reads_reverse; /* :29-36 */
	else ++s->reads_forward;
	s->bases_sequenced += cycles; /* :39-41 */
	if (cycles < SPO_MAXLEN) s->read_lengths[cycles]++;
	else s->errors++;

	for (int i = 0; i < cycles && i < SPO_MAXLEN; ++i) /* :50, Pileup::inc (src/cppNGS/Pileup.cpp:17-32) */
	{
		switch (bases[i])
		{
			case 'A': case 'a': s->pileup[i][0]++; break;
			case 'C': case 'c': s->pileup[i][1]++; break;
			case 'G': case 'g': s->pileup[i][2]++; break;
			case 'T': case 't': s->pileup[i][3]++; break;
			case 'N': case 'n': s->pileup[i][4]++; break;
			case '-': case '~': break;
			default: s->errors++; /* the reference throws "Unknown base" */
		}
	}

	double q_sum = 0.0; /* :53-70 */
	for (int i = 0; i < cycles && i < SPO_MAXLEN; ++i)
	{
		int q = (int)(signed char)quals[i] - 33; /* FastqEntry::quality(i) with the default offset */
		q_sum += q;
		if (q < 0 || q >= 100) /* >= 100 throws in the reference, < 0 indexes base_qualities_ out of bounds */
		{
			s->errors++;
			continue;
		}
		if (q >= 20.0) ++s->base_q20;
		if (q >= 30.0) ++s->base_q30;
		s->base_qualities[q]++; /* :61 */
		if (reverse) s->qsum_reverse[i] += q;
		else s->qsum_forward[i] += q;
	}
	double mean_qscore = q_sum / cycles; /* :71-80 */
	if (isfinite(mean_qscore))
	{
		long rq = (long)round(mean_qscore); /* :74 read_qualities_[std::round(mean_qscore)]++ */
		if (rq >= 0 && rq < 100) s->read_qualities[rq]++;
		else s->errors++;
		/* :76-77 Histogram(0, 60, 1)::inc(mean, true): index = floor((val-min_) / (max_-min_) * bins_.size()), bounded to the bins
		   (src/cppCORE/Histogram.cpp:117-126) */
		const double hmin = 0.0, hmax = 60.0;
		const long nbins = 60;
		long bi = (long)floor((mean_qscore - hmin) / (hmax - hmin) * nbins);
		if (bi < 0) bi = 0;
		if (bi > nbins - 1) bi = nbins - 1;
		if (reverse) s->qscore_dist_reverse[bi]++;
		else s->qscore_dist_forward[bi]++;
		if (mean_qscore >= 20.0) ++s->read_q20;
	}
}

void spo_qc_update_batch(const uint8_t* bases1, const uint8_t* quals1, const uint8_t* bases2, const uint8_t* quals2, const uint16_t* len1, const uint16_t* len2,
                         int stride, int64_t n, spo_qc_stats* out)
{
	for (int64_t r = 0; r < n; ++r) /* AnalysisWorker.cpp:86-95 */
	{
		size_t off = (size_t)r * (size_t)stride;
		qc_update(bases1 + off, quals1 + off, len1[r], 0, out);
		qc_update(bases2 + off, quals2 + off, len2[r], 1, out);
	}
}

void spo_qc_update_read(const uint8_t* bases, const uint8_t* quals, int len, int reverse, spo_qc_stats* out) { qc_update(bases, quals, len, reverse, out); }

/* FastqEntry::validate (src/cppNGS/FastqFileStream.cpp:3-48), short reads; the checks in the reference's order */
int spo_validate_entry(const char* header, int header_len, const char* bases, int bases_len, const char* header2, int header2_len, const char* quals, int quals_len)
{
	if (header_len == 0 || header[0] != '@') return 1;    /* :7-10 */
	if (header2_len == 0 || header2[0] != '+') return 2;  /* :11-14 */
	if (bases_len != quals_len) return 3;                 /* :15-18 */
	for (int i = 0; i < bases_len; ++i)                   /* :19-25 */
	{
		char c = bases[i];
		if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') return 4;
	}
	for (int i = 0; i < quals_len; ++i) /* :26-45 */
	{
		int value = quals[i];
		if (value < 33 || value > 74) return 5;
	}
	return 0;
}

/* src/FastqTrim/main.cpp:47-77 */
int spo_fastq_trim(int len, int start, int end, int max_bases, int max_len, int* first, int* count)
{
	*first = 0;
	*count = len;
	if (max_len > 0 && len >= max_len) return 1; /* :52-56: written unchanged */
	if (start > 0 || end > 0)                     /* :58-64 */
	{
		if (len <= start + end) return 0;
		*first = start;
		*count = len - start - end;
	}
	if (max_bases > 0 && *count > max_bases) *count = max_bases; /* :66-70 */
	return 1;
}
