/*
 * seqpurge_oracle.h -- CPU restatement of the SeqPurge per-read-pair trimming path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and only as the checker / the CPU baseline.  The product path (ngs-bits_b200/) never links,
 * imports or executes this code.
 *
 * Parity status: PINNED.  oracle/seqpurge_oracle_cli reproduces, record for record, all 23 golden
 * output files of the reference's ten single-thread SeqPurge tool tests
 * (src/tools-TEST/SeqPurge_Test.cpp:100-208), plus the unit-test known answers for
 * trimQuality / trimN (src/cppNGS-TEST/FastqFileStream_Test.cpp:9-128) and
 * matchProbability / factorial (src/cppCORE-TEST/BasicStatistics_Test.cpp:144-162);
 * see tests/test_oracle_golden.py.
 *
 * The reference itself (Qt6 + qmake + moc + cppCORE/cppNGS/htslib) cannot be compiled in this
 * image (no Qt headers, no moc), so there is no oracle/_ref; this restatement is the oracle.
 *
 * All file:line citations are relative to the reference checkout (imgag/ngs-bits).
 */
#ifndef SEQPURGE_ORACLE_H
#define SEQPURGE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPO_MAXLEN 1000 /* src/SeqPurge/Auxilary.h:12 */

/* flags of a result record (same bit assignment as include/seqpurge_b200.h) */
#define SPO_F_INSERT 0x01u  /* insert match found        (AnalysisWorker.cpp:269-302) */
#define SPO_F_ADAPTER 0x02u /* adapter-only hit          (AnalysisWorker.cpp:410-426) */
#define SPO_F_Q1 0x04u      /* read 1 shortened by trimQuality (AnalysisWorker.cpp:432) */
#define SPO_F_Q2 0x08u      /* read 2 shortened by trimQuality (AnalysisWorker.cpp:433) */
#define SPO_F_N1 0x10u      /* read 1 shortened by trimN (AnalysisWorker.cpp:439) */
#define SPO_F_N2 0x20u      /* read 2 shortened by trimN (AnalysisWorker.cpp:440) */

/* status of a result record */
#define SPO_OK 0
#define SPO_E_BASE_R2 1 /* byte outside ACGTN in read 2 -> Sequence::complement throws (Sequence.cpp:46-71) */
#define SPO_E_MAXLEN 2  /* max(len1,len2) >= MAXLEN (AnalysisWorker.cpp:131-134) */

/* run constants, mirrors TrimmingParameters (src/SeqPurge/Auxilary.h:100-133, main.cpp:20-43,66-71) */
typedef struct spo_params
{
	const char* a1; /* forward adapter bytes */
	int a1_len;
	const char* a2; /* reverse adapter bytes */
	int a2_len;
	int a_size;          /* min(20, |a1|, |a2|)  (main.cpp:71) */
	int adapter_overlap; /* 10 (Auxilary.h:103) */
	double match_perc;   /* 80.0 */
	double mep;          /* 1e-6 */
	int qcut, qwin, qoff; /* 15, 5, 33 */
	int ncut;            /* 7 */
	int ec;              /* error correction on insert hits */
} spo_params;

/* 8-byte per-pair result record: everything the writer / statistics need */
typedef struct spo_record
{
	uint16_t len1;       /* length of read 1 after all trimming */
	uint16_t len2;       /* length of read 2 after all trimming */
	int16_t best_offset; /* insert-match offset, -1 if none */
	uint8_t flags;       /* SPO_F_* */
	uint8_t status;      /* SPO_OK / SPO_E_* */
} spo_record;

/* error-correction histograms, mirrors ErrorCorrectionStatistics (Auxilary.h:224-271) */
typedef struct spo_ecstats
{
	int64_t mismatch_r1[SPO_MAXLEN];
	int64_t mismatch_r2[SPO_MAXLEN];
	int64_t errors_per_read[SPO_MAXLEN];
} spo_ecstats;

/* accumulators of StatisticsReads::update(const FastqEntry&, ReadDirection) (src/cppNGS/StatisticsReads.cpp:26-81) that the
   paired-end qcML report uses (StatisticsReads::getResult, :140-330); same layout as spg_qc_stats */
typedef struct spo_qc_stats
{
	int64_t reads_forward, reads_reverse, bases_sequenced, read_q20, base_q20, base_q30, errors;
	int64_t read_lengths[SPO_MAXLEN];
	int64_t pileup[SPO_MAXLEN][5]; /* A C G T N */
	int64_t qsum_forward[SPO_MAXLEN];
	int64_t qsum_reverse[SPO_MAXLEN];
	/* the accumulators behind the three qcML plots (StatisticsReads.cpp:60-61,74-77) */
	int64_t base_qualities[100];      /* base_qualities_[q] */
	int64_t read_qualities[100];      /* read_qualities_[round(mean q of the read)] */
	int64_t qscore_dist_forward[60];  /* qscore_dist_r1: Histogram(0, 60, 1) of the mean quality of forward reads */
	int64_t qscore_dist_reverse[60];  /* qscore_dist_r2 */
} spo_qc_stats;

/* StatisticsReads::update for read 1 (FORWARD) and read 2 (REVERSE) of every pair of a SoA batch, added to *out */
void spo_qc_update_batch(const uint8_t* bases1, const uint8_t* quals1, const uint8_t* bases2, const uint8_t* quals2, const uint16_t* len1, const uint16_t* len2,
                         int stride, int64_t n, spo_qc_stats* out);

/* StatisticsReads::update for ONE read (direction 0 = FORWARD, 1 = REVERSE), as the ReadQC tool calls it per entry of -in1 / -in2
   (src/ReadQC/main.cpp:64-86) */
void spo_qc_update_read(const uint8_t* bases, const uint8_t* quals, int len, int reverse, spo_qc_stats* out);

/* FastqEntry::validate for short reads (src/cppNGS/FastqFileStream.cpp:3-48), which ReadQC's reader runs on every entry:
   0 = valid, 1 = header does not start with '@', 2 = header2 does not start with '+', 3 = |bases| != |qualities|,
   4 = base outside ACGTN, 5 = quality character outside 33..74 */
int spo_validate_entry(const char* header, int header_len, const char* bases, int bases_len, const char* header2, int header2_len, const char* quals, int quals_len);

/* The trimming rule of the FastqTrim tool for one read of `len` bases (src/FastqTrim/main.cpp:47-77). Returns 0 if the read is
   dropped, else 1 with the kept range [*first, *first + *count). */
int spo_fastq_trim(int len, int start, int end, int max_bases, int max_len, int* first, int* count);

void spo_default_params(spo_params* p);

/* BasicStatistics::factorial / matchProbability (src/cppCORE/BasicStatistics.cpp:249-307).
   spo_factorial returns NaN beyond the cache (n>170). */
double spo_factorial(int n);
double spo_match_probability(double p, int n, int count);

/* FastqEntry::trimQuality / trimN (src/cppNGS/FastqFileStream.cpp:52-117) on a read of length *len
   (bases and qualities have the same length). Return the number of removed bases. */
int spo_trim_quality(const char* quals, int* len, int cutoff, int window, int offset);
int spo_trim_n(const char* bases, int* len, int num_n);

/* Sequence::toReverseComplement (src/cppNGS/Sequence.cpp:41-84). Returns 0, or -1 on a byte outside ACGTN. */
int spo_revcomp(const char* in, int len, char* out);

/* One pass of AnalysisWorker::run's per-pair body (src/SeqPurge/AnalysisWorker.cpp:122-441), without the
   header check (host side, :110-120). r1/q1/r2/q2 hold len1/len2 bytes; with p->ec they may be edited in place
   (correctErrors, :19-77). ec may be NULL. */
void spo_process_pair(const spo_params* p, char* r1, char* q1, int len1, char* r2, char* q2, int len2,
                      spo_record* out, spo_ecstats* ec);

/* Batch form over the SoA slot layout of the C-ABI (fixed-stride rows, include/seqpurge_b200.h):
   bases1/quals1/bases2/quals2 are n rows of `stride` bytes, len1/len2 the row lengths.
   Rows are edited in place when p->ec. Block-parallel over `threads` pthreads like the reference's
   analysis pool (ThreadCoordinator.cpp:92-98); per-thread ec histograms are summed deterministically. */
void spo_trim_batch(const spo_params* p, uint8_t* bases1, uint8_t* quals1, uint8_t* bases2, uint8_t* quals2,
                    const uint16_t* len1, const uint16_t* len2, int stride, int64_t n, spo_record* out,
                    spo_ecstats* ec, int threads);

#ifdef __cplusplus
}
#endif
#endif
