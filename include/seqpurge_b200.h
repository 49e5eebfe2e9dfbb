/*
 * seqpurge_b200.h -- C ABI of the B200 SeqPurge trimming engine (libseqpurge_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of imgag/ngs-bits that this repository accelerates:
 * the per-read-pair body of AnalysisWorker::run (src/SeqPurge/AnalysisWorker.cpp:97-448 in the reference).
 * The reference has no FFI for this path -- the seam is the C++ class boundary
 *     ThreadCoordinator::analyze(int i)          src/SeqPurge/ThreadCoordinator.cpp:92-98
 *       -> new AnalysisWorker(job, params, stats, ecstats); run()   src/SeqPurge/AnalysisWorker.h:15-17
 * so every entry point below names the piece of that seam it replaces. INTEGRATION.md shows the
 * GpuAnalysisWorker a maintainer adds under src/SeqPurge to bind them.
 *
 * Conventions: plain pointers and sizes only, no C++/Qt/torch types, no exceptions. Every call returns
 * SPG_OK (0) or a negative SPG_ERR_* code; spg_last_error() gives the text (the shim turns it into
 * `emit error(i, msg)`, which the reference maps to THROW + exit(1), ThreadCoordinator.cpp:108-111).
 * There is no CPU fallback behind this ABI: without a CUDA device spg_create fails with SPG_ERR_CUDA.
 */
#ifndef SEQPURGE_B200_H
#define SEQPURGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPG_OK 0
#define SPG_ERR_PARAM (-1) /* invalid argument / parameter combination */
#define SPG_ERR_CUDA (-2)  /* CUDA runtime error or no device */
#define SPG_ERR_STATE (-3) /* slot used out of protocol (submit twice, wait without submit, ...) */
#define SPG_ERR_NOMEM (-4)

#define SPG_MAXLEN 1000 /* MAXLEN, src/SeqPurge/Auxilary.h:12: reads must be shorter than this */

/* spg_result.flags */
#define SPG_F_INSERT 0x01u  /* insert match: both reads cut to len2-best_offset   (AnalysisWorker.cpp:269-302) */
#define SPG_F_ADAPTER 0x02u /* adapter-only hit on at least one read               (AnalysisWorker.cpp:410-426) */
#define SPG_F_Q1 0x04u      /* read 1 lost bases to trimQuality                    (AnalysisWorker.cpp:432) */
#define SPG_F_Q2 0x08u      /* read 2 lost bases to trimQuality                    (AnalysisWorker.cpp:433) */
#define SPG_F_N1 0x10u      /* read 1 lost bases to trimN                          (AnalysisWorker.cpp:439) */
#define SPG_F_N2 0x20u      /* read 2 lost bases to trimN                          (AnalysisWorker.cpp:440) */

/* spg_result.status: per-pair conditions on which the reference throws */
#define SPG_PAIR_OK 0
#define SPG_PAIR_BAD_BASE_R2 1 /* byte outside ACGTN in read 2: Sequence::complement throws (src/cppNGS/Sequence.cpp:46-71) */
#define SPG_PAIR_TOO_LONG 2    /* max(len1,len2) >= MAXLEN: ArgumentException (AnalysisWorker.cpp:131-134) */
#define SPG_PAIR_BAD_BASE_EC 3 /* -ec had to complement a non-ACGTN byte of read 1 (AnalysisWorker.cpp:50, Sequence.cpp:103-112) */

/* Run constants. Replaces the fields of TrimmingParameters that AnalysisWorker reads
   (src/SeqPurge/Auxilary.h:100-133; defaults src/SeqPurge/main.cpp:25-43). a_size = min(20,|a1|,|a2|) is derived
   inside (main.cpp:71). */
typedef struct spg_params
{
	const char* a1;      /* forward adapter, >= 15 bytes (main.cpp:68) */
	int32_t a1_len;
	const char* a2;      /* reverse adapter, >= 15 bytes (main.cpp:70) */
	int32_t a2_len;
	int32_t adapter_overlap; /* 10 (Auxilary.h:103); 1..32 */
	double match_perc;   /* -match_perc, 80.0 */
	double mep;          /* -mep, 1e-6 */
	int32_t qcut;        /* -qcut 15 (0 disables) */
	int32_t qwin;        /* -qwin 5 */
	int32_t qoff;        /* -qoff 33 */
	int32_t ncut;        /* -ncut 7 (0 disables) */
	int32_t ec;          /* -ec: error-correct insert-hit pairs; edited rows are returned in the slot */
	int32_t qc;          /* -qc: also accumulate the raw-read statistics of every submitted batch (spg_qc_stats_get). 2: with the checks of
	                        FastqEntry::validate as ReadQC runs them (src/cppNGS/FastqFileStream.cpp:3-48): bases of exactly A,C,G,T,N and
	                        qualities of '!'..'J' (33..74), anything else counts in spg_qc_stats.errors. +4 (i.e. 5 or 6): also fill the histograms behind
	                        the three qcML plots (base_qualities, read_qualities, qscore_dist_*), which cost one shared-memory atomic per 32 bases */
} spg_params;

/* Per-pair output: everything OutputWorker / FastqWriter / TrimmingStatistics need (OutputWorker.cpp:36-77). 8 bytes. */
typedef struct spg_result
{
	uint16_t len1;       /* read 1: keep bases/qualities [0,len1) */
	uint16_t len2;       /* read 2: keep bases/qualities [0,len2) */
	int16_t best_offset; /* insert-match offset (for the adapter-consensus counters, AnalysisWorker.cpp:279-290), -1 if none */
	uint8_t flags;       /* SPG_F_* */
	uint8_t status;      /* SPG_PAIR_* ; when != 0 the other fields are 0 / -1 */
} spg_result;

/* One slot == one AnalysisJob of the reference's job pool (Auxilary.h:23-66): pinned host SoA the caller fills.
   Row r of each byte plane starts at r*stride; bytes beyond the read length are ignored. */
typedef struct spg_slot_view
{
	uint8_t* bases1; /* FastqEntry::bases of job.r1[r]      */
	uint8_t* quals1; /* FastqEntry::qualities of job.r1[r]  */
	uint8_t* bases2;
	uint8_t* quals2;
	uint16_t* len1;  /* bases1/quals1 row length (the reference assumes |bases|==|qualities|) */
	uint16_t* len2;
	int32_t stride;    /* bytes per row: max_len rounded up to an even number (>= 16) */
	int32_t max_pairs; /* capacity (the reference's -block_size) */
} spg_slot_view;

/* Error-correction histograms, replaces ErrorCorrectionStatistics (Auxilary.h:224-236) */
typedef struct spg_ec_stats
{
	int64_t mismatch_r1[SPG_MAXLEN];
	int64_t mismatch_r2[SPG_MAXLEN];
	int64_t errors_per_read[SPG_MAXLEN];
} spg_ec_stats;

/* Raw-read statistics of the untrimmed reads, the accumulators of StatisticsReads::update(FastqEntry, direction)
   (src/cppNGS/StatisticsReads.cpp:26-81) that a paired-end qcML report needs (StatisticsReads::getResult, :140-330);
   replaces `stats_.qc.update(...)` in AnalysisWorker::run (src/SeqPurge/AnalysisWorker.cpp:86-95). Qualities are Phred+33
   like FastqEntry::quality()'s default. */
typedef struct spg_qc_stats
{
	int64_t reads_forward;                /* c_forward_ */
	int64_t reads_reverse;                /* c_reverse_ */
	int64_t bases_sequenced;              /* bases_sequenced_ */
	int64_t read_q20;                     /* c_read_q20_: reads with mean quality >= 20 (reads of length 0 do not count) */
	int64_t base_q20;                     /* c_base_q20_ */
	int64_t base_q30;                     /* c_base_q30_ */
	int64_t errors;                       /* != 0: a base Pileup::inc throws on, or a quality outside 0..99 (the reference throws) */
	int64_t read_lengths[SPG_MAXLEN];     /* read_lengths_[cycles] */
	int64_t pileup[SPG_MAXLEN][5];        /* pileups_[cycle]: A, C, G, T, N (lower case counted like Pileup::inc) */
	int64_t qsum_forward[SPG_MAXLEN];     /* qualities1_[cycle] before the division by the depth */
	int64_t qsum_reverse[SPG_MAXLEN];     /* qualities2_[cycle] */
	/* the accumulators behind the three qcML plots (StatisticsReads.cpp:60-61,74-77; plotted in getResult, :300-440) */
	int64_t base_qualities[100];          /* base_qualities_[q]: bases by quality */
	int64_t read_qualities[100];          /* read_qualities_[round(mean quality of the read)] (reads of length 0 do not count) */
	int64_t qscore_dist_forward[60];      /* qscore_dist_r1 = Histogram(0, 60, 1) of the mean read quality, forward reads */
	int64_t qscore_dist_reverse[60];      /* qscore_dist_r2, reverse reads */
} spg_qc_stats;

typedef struct spg_ctx spg_ctx;

/* Replaces: main.cpp:92 (precalculateFactorials), ThreadCoordinator.cpp:40-48 (analysis pool + job pool allocation).
   Builds the decision tables (match-probability ranks etc.) with the host's libm exactly as
   BasicStatistics::matchProbability does, uploads them to every device, allocates n_slots pinned host slots and the
   matching device buffers. Slot s runs on device_ids[s % n_devices] (host-side round robin, no collective).
   max_len: longest read the caller will submit (<= 999); when it is one of the usual read lengths (75, 76, 100, 101, 125, 126, 150,
   151, 200, 201, 250, 251, 300, 301) the kernel variant compiled for that length runs (a fast path for full-length pairs; the results
   do not depend on it, see SPG_OPT_FULL_LEN). n_slots may be 0 (only spg_trim_device is used). */
int spg_create(spg_ctx** ctx, const spg_params* params, const int* device_ids, int n_devices, int n_slots, int max_pairs, int max_len);

/* Replaces: AnalysisJob storage (Auxilary.h:37-39). Pointers stay valid until spg_destroy. */
int spg_slot_buffers(spg_ctx* ctx, int slot, spg_slot_view* view);

/* Quality tails of a slot (optional, SPG_OPT_QUAL_TAILS): two more pinned planes of max_pairs x SPG_QTAIL bytes. Row r of qtail1 holds the
   last SPG_QTAIL qualities of read 1 of pair r, i.e. quals1[r*stride + len1[r] - SPG_QTAIL + j] for j = 0 .. SPG_QTAIL-1 (a shorter read
   right-aligned, the bytes in front of it are ignored); qtail2 likewise. The reference has no counterpart: FastqEntry::trimQuality
   (src/cppNGS/FastqFileStream.cpp:52-87) walks the quality string from its 3' end, and for most reads the decision falls within the last
   few bases. With the option set, spg_submit ships these 2 x 16 bytes per pair next to the bases while the quality rows stay in the
   pinned slot; the kernel goes to a row only for reads that were cut by the adapter steps or whose trimming point lies further left.
   The caller's stager writes the tails while it copies the quality string (GpuAnalysisWorker::start). */
#define SPG_QTAIL 16
int spg_slot_qtails(spg_ctx* ctx, int slot, uint8_t** qtail1, uint8_t** qtail2);

/* Replaces: thread_pool_analyze_.start(worker) (ThreadCoordinator.cpp:97). Asynchronous: H2D copy of the first n_pairs
   rows, the trimming kernel, D2H of the results (and of the edited rows with -ec) are queued on the slot's device stream. */
int spg_submit(spg_ctx* ctx, int slot, int n_pairs);

/* Replaces: AnalysisWorker's done(int)/error(int,QString) signals (AnalysisWorker.h:19-21). Blocks until the slot's
   work has finished; *results points at n_pairs records in pinned host memory, valid until the slot is resubmitted. */
int spg_wait(spg_ctx* ctx, int slot, const spg_result** results);

/* Device-resident form of the same operation (no copies): all pointers are device pointers on device_ids[device_index],
   16-byte aligned; row planes and len arrays must be readable up to a multiple of 8 rows / entries; stride even, 16..1008.
   cuda_stream is a cudaStream_t (NULL = default stream). With -ec the row planes are edited in place. */
int spg_trim_device(spg_ctx* ctx, int device_index, void* bases1, void* quals1, void* bases2, void* quals2, const uint16_t* len1, const uint16_t* len2,
                    int stride, int64_t n_pairs, spg_result* results, void* cuda_stream);

/* Accumulated -ec histograms of everything waited for so far (Auxilary.h:224-236). */
int spg_ec_stats_get(spg_ctx* ctx, spg_ec_stats* out);

/* Raw-read statistics of a device-resident batch (same pointer rules as spg_trim_device), added to the context's accumulators
   on that device. spg_submit does this by itself for every batch when params.qc is set. */
int spg_qc_device(spg_ctx* ctx, int device_index, const void* bases1, const void* quals1, const void* bases2, const void* quals2, const uint16_t* len1,
                  const uint16_t* len2, int stride, int64_t n_pairs, void* cuda_stream);

/* Sum of the -qc accumulators of all devices (synchronises the devices' streams first). */
int spg_qc_stats_get(spg_ctx* ctx, spg_qc_stats* out);

/* Text of the last error on this context (or of the last failed spg_create when ctx is NULL). */
const char* spg_last_error(spg_ctx* ctx);

/* Replaces: ~ThreadCoordinator. Waits for queued work, frees everything. */
void spg_destroy(spg_ctx* ctx);

/* ---- FASTQ text in, FASTQ text out ------------------------------------------------------------------
 * Replaces, around the trimming path: the line splitting of FastqFileStream::readEntry (src/cppNGS/FastqFileStream.cpp:135-160 over
 * VersatileFile::readLine, src/cppCORE/VersatileFile.cpp:274-399), the AoS->SoA flattening of the worker, the routing of
 * OutputWorker::run (src/SeqPurge/OutputWorker.cpp:36-57) and the record layout of FastqOutfileStream::write
 * (src/cppNGS/FastqFileStream.cpp:183-198), plus the adapter-consensus counters of AnalysisWorker.cpp:279-290. The host only
 * inflates the inputs into the slot's text buffers and deflates the output text.
 *
 * A text chunk must start at a record start; it may end anywhere. spg_fq_wait reports how many bytes of each chunk belong to the
 * pairs that were processed: the caller moves the rest to the front of the next chunk of that file. */
typedef struct spg_fq spg_fq;

typedef struct spg_fq_config
{
	int32_t n_slots;   /* chunks in flight; slot s runs on device s % n_devices of the context */
	int32_t max_pairs; /* pairs per chunk at most */
	int32_t max_len;   /* longest read the rows can take (< 1000); spg_fq_output.max_len tells when a chunk needs more */
	int64_t text_cap;  /* bytes of text per file and chunk */
	int32_t min_len;   /* -min_len: reads shorter than this after trimming are dropped (OutputWorker.cpp:41-56) */
	int32_t singles;   /* 1: -out3 given, reads whose mate was dropped go to out[2] (read 1) / out[3] (read 2) */
	int32_t stats_only; /* 1: framing and the -qc statistics only (the ReadQC tool, src/ReadQC/main.cpp:58-101): no trimming, no output text;
	                       spg_fq_output carries the pair count, the consumed bytes, the lengths and the framing status */
	int32_t single_end; /* 1 (with stats_only): only text1 is given; every record is one forward read, nothing is counted as reverse read */
	int32_t validate;   /* 1: records are checked like FastqEntry::validate does (header starts with '@', header2 with '+', equal lengths) */
	int32_t fixed_trim; /* 1 (with single_end): the FastqTrim tool (src/FastqTrim/main.cpp:47-77) instead of SeqPurge's trimming: every read of
	                       text1 loses trim_start bases at its start and trim_end at its end (reads of at most trim_start+trim_end bases are
	                       dropped), is then cut to trim_len bases (0 = no limit); reads of trim_max_len bases and more (0 = off) pass
	                       unchanged. Output text in out[0]; results[].len1 = new length, best_offset = first base kept, flags 0x80 = dropped */
	int32_t trim_start, trim_end, trim_len, trim_max_len;
} spg_fq_config;
#define SPG_F_DROPPED 0x80u /* fixed_trim: the read is not written */

typedef struct spg_fq_input
{
	uint8_t* text1; /* pinned host buffers of text_cap bytes */
	uint8_t* text2;
	int64_t cap;
} spg_fq_input;

/* framing status of a pair (spg_fq_output.frame_status) */
#define SPG_FQ_OK 0
#define SPG_FQ_HEADER_MISMATCH 1 /* "Headers of reads do not match" (AnalysisWorker.cpp:110-120) */
#define SPG_FQ_LENGTH_MISMATCH 2 /* |bases| != |qualities| */
#define SPG_FQ_TOO_LONG 3        /* read longer than max_len of this stream (or >= 1000) */
#define SPG_FQ_BAD_HEADER 4      /* validate: "First header line does not start with '@'" (FastqFileStream.cpp:7-10) */
#define SPG_FQ_BAD_HEADER2 5     /* validate: "Second header line does not start with '+'" (FastqFileStream.cpp:11-14) */

/* Summary counters of one chunk, what OutputWorker::run adds to TrimmingStatistics (src/SeqPurge/OutputWorker.cpp:59-77), reduced on the
   device from the result records. bases_perc_trim_sum of the reference is sum over reads of (o - len) / o in double, in job order; here
   the exact integer sums per original length o are returned and the caller adds trimmed_bases_by_length[o] / o -- the same value up to the
   rounding order of a floating-point sum (the reference's own order depends on which job finishes first). */
typedef struct spg_fq_stats
{
	int64_t reads_trimmed_insert;  /* 2 per pair with an insert match */
	int64_t reads_trimmed_adapter; /* 2 per pair with an adapter-only hit */
	int64_t reads_trimmed_q;       /* reads that lost bases to trimQuality */
	int64_t reads_trimmed_n;       /* reads that lost bases to trimN */
	int64_t reads_removed;         /* reads not written (shorter than min_len, or mate removed without -out3) */
	int64_t bases_remaining[SPG_MAXLEN];         /* [len]: reads of that length after trimming */
	int64_t trimmed_bases_by_length[SPG_MAXLEN]; /* [o]: bases removed from reads of original length o */
} spg_fq_stats;

typedef struct spg_fq_output
{
	int32_t n_pairs;             /* pairs processed = min(records1, records2, max_pairs) */
	int32_t records1, records2;  /* records found in the two chunks */
	int64_t consumed1, consumed2; /* bytes of the chunks that belong to the processed pairs */
	const uint8_t* out[4];       /* out1, out2, out3 (read-1 singletons), out4 (read-2 singletons): FASTQ text, pinned host memory */
	int64_t out_bytes[4];
	const spg_result* results;   /* [n_pairs] as spg_wait */
	const uint16_t* len1;        /* [n_pairs] untrimmed lengths */
	const uint16_t* len2;
	const uint8_t* frame_status; /* [n_pairs] SPG_FQ_* */
	int32_t error_pair;          /* first pair with frame_status != 0 or results[].status != 0, or -1 */
	int32_t max_len;             /* longest bases/qualities line of the chunk */
	int32_t invalid_chars;       /* stats_only: != 0 if the chunk holds a base or quality that counts in spg_qc_stats.errors */
	const spg_fq_stats* stats;   /* summary counters of this chunk (NULL with stats_only / fixed_trim) */
} spg_fq_output;

/* Attaches a FASTQ stream to a context (which provides parameters, tables, devices; it may have been created with n_slots = 0). */
int spg_fq_open(spg_ctx* ctx, const spg_fq_config* cfg, spg_fq** out);
int spg_fq_buffers(spg_fq* fq, int slot, spg_fq_input* in);
/* Queues: H2D of the two chunks, framing, [-qc statistics], trimming, output assembly, D2H. final1/final2: the chunk holds the end
   of its file (an unterminated last line and an incomplete last record then count, as for the reference's reader). */
int spg_fq_submit(spg_fq* fq, int slot, int64_t bytes1, int64_t bytes2, int final1, int final2);
int spg_fq_wait(spg_fq* fq, int slot, spg_fq_output* out);
/* Adapter-consensus counters of all chunks so far: counts[read][position 0..39][A,C,G,T,N]; *unknown_base != 0 if a base outside
   ACGTN (and '-', '~') was met, where the reference's Pileup::inc throws. */
int spg_fq_consensus_get(spg_fq* fq, int64_t counts[2][40][5], int32_t* unknown_base);
void spg_fq_close(spg_fq* fq);

/* ---- tuning / introspection (not part of the reference seam) ------------------------------------------------------------ */

#define SPG_OPT_FORCE_BYTEWISE 1 /* value 1: route every pair through the byte-wise kernel path (cross-check of the bit-plane path) */
#define SPG_OPT_GRID_CTAS_PER_SM 2 /* cap on resident CTAs per SM (0 = what the occupancy calculator allows) */
#define SPG_OPT_MIN_BLOCKS 3      /* kernel variant compiled for at least 2, 3 or 4 resident CTAs per SM (register budget) */
#define SPG_OPT_TILE_PAIRS 4      /* pairs per TMA-staged tile (multiple of 8; 0 = automatic) */
#define SPG_OPT_STAGES 5          /* depth of the TMA ring, 2..4 (0 = automatic) */
#define SPG_OPT_FULL_LEN 6        /* read length the kernel variant is chosen for (fast path for full-length pairs): -1 = automatic, 0 = general kernel only */
#define SPG_OPT_KERNEL 7          /* thread layout of the read-length variants: 0 = automatic (one lane per pair where it applies), 1 = one warp per pair, 2 = as 0 */
#define SPG_OPT_SEED_SCAN 8       /* lane-per-pair kernel: 1 (default) = exact-block filter in front of the adapter scans where the parameters allow it, 0 = every offset */
#define SPG_OPT_ZERO_COPY_QUALS 9 /* slots (spg_submit): 1 (default) = when the lane-per-pair kernel runs, the quality planes are not copied; the kernel reads
                                     the few quality bytes it needs from the pinned slot over PCIe. 0 = all four planes are copied */
#define SPG_OPT_N_LANES 10        /* lane-per-pair kernel: 1 (default) = pairs with N take its N-aware path, 0 = the warp-cooperative general path */
#define SPG_OPT_QUAL_TAILS 11     /* slots (spg_submit): 1 = the caller fills qtail1 / qtail2 of every slot it submits (spg_slot_qtails); they are copied with the
                                     bases when the quality planes stay in the slot (SPG_OPT_ZERO_COPY_QUALS). 0 (default) = the tails are not looked at */
int spg_set_option(spg_ctx* ctx, int option, int value);

int spg_get_option(spg_ctx* ctx, int option, int* value);

/* Which trimming kernel the last launch of this context ran (bench.py reports it next to the roofline): writes the instantiation's name
   into name[cap]; returns layout * 100000 + NW * 1000 + FULL (layout 0: general kernel, 1: warp per pair, 2: lane per pair), 0 before
   the first launch. */
int spg_last_kernel(spg_ctx* ctx, char* name, int cap);

/* number of kernel launches issued by this context so far (bench.py reports it as gpu_launches) */
int64_t spg_launch_count(spg_ctx* ctx);

/* ---- synthetic read pairs, generated on the device (bench / tests) --------------------------------------- */
typedef struct spg_synth_config
{
	int32_t read_len;      /* L */
	float insert_mean;     /* mu */
	float insert_sd;       /* sigma */
	int32_t insert_min;    /* clip */
	int32_t insert_max;
	float error_rate;      /* i.i.d. substitution rate */
	float n_rate;          /* per-base N rate */
	float lowq_tail_mean;  /* mean length of the low-quality 3' tail ('#'), 0 = none */
	float n_run_rate;      /* fraction of reads that get a run of >= 7 N */
	int32_t binned_quals;  /* 1: NovaSeq-like binned qualities F : , # */
	uint64_t seed;
	const char* a1;        /* adapters appended after the insert (33 bytes used at most; NULL = Illumina defaults) */
	const char* a2;
} spg_synth_config;

/* Fills device row planes for pairs [first_pair, first_pair+n_pairs) of the stream defined by cfg (counter based:
   any slice can be regenerated). Pointers as in spg_trim_device. */
int spg_synth_device(int device_id, const spg_synth_config* cfg, int64_t first_pair, int64_t n_pairs, void* bases1, void* quals1, void* bases2, void* quals2,
                     uint16_t* len1, uint16_t* len2, int stride, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
