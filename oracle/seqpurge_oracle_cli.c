/*
 * seqpurge_oracle_cli.c -- Qt-free command line around the oracle (TEST INFRASTRUCTURE ONLY).
 *
 * Restates the I/O shell of the reference tool so that the oracle can be pinned against the reference's golden
 * FASTQ files:
 *   FASTQ reading   src/cppNGS/FastqFileStream.cpp:135-158 (readEntry with one-line look-ahead),
 *                   src/cppCORE/VersatileFile.cpp:286-308,395-414 (gzgets into a 1 KiB buffer, trim \r\n, gzeof)
 *   block loading   src/SeqPurge/InputWorker.cpp:16-77
 *   header check    src/SeqPurge/AnalysisWorker.cpp:110-120
 *   routing/stats   src/SeqPurge/OutputWorker.cpp:36-77, src/SeqPurge/FastqWriter.cpp:17-38
 *   FASTQ writing   src/cppNGS/FastqFileStream.cpp:160-193 (gzopen wb, gzbuffer 131072, gzsetparams, 8 gzputs)
 *   summary         src/SeqPurge/Auxilary.h:166-221,238-269
 * Flags and defaults: src/SeqPurge/main.cpp:20-43.  Not supported here: -qc (qcML), -debug, -progress.
 *
 * With -threads N the pairs of one block are analysed by N threads but blocks retire in input order, i.e. the
 * output always equals the reference's `-threads 1` output (the only order the reference's tests pin).
 */
#include "seqpurge_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

static void die(const char* msg, const char* a, const char* b)
{
	fprintf(stderr, "seqpurge_oracle: %s%s%s\n", msg, a ? a : "", b ? b : "");
	exit(1);
}

/* ---- growable byte string ------------------------------------------------------------------------------------ */
typedef struct str
{
	char* d;
	int n, cap;
} str;

static void str_reserve(str* s, int cap)
{
	if (cap <= s->cap) return;
	int c = s->cap ? s->cap : 64;
	while (c < cap) c *= 2;
	s->d = (char*)realloc(s->d, (size_t)c);
	s->cap = c;
}
static void str_append(str* s, const char* p, int n)
{
	str_reserve(s, s->n + n + 1);
	memcpy(s->d + s->n, p, (size_t)n);
	s->n += n;
	s->d[s->n] = 0;
}
static void str_assign(str* s, const str* o)
{
	s->n = 0;
	str_reserve(s, o->n + 1);
	if (o->n) memcpy(s->d, o->d, (size_t)o->n);
	s->n = o->n;
	s->d[s->n] = 0;
}

/* ---- FASTQ input stream ------------------------------------------------------------------------------------------ */
typedef struct fq_in
{
	gzFile gz;
	const char* name;
	int first;
	str last; /* look-ahead line */
	char buf[1024];
} fq_in;

static void fq_open(fq_in* f, const char* name)
{
	memset(f, 0, sizeof(*f));
	f->name = name;
	f->gz = gzopen(name, "rb");
	if (!f->gz) die("Could not open file for reading: ", name, NULL);
	gzbuffer(f->gz, 128 * 1024); /* FastqFileStream.cpp:125-128 */
	f->first = 1;
}
static void fq_close(fq_in* f)
{
	if (f->gz) gzclose(f->gz);
	f->gz = NULL;
	free(f->last.d);
	f->last.d = NULL;
	f->last.n = f->last.cap = 0;
}
static void fq_read_line(fq_in* f, str* out) /* VersatileFile::readLine(true) */
{
	out->n = 0;
	str_reserve(out, 1);
	out->d[0] = 0;
	for (;;)
	{
		char* s = gzgets(f->gz, f->buf, (int)sizeof(f->buf));
		if (s == NULL)
		{
			int err = Z_OK;
			const char* msg = gzerror(f->gz, &err);
			if (err != Z_OK && err != Z_STREAM_END) die("Error while reading file: ", f->name, msg);
			break;
		}
		str_append(out, s, (int)strlen(s));
		if (out->n > 0 && out->d[out->n - 1] == '\n') break;
	}
	while (out->n > 0 && (out->d[out->n - 1] == '\n' || out->d[out->n - 1] == '\r')) out->d[--out->n] = 0;
}

typedef struct entry
{
	str header, bases, header2, quals;
} entry;

static void fq_read_entry(fq_in* f, entry* e) /* FastqFileStream::readEntry */
{
	if (f->first)
	{
		fq_read_line(f, &f->last);
		f->first = 0;
	}
	str_assign(&e->header, &f->last);
	fq_read_line(f, &e->bases);
	fq_read_line(f, &e->header2);
	fq_read_line(f, &e->quals);
	fq_read_line(f, &f->last);
}
static int fq_at_end(fq_in* f) { return gzeof(f->gz); }

/* ---- FASTQ output stream ----------------------------------------------------------------------------------------- */
static gzFile out_open(const char* name, int level)
{
	gzFile g = gzopen(name, "wb");
	if (!g) die("Could not open file for writing: ", name, NULL);
	gzbuffer(g, 131072);
	if (level < 0 || level > 9) die("Invalid gzip compression level for FASTQ file ", name, NULL);
	gzsetparams(g, level, Z_DEFAULT_STRATEGY);
	return g;
}
static void out_write(gzFile g, const entry* e, int len)
{
	/* the entry's bases/quals were truncated to len: terminate, write, restore is not needed (entries are reloaded) */
	e->bases.d[len] = 0;
	e->quals.d[len] = 0;
	if (gzputs(g, e->header.d) == -1 || gzputs(g, "\n") == -1 || gzputs(g, e->bases.d) == -1 || gzputs(g, "\n") == -1 || gzputs(g, e->header2.d) == -1
	    || gzputs(g, "\n") == -1 || gzputs(g, e->quals.d) == -1 || gzputs(g, "\n") == -1)
	{
		die("Could not write to output file", NULL, NULL);
	}
}

/* ---- header check (AnalysisWorker.cpp:110-120) ---------------------------------------------------------------------- */
static void check_headers(const str* h1, const str* h2)
{
	int n1 = 0, n2 = 0;
	while (n1 < h1->n && h1->d[n1] != ' ') ++n1;
	while (n2 < h2->n && h2->d[n2] != ' ') ++n2;
	if (n1 >= 2 && n2 >= 2 && h1->d[n1 - 2] == '/' && h1->d[n1 - 1] == '1' && h2->d[n2 - 2] == '/' && h2->d[n2 - 1] == '2')
	{
		n1 -= 2;
		n2 -= 2;
	}
	if (n1 != n2 || memcmp(h1->d, h2->d, (size_t)n1) != 0)
	{
		fprintf(stderr, "seqpurge_oracle: Headers of reads do not match:\n%.*s\n%.*s\n", n1, h1->d, n2, h2->d);
		exit(1);
	}
}

/* ---- statistics (Auxilary.h:136-221) --------------------------------------------------------------------------------- */
typedef struct pile
{
	long a, c, g, t, n;
} pile;
static void pile_inc(pile* p, char b) /* Pileup::inc (src/cppNGS/Pileup.cpp:17-32) */
{
	switch (b)
	{
		case 'A': case 'a': ++p->a; break;
		case 'C': case 'c': ++p->c; break;
		case 'G': case 'g': ++p->g; break;
		case 'T': case 't': ++p->t; break;
		case 'N': case 'n': ++p->n; break;
		case '-': case '~': break;
		default: die("Unknown base in pileup!", NULL, NULL);
	}
}
typedef struct stats
{
	long read_num;
	double bases_remaining[SPO_MAXLEN];
	pile acons1[40], acons2[40];
	double trimmed_insert, trimmed_adapter, trimmed_q, trimmed_n, removed, bases_perc_trim_sum;
} stats;

static void consensus(FILE* f, const pile* ac)
{
	for (int i = 0; i < 40; ++i)
	{
		long depth = ac[i].a + ac[i].c + ac[i].g + ac[i].t;
		if (depth < 20) break;
		long mx = ac[i].a;
		if (ac[i].c > mx) mx = ac[i].c;
		if (ac[i].g > mx) mx = ac[i].g;
		if (ac[i].t > mx) mx = ac[i].t;
		if ((double)mx / depth <= 0.5) fputc('N', f);
		else if (ac[i].a == mx) fputc('A', f);
		else if (ac[i].c == mx) fputc('C', f);
		else if (ac[i].g == mx) fputc('G', f);
		else if (ac[i].t == mx) fputc('T', f);
	}
	fputc('\n', f);
}
static void write_summary(FILE* f, const stats* s, const spo_params* p, const spo_ecstats* ec)
{
	fprintf(f, "Reads (forward + reverse): %ld\n\n", s->read_num);
	fprintf(f, "Reads trimmed by insert match: %ld\n", (long)s->trimmed_insert);
	fprintf(f, "Reads trimmed by adapter match: %ld\n", (long)s->trimmed_adapter);
	fprintf(f, "Reads trimmed by quality: %ld\n", (long)s->trimmed_q);
	fprintf(f, "Reads trimmed by N stretches: %ld\n", (long)s->trimmed_n);
	double trimmed = s->trimmed_insert + s->trimmed_adapter;
	fprintf(f, "Trimmed reads: %ld of %ld (%.2f%%)\n", (long)trimmed, s->read_num, 100.0 * trimmed / s->read_num);
	fprintf(f, "Removed reads: %ld of %ld (%.2f%%)\n", (long)s->removed, s->read_num, 100.0 * s->removed / s->read_num);
	fprintf(f, "Removed bases: %.2f%%\n\n", 100.0 * s->bases_perc_trim_sum / s->read_num);
	fprintf(f, "Forward adapter sequence (given)    : %.*s\n", p->a1_len, p->a1);
	fprintf(f, "Forward adapter sequence (consensus): ");
	consensus(f, s->acons1);
	fprintf(f, "Reverse adapter sequence (given)    : %.*s\n", p->a2_len, p->a2);
	fprintf(f, "Reverse adapter sequence (consensus): ");
	consensus(f, s->acons2);
	fprintf(f, "\nRead length distribution after trimming:\n");
	int max = SPO_MAXLEN - 1;
	while (max > 0 && s->bases_remaining[max] == 0) --max;
	for (int i = 0; i <= max; ++i) fprintf(f, "%4d: %ld\n", i, (long)s->bases_remaining[i]);
	if (ec)
	{
		const int64_t* arrs[3] = {ec->mismatch_r1, ec->mismatch_r2, ec->errors_per_read};
		const char* titles[3] = {"Read error per cycle (read 1):", "Read error per cycle (read 2):", "Read error count distribution:"};
		for (int k = 0; k < 3; ++k)
		{
			fprintf(f, "\n%s\n", titles[k]);
			int m = SPO_MAXLEN - 1;
			while (m > 0 && arrs[k][m] == 0) --m;
			for (int i = 1; i <= m; ++i) fprintf(f, "%4d: %ld\n", i, (long)arrs[k][i]);
		}
	}
}

/* ---- block analysis over threads ---------------------------------------------------------------------------------------- */
typedef struct block
{
	const spo_params* p;
	entry *r1, *r2;
	spo_record* rec;
	int count;
	int next;
	pthread_mutex_t mu;
	spo_ecstats* ec; /* shared; only touched with -ec, guarded by running -ec single-threaded per block */
} block;

static void* block_worker(void* arg)
{
	block* b = (block*)arg;
	for (;;)
	{
		pthread_mutex_lock(&b->mu);
		int r = b->next;
		b->next += 64;
		pthread_mutex_unlock(&b->mu);
		if (r >= b->count) break;
		int end = r + 64 < b->count ? r + 64 : b->count;
		for (; r < end; ++r)
		{
			spo_process_pair(b->p, b->r1[r].bases.d, b->r1[r].quals.d, b->r1[r].bases.n, b->r2[r].bases.d, b->r2[r].quals.d, b->r2[r].bases.n, &b->rec[r], b->ec);
		}
	}
	return NULL;
}

/* ---- main ------------------------------------------------------------------------------------------------------------------ */
#define MAXFILES 256
int main(int argc, char** argv)
{
	const char* in1[MAXFILES];
	const char* in2[MAXFILES];
	int n_in1 = 0, n_in2 = 0;
	const char *out1 = NULL, *out2 = NULL, *out3 = NULL, *summary = NULL;
	const char* a1 = "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA";
	const char* a2 = "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT";
	spo_params p;
	spo_default_params(&p);
	int min_len = 30, threads = 1, block_size = 10000, level = Z_BEST_SPEED;

	for (int i = 1; i < argc; ++i)
	{
		const char* f = argv[i];
#define NEXT() (i + 1 < argc ? argv[++i] : (die("missing value for ", f, NULL), ""))
		if (!strcmp(f, "-in1")) { while (i + 1 < argc && argv[i + 1][0] != '-') in1[n_in1++] = argv[++i]; }
		else if (!strcmp(f, "-in2")) { while (i + 1 < argc && argv[i + 1][0] != '-') in2[n_in2++] = argv[++i]; }
		else if (!strcmp(f, "-out1")) out1 = NEXT();
		else if (!strcmp(f, "-out2")) out2 = NEXT();
		else if (!strcmp(f, "-out3")) out3 = NEXT();
		else if (!strcmp(f, "-summary")) summary = NEXT();
		else if (!strcmp(f, "-a1")) a1 = NEXT();
		else if (!strcmp(f, "-a2")) a2 = NEXT();
		else if (!strcmp(f, "-match_perc")) p.match_perc = atof(NEXT());
		else if (!strcmp(f, "-mep")) p.mep = atof(NEXT());
		else if (!strcmp(f, "-qcut")) p.qcut = atoi(NEXT());
		else if (!strcmp(f, "-qwin")) p.qwin = atoi(NEXT());
		else if (!strcmp(f, "-qoff")) p.qoff = atoi(NEXT());
		else if (!strcmp(f, "-ncut")) p.ncut = atoi(NEXT());
		else if (!strcmp(f, "-min_len")) min_len = atoi(NEXT());
		else if (!strcmp(f, "-threads")) threads = atoi(NEXT());
		else if (!strcmp(f, "-block_size")) block_size = atoi(NEXT());
		else if (!strcmp(f, "-block_prefetch")) (void)NEXT();
		else if (!strcmp(f, "-progress")) (void)NEXT();
		else if (!strcmp(f, "-compression_level")) level = atoi(NEXT());
		else if (!strcmp(f, "-ec")) p.ec = 1;
		else die("unknown or unsupported flag ", f, NULL);
	}
	if (n_in1 == 0 || n_in2 == 0 || !out1 || !out2) die("usage: -in1 .. -in2 .. -out1 . -out2 . [SeqPurge flags]", NULL, NULL);
	if (n_in1 != n_in2) die("Input file lists 'in1' and 'in2' differ in counts!", NULL, NULL);
	p.a1 = a1;
	p.a1_len = (int)strlen(a1);
	p.a2 = a2;
	p.a2_len = (int)strlen(a2);
	if (p.a1_len < 15) die("Forward adapter too short: ", a1, NULL);
	if (p.a2_len < 15) die("Reverse adapter too short: ", a2, NULL);
	p.a_size = 20;
	if (p.a1_len < p.a_size) p.a_size = p.a1_len;
	if (p.a2_len < p.a_size) p.a_size = p.a2_len;
	if (threads < 1) threads = 1;
	if (p.ec) threads = 1; /* shared ec histograms; the reference races here with -threads>1 */

	fq_in s1, s2;
	int cur = 0;
	fq_open(&s1, in1[0]);
	fq_open(&s2, in2[0]);
	gzFile o1 = out_open(out1, level), o2 = out_open(out2, level), o3 = NULL, o4 = NULL;
	if (out3 && out3[0])
	{
		char name[4096];
		snprintf(name, sizeof(name), "%s_R1.fastq.gz", out3);
		o3 = out_open(name, level);
		snprintf(name, sizeof(name), "%s_R2.fastq.gz", out3);
		o4 = out_open(name, level);
	}

	entry* r1 = (entry*)calloc((size_t)block_size, sizeof(entry));
	entry* r2 = (entry*)calloc((size_t)block_size, sizeof(entry));
	spo_record* rec = (spo_record*)calloc((size_t)block_size, sizeof(spo_record));
	stats* st = (stats*)calloc(1, sizeof(stats));
	spo_ecstats* ec = p.ec ? (spo_ecstats*)calloc(1, sizeof(spo_ecstats)) : NULL;
	pthread_t* tids = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
	spo_factorial(0);

	int end_of_data = 0;
	while (!end_of_data)
	{
		/* InputWorker::run (InputWorker.cpp:16-77) */
		int pairs = 0;
		while (pairs < block_size && !end_of_data)
		{
			if (fq_at_end(&s1) && fq_at_end(&s2))
			{
				++cur;
				if (cur >= n_in1) end_of_data = 1;
				else
				{
					fq_close(&s1);
					fq_close(&s2);
					fq_open(&s1, in1[cur]);
					fq_open(&s2, in2[cur]);
				}
			}
			else if (fq_at_end(&s1)) die("File has more entries than its mate: ", s2.name, NULL);
			else if (fq_at_end(&s2)) die("File has more entries than its mate: ", s1.name, NULL);
			if (!end_of_data)
			{
				fq_read_entry(&s1, &r1[pairs]);
				fq_read_entry(&s2, &r2[pairs]);
				++pairs;
			}
		}
		if (pairs == 0) break;

		/* AnalysisWorker::run */
		for (int r = 0; r < pairs; ++r)
		{
			check_headers(&r1[r].header, &r2[r].header);
			if (r1[r].bases.n != r1[r].quals.n || r2[r].bases.n != r2[r].quals.n) die("bases/qualities length mismatch (unsupported by the oracle): ", r1[r].header.d, NULL);
			str_reserve(&r1[r].bases, 1);
			str_reserve(&r1[r].quals, 1);
			str_reserve(&r2[r].bases, 1);
			str_reserve(&r2[r].quals, 1);
		}
		block b;
		b.p = &p;
		b.r1 = r1;
		b.r2 = r2;
		b.rec = rec;
		b.count = pairs;
		b.next = 0;
		b.ec = ec;
		pthread_mutex_init(&b.mu, NULL);
		if (threads == 1) block_worker(&b);
		else
		{
			for (int t = 0; t < threads; ++t) pthread_create(&tids[t], NULL, block_worker, &b);
			for (int t = 0; t < threads; ++t) pthread_join(tids[t], NULL);
		}
		pthread_mutex_destroy(&b.mu);

		/* OutputWorker::run + FastqWriter::run, and the statistics of AnalysisWorker.cpp:279-293 */
		int removed = 0;
		for (int r = 0; r < pairs; ++r)
		{
			const spo_record* k = &rec[r];
			if (k->status == SPO_E_BASE_R2) die("Could not convert base to complement! read: ", r2[r].header.d, NULL);
			if (k->status == SPO_E_MAXLEN) die("Read length unsupported! A maximum read length of 1000 is supported!", NULL, NULL);
			if (k->status != SPO_OK) die("Could not convert base to complement (error correction)! read: ", r1[r].header.d, NULL);
			int l1o = r1[r].bases.n, l2o = r2[r].bases.n;
			if (k->flags & SPO_F_INSERT)
			{
				int new_length = l2o - k->best_offset;
				for (int i = 0; i < 40 && new_length + i < l1o; ++i) pile_inc(&st->acons1[i], r1[r].bases.d[new_length + i]);
				for (int i = 0; i < 40 && i < k->best_offset; ++i) pile_inc(&st->acons2[i], r2[r].bases.d[l2o - k->best_offset + i]);
				st->trimmed_insert += 2;
			}
			if (k->flags & SPO_F_ADAPTER) st->trimmed_adapter += 2;
			st->trimmed_q += ((k->flags & SPO_F_Q1) != 0) + ((k->flags & SPO_F_Q2) != 0);
			st->trimmed_n += ((k->flags & SPO_F_N1) != 0) + ((k->flags & SPO_F_N2) != 0);

			if (k->len1 >= min_len && k->len2 >= min_len)
			{
				out_write(o1, &r1[r], k->len1);
				out_write(o2, &r2[r], k->len2);
			}
			else if (o3 && k->len1 >= min_len)
			{
				removed += 1;
				out_write(o3, &r1[r], k->len1);
			}
			else if (o4 && k->len2 >= min_len)
			{
				removed += 1;
				out_write(o4, &r2[r], k->len2);
			}
			else removed += 2;

			st->bases_remaining[k->len1] += 1;
			st->bases_remaining[k->len2] += 1;
			if (l1o > 0) st->bases_perc_trim_sum += (double)(l1o - k->len1) / l1o;
			if (l2o > 0) st->bases_perc_trim_sum += (double)(l2o - k->len2) / l2o;
		}
		st->read_num += 2L * pairs;
		st->removed += removed;
	}

	gzclose(o1);
	gzclose(o2);
	if (o3) gzclose(o3);
	if (o4) gzclose(o4);
	FILE* sf = summary && summary[0] ? fopen(summary, "w") : stdout;
	if (!sf) die("Could not open summary file: ", summary, NULL);
	write_summary(sf, st, &p, ec);
	if (sf != stdout) fclose(sf);
	return 0;
}
