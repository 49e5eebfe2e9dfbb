// spg_fastq.cuh -- FASTQ framing and output assembly on the device.
//
// The host hands over inflated FASTQ text of the two input files (any cut: a chunk starts at a record start and may end anywhere).
// The kernels below do what FastqFileStream::readEntry (src/cppNGS/FastqFileStream.cpp:135-160, over VersatileFile::readLine,
// src/cppCORE/VersatileFile.cpp:274-399: a line ends at '\n', trailing '\n'/'\r' are chopped, four lines make an entry, no
// validation) and the AoS->SoA flattening in front of the trimming kernel do, then -- behind the trimming kernel -- what
// OutputWorker::run + FastqOutfileStream::write do (src/SeqPurge/OutputWorker.cpp:36-57, FastqFileStream.cpp:183-198): route every
// pair by the trimmed lengths (both >= min_len -> out1/out2; one -> out3/out4 singleton files, if given; else dropped) and lay the
// records out as text "header\nbases\nheader2\nqualities\n" in submission order. The adapter-consensus counters of
// AnalysisWorker.cpp:279-290 are reduced in the same pass.
//
//   fq_count_newlines -> fq_scan_counts -> fq_scatter_newlines     line index of both texts (positions of '\n')
//   fq_plan                                                        records per file, pairs of this batch, bytes consumed
//   fq_pack                                                        rows + lengths for the trimming kernel, header check (a3)
//   [qc_kernel] trim_kernel                                        (spg_qc.cuh, spg_kernel.cuh; pair count read from the plan)
//   fq_out_sizes -> fq_out_scan -> fq_out_write                    output text of the four streams, consensus counters
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/seqpurge_b200.h"

namespace spg
{

constexpr int kFqBlockThreads = 256;
constexpr int kFqBytesPerThread = 64;                                   // four 16-byte loads
constexpr int kFqBlockBytes = kFqBlockThreads * kFqBytesPerThread;      // 16 KiB of text per CTA
constexpr int kFqOutPairs = 256;                                        // pairs per CTA in the output kernels

// per-pair framing status (first failing check in the order of the reference's worker)
enum : uint8_t
{
	FQ_OK = 0,
	FQ_HEADER_MISMATCH = 1, // AnalysisWorker.cpp:110-120
	FQ_LENGTH_MISMATCH = 2, // |bases| != |qualities|
	FQ_TOO_LONG = 3,        // longer than the row stride of this engine (or >= MAXLEN)
	FQ_BAD_HEADER = 4,      // validate: header does not start with '@'
	FQ_BAD_HEADER2 = 5      // validate: header2 does not start with '+'
};

struct __align__(16) FqRec // where the two header lines of a record are in the text
{
	uint32_t hs, hl;   // start offset and length (without line ending) of the header line
	uint32_t h2s, h2l; // same for header2
};

struct FqPlan // device-resident per slot; copied to the host after the batch
{
	int n_pairs;
	int lines[2];        // lines of each text that take part (complete lines; + the unterminated last line of a final chunk)
	int records[2];      // records available in each text
	uint32_t consumed[2]; // bytes of each text that belong to the n_pairs records
	uint32_t out_bytes[4];
	int err_pair;        // smallest pair index with a framing status != 0, or INT_MAX
	int max_len;         // longest bases/qualities line seen in the batch
	int acons_error;     // a base outside ACGTN in the consensus window (Pileup::inc throws there)
	int pad;
};

struct FqArgs
{
	const uint8_t* text[2];
	uint32_t bytes[2];
	int final_[2];       // the text holds the end of its file
	uint32_t* nl[2];     // [nl_cap] positions of the line ends
	int nl_cap;
	uint32_t* block_counts[2]; // [n_blocks] newline count per 16 KiB block, then exclusive offsets
	int n_blocks[2];
	int max_pairs;
	FqPlan* plan;
	// rows for the trimming kernel
	uint8_t* rows[4]; // b1, q1, b2, q2
	uint16_t* len[2];
	int stride;
	FqRec* rec[2];
	unsigned long long* stats; // spg_fq_stats of this chunk as 64-bit words (5 counters, then the two histograms), or null
	uint8_t* fstat;   // [max_pairs]
	// output
	const spg_result* res;
	uint8_t* out[4];  // out1, out2, out3 (read 1 singletons), out4 (read 2 singletons)
	uint32_t* out_block[4]; // [n_out_blocks] per-CTA byte counts, then exclusive offsets
	int min_len;
	int singles;      // -out3 given
	unsigned long long* acons; // [2][40][5] A,C,G,T,N
	int single_end;   // only text 0 holds records (ReadQC without -in2): read 2 of every pair is empty
	int validate;     // FastqEntry::validate checks on the header lines
	int fixed_trim;   // FastqTrim (src/FastqTrim/main.cpp:47-77) on the reads of text 0
	int ft_start, ft_end, ft_len, ft_max_len;
	spg_result* res_w; // fixed_trim: the records this mode writes itself (same buffer as res)
};

__device__ __forceinline__ uint32_t fq_newline_flags(uint32_t w) // bit 7 of every byte that equals '\n' (exact, no borrow artefacts)
{
	const uint32_t x = w ^ 0x0A0A0A0Au;
	const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
	return ~(t | x | 0x7F7F7F7Fu);
}

// newline flags of the 64 bytes a thread owns: bit i of the result = byte i is '\n' (bytes at or beyond `bytes` never are)
__device__ __forceinline__ unsigned long long fq_thread_flags(const uint8_t* text, uint32_t bytes, uint32_t base)
{
	unsigned long long flags = 0;
	if (base >= bytes) return 0;
#pragma unroll
	for (int v = 0; v < 4; ++v)
	{
		const uint32_t off = base + 16u * v;
		if (off >= bytes) break;
		const uint4 q = *reinterpret_cast<const uint4*>(text + off); // the buffers are padded to a multiple of 16 bytes
		const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
		for (int k = 0; k < 4; ++k)
		{
			uint32_t f = fq_newline_flags(w[k]); // bits 7, 15, 23, 31
			// gather to 4 bits: byte j -> bit j
			f = ((f >> 7) & 1u) | ((f >> 14) & 2u) | ((f >> 21) & 4u) | ((f >> 28) & 8u);
			flags |= (unsigned long long)f << (16 * v + 4 * k);
		}
	}
	const uint32_t left = bytes - base;
	if (left < 64u) flags &= (1ull << left) - 1ull;
	return flags;
}

// grid: (max blocks of the two texts, 2)
__global__ void __launch_bounds__(kFqBlockThreads) fq_count_newlines(const __grid_constant__ FqArgs A)
{
	const int f = blockIdx.y;
	if ((int)blockIdx.x >= A.n_blocks[f]) return;
	const uint32_t base = (uint32_t)blockIdx.x * kFqBlockBytes + (uint32_t)threadIdx.x * kFqBytesPerThread;
	int c = __popcll(fq_thread_flags(A.text[f], A.bytes[f], base));
	c = __reduce_add_sync(0xffffffffu, c);
	__shared__ int s[kFqBlockThreads / 32];
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		int t = 0;
		for (int i = 0; i < kFqBlockThreads / 32; ++i) t += s[i];
		A.block_counts[f][blockIdx.x] = (uint32_t)t;
	}
}

// grid: 2 CTAs (one per text) of 1024 threads: exclusive scan of the block counts in place; the plan gets the line totals
__global__ void __launch_bounds__(1024) fq_scan_counts(const __grid_constant__ FqArgs A)
{
	const int f = blockIdx.x;
	const int n = A.n_blocks[f];
	__shared__ uint32_t warp_sums[32];
	__shared__ uint32_t carry;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int base = 0; base < n; base += 1024)
	{
		const int i = base + (int)threadIdx.x;
		const uint32_t v = i < n ? A.block_counts[f][i] : 0u;
		uint32_t x = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
			if ((int)(threadIdx.x & 31) >= d) x += y;
		}
		if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
		__syncthreads();
		if (threadIdx.x < 32)
		{
			uint32_t w = warp_sums[threadIdx.x];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1)
			{
				const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
				if ((int)threadIdx.x >= d) w += y;
			}
			warp_sums[threadIdx.x] = w; // inclusive
		}
		__syncthreads();
		const uint32_t before = carry + (threadIdx.x >= 32 ? warp_sums[(threadIdx.x >> 5) - 1] : 0u) + x - v;
		if (i < n) A.block_counts[f][i] = before;
		__syncthreads();
		if (threadIdx.x == 1023) carry = before + v;
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		// complete lines, plus the unterminated last line of a file's final chunk (readLine returns it like any other line)
		long long lines = carry;
		const uint32_t bytes = A.bytes[f];
		if (A.final_[f] && bytes > 0 && A.text[f][bytes - 1] != '\n') lines += 1;
		A.plan->lines[f] = (int)min(lines, (long long)A.nl_cap);
	}
}

// same grid as fq_count_newlines: positions of the newlines in rank order
__global__ void __launch_bounds__(kFqBlockThreads) fq_scatter_newlines(const __grid_constant__ FqArgs A)
{
	const int f = blockIdx.y;
	if ((int)blockIdx.x >= A.n_blocks[f]) return;
	const uint32_t base = (uint32_t)blockIdx.x * kFqBlockBytes + (uint32_t)threadIdx.x * kFqBytesPerThread;
	unsigned long long flags = fq_thread_flags(A.text[f], A.bytes[f], base);
	const int c = __popcll(flags);
	// exclusive scan of the per-thread counts inside the CTA
	int x = c;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const int y = __shfl_up_sync(0xffffffffu, x, d);
		if ((int)(threadIdx.x & 31) >= d) x += y;
	}
	__shared__ int s[kFqBlockThreads / 32];
	if ((threadIdx.x & 31) == 31) s[threadIdx.x >> 5] = x;
	__syncthreads();
	int before = x - c;
	for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += s[w];
	uint32_t rank = A.block_counts[f][blockIdx.x] + (uint32_t)before;
	uint32_t* nl = A.nl[f];
	while (flags)
	{
		const int b = __ffsll((long long)flags) - 1;
		flags &= flags - 1;
		if (rank < (uint32_t)A.nl_cap) nl[rank] = base + (uint32_t)b;
		++rank;
	}
	// the virtual line end of an unterminated final line
	if (blockIdx.x == 0 && threadIdx.x == 0)
	{
		const uint32_t bytes = A.bytes[f];
		if (A.final_[f] && bytes > 0 && A.text[f][bytes - 1] != '\n')
		{
			// its rank is the number of real newlines = exclusive offset of the last block + that block's count; recomputed here from
			// the last block to stay independent of the other CTAs' progress
			const int lb = A.n_blocks[f] - 1;
			uint32_t cnt = 0;
			const uint32_t lbase = (uint32_t)lb * kFqBlockBytes;
			for (uint32_t i = lbase; i < bytes; ++i) cnt += A.text[f][i] == '\n';
			const uint32_t r = A.block_counts[f][lb] + cnt;
			if (r < (uint32_t)A.nl_cap) nl[r] = bytes;
		}
	}
}

// one thread: how many records each text holds, how many pairs this batch gets, how many bytes they cover
__global__ void fq_plan(const __grid_constant__ FqArgs A)
{
	if (threadIdx.x != 0 || blockIdx.x != 0) return;
	FqPlan& P = *A.plan;
	int rec[2];
	for (int f = 0; f < 2; ++f)
	{
		const int lines = P.lines[f];
		// a final chunk: the last record may lack lines (readLine returns empty strings at the end of the file) -- ceil
		rec[f] = A.final_[f] ? (lines + 3) / 4 : lines / 4;
		P.records[f] = rec[f];
	}
	int n = min(A.single_end ? rec[0] : min(rec[0], rec[1]), A.max_pairs);
	P.n_pairs = n;
	for (int f = 0; f < 2; ++f)
	{
		const long long last_line = 4ll * n - 1; // its end closes the n-th record
		uint32_t c;
		if (n == 0) c = 0;
		else if (f == 1 && A.single_end) c = 0;
		else if (last_line < P.lines[f]) c = min(A.nl[f][last_line] + 1u, A.bytes[f]);
		else c = A.bytes[f]; // the final, incomplete record
		P.consumed[f] = c;
	}
	for (int k = 0; k < 4; ++k) P.out_bytes[k] = 0;
	P.err_pair = 0x7fffffff;
	P.max_len = 0;
	P.acons_error = 0;
}

struct FqLine
{
	uint32_t s, e; // [s, e) without line ending
};

// Copies of the row / text segments of a pair, warp-cooperative, in two phases: first every load of the pair is issued, then the
// stores follow, so a pair costs one trip to memory (the byte-wise loop of round 1 waited for memory once per 32 bytes of each segment:
// the two copy kernels were bound by that chain of latencies at full occupancy).
//
// FqTwo: two segments of the same length n (the bases and the qualities of a read). Four bytes at a time: lane w forms the w-th aligned
// destination word from the two aligned source words that hold its bytes (funnel shift; the second one comes from the neighbouring
// lane); two words per lane in flight cover reads of up to 256 bases, longer ones finish in a plain loop. The up to three bytes in front
// of and behind the words go one per lane (lanes 0-2, 4-6). A source word is only read if it holds a byte of the segment, and all
// buffers are multiples of 16 bytes, so nothing outside them is touched.
struct FqTwo
{
	uint32_t nw[2], sh[2], ei[2], eb[2], lo0[2], lo1[2], ex[2];
	const uint32_t* lp[2]; // aligned source word of this lane
	uint32_t* sp[2];       // destination word of this lane
};
__device__ __forceinline__ void fq_two_load(FqTwo& T, uint8_t* d0, const uint8_t* s0, uint8_t* d1, const uint8_t* s1, uint32_t n, int lane)
{
#pragma unroll
	for (int k = 0; k < 2; ++k)
	{
		uint8_t* dst = k ? d1 : d0;
		const uint8_t* src = k ? s1 : s0;
		const uint32_t head = min(n, (0u - (uint32_t)(uintptr_t)dst) & 3u);
		const uint32_t nw = (n - head) >> 2;
		const uint32_t mis = ((uint32_t)(uintptr_t)src + head) & 3u;
		const uint32_t nld = nw + (mis ? 1u : 0u); // source words that hold bytes of the destination words
		T.nw[k] = nw;
		T.sh[k] = 8u * mis;
		T.lp[k] = reinterpret_cast<const uint32_t*>(src + (int)(head - mis)) + lane;
		T.sp[k] = reinterpret_cast<uint32_t*>(dst + head) + lane;
		const uint32_t done = head + 4u * nw, t = (uint32_t)lane - 4u;
		T.ei[k] = (uint32_t)lane < head ? (uint32_t)lane : (t < n - done ? done + t : 0xFFFFFFFFu);
		T.eb[k] = T.lo0[k] = T.lo1[k] = T.ex[k] = 0;
		if (T.ei[k] != 0xFFFFFFFFu) T.eb[k] = src[T.ei[k]];
		if ((uint32_t)lane < nld) T.lo0[k] = T.lp[k][0];
		if ((uint32_t)lane + 32u < nld) T.lo1[k] = T.lp[k][32];
		if (lane == 31 && 64u < nld) T.ex[k] = T.lp[k][33];
	}
}
__device__ __forceinline__ void fq_two_store(const FqTwo& T, uint8_t* d0, const uint8_t* s0, uint8_t* d1, const uint8_t* s1, int lane)
{
#pragma unroll
	for (int k = 0; k < 2; ++k)
	{
		const uint32_t n0 = __shfl_down_sync(0xffffffffu, T.lo0[k], 1), n1 = __shfl_down_sync(0xffffffffu, T.lo1[k], 1);
		const uint32_t first1 = __shfl_sync(0xffffffffu, T.lo1[k], 0);
		const uint32_t hi0 = lane == 31 ? first1 : n0, hi1 = lane == 31 ? T.ex[k] : n1;
		if ((uint32_t)lane < T.nw[k]) T.sp[k][0] = __funnelshift_r(T.lo0[k], hi0, T.sh[k]);
		if ((uint32_t)lane + 32u < T.nw[k]) T.sp[k][32] = __funnelshift_r(T.lo1[k], hi1, T.sh[k]);
		if (T.ei[k] != 0xFFFFFFFFu) (k ? d1 : d0)[T.ei[k]] = (uint8_t)T.eb[k];
		for (uint32_t w = 64u + (uint32_t)lane; w < T.nw[k]; w += 32u) // reads of more than 256 bases
		{
			const uint32_t lo = T.lp[k][w - (uint32_t)lane], hi = T.sh[k] ? T.lp[k][w - (uint32_t)lane + 1u] : 0u;
			T.sp[k][w - (uint32_t)lane] = __funnelshift_r(lo, hi, T.sh[k]);
		}
	}
}

// FqText: a short segment (a header line) byte-wise: the first 64 bytes in flight, longer ones finish in a plain loop
struct FqText
{
	uint32_t b0, b1;
};
__device__ __forceinline__ void fq_text_load(FqText& T, const uint8_t* src, uint32_t n, int lane)
{
	T.b0 = T.b1 = 0;
	if ((uint32_t)lane < n) T.b0 = src[lane];
	if ((uint32_t)lane + 32u < n) T.b1 = src[lane + 32];
}
__device__ __forceinline__ void fq_text_store(const FqText& T, uint8_t* dst, const uint8_t* src, uint32_t n, int lane)
{
	if ((uint32_t)lane < n) dst[lane] = (uint8_t)T.b0;
	if ((uint32_t)lane + 32u < n) dst[lane + 32] = (uint8_t)T.b1;
	for (uint32_t i = 64u + (uint32_t)lane; i < n; i += 32u) dst[i] = src[i];
}

// first ' ' of a line (or its length), warp-cooperative; `c0` is this lane's byte of the first 32 (0x100 beyond the line)
__device__ __forceinline__ uint32_t fq_token_len(const uint8_t* t, FqLine L, uint32_t c0, int lane)
{
	const uint32_t len = L.e - L.s;
	for (uint32_t base = 0; base < len; base += 32)
	{
		const uint32_t i = base + (uint32_t)lane;
		const uint32_t c = base == 0 ? c0 : (i < len ? (uint32_t)t[L.s + i] : 0x100u);
		const uint32_t m = __ballot_sync(0xffffffffu, c == (uint32_t)' ');
		if (m) return base + (uint32_t)(__ffs((int)m) - 1);
	}
	return len;
}

// bounds of line k of text f as readLine delivers it: from the end of the line before to its own end, trailing '\r' chopped; lines beyond
// the last one are empty. `a`, `b` = nl[k-1], nl[k] (loaded by the caller).
__device__ __forceinline__ FqLine fq_line_of(const FqArgs& A, int f, long long k, int lines, uint32_t a, uint32_t b)
{
	FqLine L;
	if (k >= lines)
	{
		L.s = L.e = A.bytes[f];
		return L;
	}
	L.s = k == 0 ? 0u : a + 1u;
	L.e = b;
	const uint8_t* t = A.text[f];
	while (L.e > L.s && t[L.e - 1] == '\r') --L.e;
	return L;
}

// one warp per pair: rows, lengths, header locations, header check. Three trips to memory per pair: the line index (lanes 0-7 load the
// entries of the eight lines, one pair ahead), the bytes in front of the line ends + the first 32 bytes of the two headers, then all
// four rows at once.
__global__ void __launch_bounds__(256, 3) fq_pack(const __grid_constant__ FqArgs A)
{
	const int lane = threadIdx.x & 31;
	const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	const int n = A.plan->n_pairs;
	const int nfiles = A.single_end ? 1 : 2;
	const int lf = (lane >> 2) & 1, li = lane & 3; // text and line of the record this lane looks up (lanes 0-7)
	const int my_lines = A.plan->lines[lf];
	const bool looks_up = lane < 4 * nfiles;
	int max_len = 0;
	uint32_t na = 0, nb = 0; // nl[k-1], nl[k] of this lane's line of the current pair
	auto load_index = [&](long long p) {
		const long long k = 4 * p + li;
		na = nb = 0;
		if (looks_up && k < my_lines)
		{
			if (k > 0) na = A.nl[lf][k - 1];
			nb = A.nl[lf][k];
		}
	};
	if (gwarp < n) load_index(gwarp);
	for (long long p = gwarp; p < n; p += nwarps)
	{
		const uint32_t ca = na, cb = nb;
		if (p + nwarps < n) load_index(p + nwarps);
		// the first 32 bytes behind the start of each header line (masked with the line's length below): issued together with the '\r' checks
		uint32_t hc[2] = {0x100u, 0x100u};
#pragma unroll
		for (int f = 0; f < 2; ++f)
		{
			const long long k0 = 4 * p;
			const uint32_t a0 = __shfl_sync(0xffffffffu, ca, 4 * f);
			const uint32_t hstart = k0 >= A.plan->lines[f] ? A.bytes[f] : (k0 == 0 ? 0u : a0 + 1u);
			if (f < nfiles && hstart + (uint32_t)lane < A.bytes[f]) hc[f] = A.text[f][hstart + (uint32_t)lane];
		}
		FqLine mine;
		mine.s = mine.e = 0;
		if (looks_up) mine = fq_line_of(A, lf, 4 * p + li, my_lines, ca, cb);
		FqLine ln[2][4];
#pragma unroll
		for (int f = 0; f < 2; ++f)
#pragma unroll
			for (int i = 0; i < 4; ++i)
			{
				ln[f][i].s = __shfl_sync(0xffffffffu, mine.s, 4 * f + i);
				ln[f][i].e = __shfl_sync(0xffffffffu, mine.e, 4 * f + i);
			}
		uint8_t st = FQ_OK;
		FqTwo two[2];
		uint8_t* rdst[2][2];
		const uint8_t* rsrc[2][2];
#pragma unroll
		for (int f = 0; f < 2; ++f)
		{
			rdst[f][0] = rdst[f][1] = nullptr;
			rsrc[f][0] = rsrc[f][1] = nullptr;
			if (f == 1 && A.single_end) // no mate: an empty read 2
			{
				if (lane == 0)
				{
					A.len[1][p] = 0;
					FqRec r;
					r.hs = r.hl = r.h2s = r.h2l = 0;
					A.rec[1][p] = r;
				}
				continue;
			}
			const uint8_t* t = A.text[f];
			const FqLine hdr = ln[f][0], b = ln[f][1], h2 = ln[f][2], q = ln[f][3];
			const uint32_t lb = b.e - b.s, lq = q.e - q.s;
			max_len = max(max_len, (int)min(max(lb, lq), 0x7fffffffu));
			uint32_t len = lb;
			if (A.validate && st == FQ_OK) // FastqEntry::validate, in its order (FastqFileStream.cpp:7-18)
			{
				if (hdr.e == hdr.s || t[hdr.s] != '@') st = FQ_BAD_HEADER;
				else if (h2.e == h2.s || t[h2.s] != '+') st = FQ_BAD_HEADER2;
			}
			if (lb != lq && st == FQ_OK) st = FQ_LENGTH_MISMATCH;
			if ((lb > (uint32_t)A.stride || lq > (uint32_t)A.stride || lb >= (uint32_t)SPG_MAXLEN) && st == FQ_OK) st = FQ_TOO_LONG;
			if (lb > (uint32_t)A.stride || lq > (uint32_t)A.stride || lb >= (uint32_t)SPG_MAXLEN || lb != lq) len = 0; // keeps the kernels behind inside the rows
			rdst[f][0] = A.rows[2 * f] + (size_t)p * A.stride;
			rdst[f][1] = A.rows[2 * f + 1] + (size_t)p * A.stride;
			rsrc[f][0] = t + b.s;
			rsrc[f][1] = t + q.s;
			fq_two_load(two[f], rdst[f][0], rsrc[f][0], rdst[f][1], rsrc[f][1], len, lane);
			if (lane == 0)
			{
				A.len[f][p] = (uint16_t)len;
				FqRec r;
				r.hs = hdr.s;
				r.hl = hdr.e - hdr.s;
				r.h2s = h2.s;
				r.h2l = h2.e - h2.s;
				A.rec[f][p] = r;
			}
		}
		fq_two_store(two[0], rdst[0][0], rsrc[0][0], rdst[0][1], rsrc[0][1], lane);
		if (!A.single_end) fq_two_store(two[1], rdst[1][0], rsrc[1][0], rdst[1][1], rsrc[1][1], lane);
		// headers must match up to the first space, "/1" and "/2" aside (AnalysisWorker.cpp:110-120); ReadQC does not compare them
		if (!A.single_end && !A.validate)
		{
			const uint8_t* t1 = A.text[0];
			const uint8_t* t2 = A.text[1];
			const FqLine h1 = ln[0][0], h2 = ln[1][0];
			const uint32_t c1 = (uint32_t)lane < h1.e - h1.s ? hc[0] : 0x100u, c2 = (uint32_t)lane < h2.e - h2.s ? hc[1] : 0x100u;
			uint32_t n1 = fq_token_len(t1, h1, c1, lane), n2 = fq_token_len(t2, h2, c2, lane);
			if (n1 >= 2 && n2 >= 2)
			{
				// the last two bytes of both tokens: from the preloaded bytes where they are among the first 32
				auto hb = [&](const uint8_t* t, FqLine L, uint32_t c0, uint32_t idx) -> uint32_t {
					return idx < 32u ? __shfl_sync(0xffffffffu, c0, (int)idx) : (uint32_t)t[L.s + idx];
				};
				const uint32_t e1 = hb(t1, h1, c1, n1 - 2), e2 = hb(t1, h1, c1, n1 - 1), e3 = hb(t2, h2, c2, n2 - 2), e4 = hb(t2, h2, c2, n2 - 1);
				if (e1 == '/' && e2 == '1' && e3 == '/' && e4 == '2')
				{
					n1 -= 2;
					n2 -= 2;
				}
			}
			bool diff = n1 != n2;
			if (!diff)
			{
				for (uint32_t base = 0; base < n1 && !diff; base += 32)
				{
					const uint32_t i = base + (uint32_t)lane;
					const uint32_t x1 = base == 0 ? c1 : (uint32_t)(i < n1 ? t1[h1.s + i] : 0);
					const uint32_t x2 = base == 0 ? c2 : (uint32_t)(i < n1 ? t2[h2.s + i] : 0);
					diff = __any_sync(0xffffffffu, i < n1 && x1 != x2);
				}
			}
			if (diff) st = FQ_HEADER_MISMATCH; // the first check of the worker: wins over the length checks
		}
		if (lane == 0)
		{
			A.fstat[p] = st;
			if (st != FQ_OK) atomicMin(&A.plan->err_pair, (int)p);
		}
	}
	max_len = __reduce_max_sync(0xffffffffu, max_len);
	if (lane == 0 && max_len > 0) atomicMax(&A.plan->max_len, max_len);
}

// FastqTrim (src/FastqTrim/main.cpp:47-77): one thread per read; len1 = bases kept, best_offset = first base kept, SPG_F_DROPPED if the
// read is not written
__global__ void __launch_bounds__(256) fq_fixed_trim(const __grid_constant__ FqArgs A)
{
	const int n = A.plan->n_pairs;
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
	{
		const int len = A.len[0][p];
		int off = 0, keep_n = len;
		bool keep = true;
		if (!(A.ft_max_len > 0 && len >= A.ft_max_len)) // reads of max_len bases and more pass unchanged (:52-56)
		{
			if (A.ft_start > 0 || A.ft_end > 0) // :58-64
			{
				if (len <= A.ft_start + A.ft_end) keep = false;
				else
				{
					off = A.ft_start;
					keep_n = len - A.ft_start - A.ft_end;
				}
			}
			if (keep && A.ft_len > 0 && keep_n > A.ft_len) keep_n = A.ft_len; // :66-70
		}
		spg_result r;
		r.len1 = (uint16_t)(keep ? keep_n : 0);
		r.len2 = 0;
		r.best_offset = (int16_t)off;
		r.flags = keep ? 0 : (uint8_t)SPG_F_DROPPED;
		r.status = 0;
		A.res_w[p] = r;
	}
}

// ---- output --------------------------------------------------------------------------------------------------------------------------------
// bytes a pair contributes to the four streams
__device__ __forceinline__ void fq_pair_sizes(const FqArgs& A, int p, uint32_t sz[4], int& l1, int& l2)
{
	const spg_result r = A.res[p];
	l1 = r.len1;
	l2 = r.len2;
	if (A.fixed_trim) // one stream, every read that was not dropped
	{
		const FqRec a = A.rec[0][p];
		sz[0] = (r.flags & SPG_F_DROPPED) ? 0u : a.hl + a.h2l + 2u * (uint32_t)l1 + 4u;
		sz[1] = sz[2] = sz[3] = 0;
		return;
	}
	const bool ok1 = l1 >= A.min_len, ok2 = l2 >= A.min_len;
	const FqRec a = A.rec[0][p], b = A.rec[1][p];
	const uint32_t s1 = a.hl + a.h2l + 2u * (uint32_t)l1 + 4u;
	const uint32_t s2 = b.hl + b.h2l + 2u * (uint32_t)l2 + 4u;
	sz[0] = sz[1] = sz[2] = sz[3] = 0;
	if (ok1 && ok2)
	{
		sz[0] = s1;
		sz[1] = s2;
	}
	else if (A.singles && ok1) sz[2] = s1;
	else if (A.singles && ok2) sz[3] = s2;
}

// exclusive scan of four values per thread over a CTA of kFqOutPairs threads; returns the CTA totals in tot
__device__ __forceinline__ void fq_block_scan4(uint32_t v[4], uint32_t tot[4])
{
	__shared__ uint32_t ws[4][kFqOutPairs / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc[4];
#pragma unroll
	for (int k = 0; k < 4; ++k)
	{
		uint32_t x = v[k];
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
			if (lane >= d) x += y;
		}
		inc[k] = x;
		if (lane == 31) ws[k][warp] = x;
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < 4; ++k)
	{
		uint32_t before = 0, total = 0;
		for (int w = 0; w < kFqOutPairs / 32; ++w)
		{
			const uint32_t s = ws[k][w];
			if (w < warp) before += s;
			total += s;
		}
		v[k] = before + inc[k] - v[k];
		tot[k] = total;
	}
	__syncthreads();
}

__global__ void __launch_bounds__(kFqOutPairs) fq_out_sizes(const __grid_constant__ FqArgs A)
{
	const int n = A.plan->n_pairs;
	const int first = blockIdx.x * kFqOutPairs;
	if (first >= n) return;
	const int p = first + (int)threadIdx.x;
	uint32_t sz[4] = {0, 0, 0, 0}, tot[4];
	int l1, l2;
	if (p < n) fq_pair_sizes(A, p, sz, l1, l2);
	fq_block_scan4(sz, tot);
	if (threadIdx.x < 4) A.out_block[threadIdx.x][blockIdx.x] = tot[threadIdx.x];
	// summary counters (OutputWorker.cpp:59-77): per-CTA histograms in shared memory, one atomic per touched bin into the chunk's totals
	if (A.stats != nullptr)
	{
		__shared__ uint32_t h_rem[SPG_MAXLEN], h_trim[SPG_MAXLEN], cnt[5];
		for (int i = threadIdx.x; i < SPG_MAXLEN; i += kFqOutPairs) h_rem[i] = h_trim[i] = 0;
		if (threadIdx.x < 5) cnt[threadIdx.x] = 0;
		__syncthreads();
		if (p < n)
		{
			const spg_result r = A.res[p];
			const int o1 = A.len[0][p], o2 = A.len[1][p];
			const bool ok1 = l1 >= A.min_len, ok2 = l2 >= A.min_len;
			const uint32_t removed = (ok1 && ok2) ? 0u : ((A.singles && (ok1 || ok2)) ? 1u : 2u);
			const uint32_t tq = ((r.flags & SPG_F_Q1) ? 1u : 0u) + ((r.flags & SPG_F_Q2) ? 1u : 0u);
			const uint32_t tn = ((r.flags & SPG_F_N1) ? 1u : 0u) + ((r.flags & SPG_F_N2) ? 1u : 0u);
			if (r.flags & SPG_F_INSERT) atomicAdd(&cnt[0], 2u);
			if (r.flags & SPG_F_ADAPTER) atomicAdd(&cnt[1], 2u);
			if (tq) atomicAdd(&cnt[2], tq);
			if (tn) atomicAdd(&cnt[3], tn);
			if (removed) atomicAdd(&cnt[4], removed);
			atomicAdd(&h_rem[min(l1, SPG_MAXLEN - 1)], 1u);
			atomicAdd(&h_rem[min(l2, SPG_MAXLEN - 1)], 1u);
			if (o1 > l1) atomicAdd(&h_trim[min(o1, SPG_MAXLEN - 1)], (uint32_t)(o1 - l1));
			if (o2 > l2) atomicAdd(&h_trim[min(o2, SPG_MAXLEN - 1)], (uint32_t)(o2 - l2));
		}
		__syncthreads();
		if (threadIdx.x < 5 && cnt[threadIdx.x]) atomicAdd(&A.stats[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
		for (int i = threadIdx.x; i < SPG_MAXLEN; i += kFqOutPairs)
		{
			if (h_rem[i]) atomicAdd(&A.stats[5 + i], (unsigned long long)h_rem[i]);
			if (h_trim[i]) atomicAdd(&A.stats[5 + SPG_MAXLEN + i], (unsigned long long)h_trim[i]);
		}
	}
}

// one CTA of 1024 threads: exclusive scan of the per-CTA byte counts of the four streams; totals into the plan
__global__ void __launch_bounds__(1024) fq_out_scan(const __grid_constant__ FqArgs A)
{
	const int n = A.plan->n_pairs;
	const int nb = (n + kFqOutPairs - 1) / kFqOutPairs;
	const int k = threadIdx.x >> 8, t = threadIdx.x & 255; // 256 threads per stream
	__shared__ uint32_t part[4][256];
	// thread t of stream k owns the blocks [t*per, (t+1)*per)
	const int per = (nb + 255) / 256;
	uint32_t sum = 0;
	for (int i = t * per; i < min(nb, (t + 1) * per); ++i) sum += A.out_block[k][i];
	part[k][t] = sum;
	__syncthreads();
	if (t == 0)
	{
		uint32_t run = 0;
		for (int i = 0; i < 256; ++i)
		{
			const uint32_t v = part[k][i];
			part[k][i] = run;
			run += v;
		}
		A.plan->out_bytes[k] = run;
	}
	__syncthreads();
	uint32_t run = part[k][t];
	for (int i = t * per; i < min(nb, (t + 1) * per); ++i)
	{
		const uint32_t v = A.out_block[k][i];
		A.out_block[k][i] = run;
		run += v;
	}
}

// one record "header\nbases\nheader2\nquals\n" at dst, in the two phases of FqTwo / FqText; lanes 28-31 write the four line ends
struct FqRecord
{
	FqText h, h2;
	FqTwo rows;
};
__device__ __forceinline__ void fq_record_load(FqRecord& R, uint8_t* dst, const uint8_t* text, FqRec r, const uint8_t* bases, const uint8_t* quals, uint32_t len, int lane)
{
	fq_text_load(R.h, text + r.hs, r.hl, lane);
	fq_text_load(R.h2, text + r.h2s, r.h2l, lane);
	fq_two_load(R.rows, dst + r.hl + 1u, bases, dst + r.hl + len + r.h2l + 3u, quals, len, lane);
}
__device__ __forceinline__ void fq_record_store(const FqRecord& R, uint8_t* dst, const uint8_t* text, FqRec r, const uint8_t* bases, const uint8_t* quals, uint32_t len, int lane)
{
	fq_text_store(R.h, dst, text + r.hs, r.hl, lane);
	fq_text_store(R.h2, dst + r.hl + len + 2u, text + r.h2s, r.h2l, lane);
	fq_two_store(R.rows, dst + r.hl + 1u, bases, dst + r.hl + len + r.h2l + 3u, quals, lane);
	if (lane >= 28)
	{
		const uint32_t at = lane == 28 ? r.hl : lane == 29 ? r.hl + len + 1u : lane == 30 ? r.hl + len + r.h2l + 2u : r.hl + 2u * len + r.h2l + 3u;
		dst[at] = '\n';
	}
}

__device__ __forceinline__ int fq_base_index(uint8_t c) // Pileup::inc (src/cppNGS/Pileup.cpp:17-32); -1: ignored, -2: unknown
{
	switch (c)
	{
		case 'A': case 'a': return 0;
		case 'C': case 'c': return 1;
		case 'G': case 'g': return 2;
		case 'T': case 't': return 3;
		case 'N': case 'n': return 4;
		case '-': case '~': return -1;
		default: return -2;
	}
}

__global__ void __launch_bounds__(kFqOutPairs, 2) fq_out_write(const __grid_constant__ FqArgs A)
{
	const int n = A.plan->n_pairs;
	const int first = blockIdx.x * kFqOutPairs;
	if (first >= n) return;
	__shared__ uint32_t offs[4][kFqOutPairs];
	__shared__ unsigned int s_acons[2 * 40 * 5];
	for (int i = threadIdx.x; i < 2 * 40 * 5; i += kFqOutPairs) s_acons[i] = 0;
	{
		const int p = first + (int)threadIdx.x;
		uint32_t sz[4] = {0, 0, 0, 0}, tot[4];
		int l1, l2;
		if (p < n) fq_pair_sizes(A, p, sz, l1, l2);
		fq_block_scan4(sz, tot);
#pragma unroll
		for (int k = 0; k < 4; ++k) offs[k][threadIdx.x] = A.out_block[k][blockIdx.x] + sz[k];
	}
	__syncthreads();
	// a warp writes 32 consecutive pairs (neighbouring rows share sectors); the result record and the header locations of a pair are
	// loaded one pair ahead, the loads of its two records are issued before the first store (fq_record_load / fq_record_store)
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int j0 = warp * 32;
	spg_result r_nx;
	FqRec a_nx, b_nx;
	r_nx.len1 = r_nx.len2 = 0;
	r_nx.best_offset = -1;
	r_nx.flags = r_nx.status = 0;
	a_nx.hs = a_nx.hl = a_nx.h2s = a_nx.h2l = 0;
	b_nx = a_nx;
	if (first + j0 < n)
	{
		r_nx = A.res[first + j0];
		a_nx = A.rec[0][first + j0];
		b_nx = A.rec[1][first + j0];
	}
	for (int jj = 0; jj < 32; ++jj)
	{
		const int j = j0 + jj, p = first + j;
		if (p >= n) break;
		const spg_result r = r_nx;
		const FqRec ra = a_nx, rb = b_nx;
		if (jj + 1 < 32 && p + 1 < n)
		{
			r_nx = A.res[p + 1];
			a_nx = A.rec[0][p + 1];
			b_nx = A.rec[1][p + 1];
		}
		const int l1 = r.len1, l2 = r.len2;
		const bool ok1 = l1 >= A.min_len, ok2 = l2 >= A.min_len;
		const uint8_t* b1 = A.rows[0] + (size_t)p * A.stride;
		const uint8_t* q1 = A.rows[1] + (size_t)p * A.stride;
		const uint8_t* b2 = A.rows[2] + (size_t)p * A.stride;
		const uint8_t* q2 = A.rows[3] + (size_t)p * A.stride;
		if (A.fixed_trim)
		{
			if (!(r.flags & SPG_F_DROPPED))
			{
				FqRecord R;
				uint8_t* d = A.out[0] + offs[0][j];
				fq_record_load(R, d, A.text[0], ra, b1 + r.best_offset, q1 + r.best_offset, (uint32_t)l1, lane);
				fq_record_store(R, d, A.text[0], ra, b1 + r.best_offset, q1 + r.best_offset, (uint32_t)l1, lane);
			}
			continue;
		}
		if (ok1 && ok2)
		{
			FqRecord R1, R2;
			uint8_t* d1 = A.out[0] + offs[0][j];
			uint8_t* d2 = A.out[1] + offs[1][j];
			fq_record_load(R1, d1, A.text[0], ra, b1, q1, (uint32_t)l1, lane);
			fq_record_load(R2, d2, A.text[1], rb, b2, q2, (uint32_t)l2, lane);
			fq_record_store(R1, d1, A.text[0], ra, b1, q1, (uint32_t)l1, lane);
			fq_record_store(R2, d2, A.text[1], rb, b2, q2, (uint32_t)l2, lane);
		}
		else if (A.singles && (ok1 || ok2))
		{
			FqRecord R;
			uint8_t* d = ok1 ? A.out[2] + offs[2][j] : A.out[3] + offs[3][j];
			const uint8_t* text = ok1 ? A.text[0] : A.text[1];
			const FqRec rr = ok1 ? ra : rb;
			const uint8_t* bb = ok1 ? b1 : b2;
			const uint8_t* qq = ok1 ? q1 : q2;
			const uint32_t ll = (uint32_t)(ok1 ? l1 : l2);
			fq_record_load(R, d, text, rr, bb, qq, ll, lane);
			fq_record_store(R, d, text, rr, bb, qq, ll, lane);
		}

		// consensus of the adapter bases behind the insert, from the untrimmed reads (AnalysisWorker.cpp:279-290); -ec never edits them
		if ((r.flags & SPG_F_INSERT) && r.status == 0)
		{
			const int o1 = A.len[0][p], o2 = A.len[1][p];
			const int new_length = o2 - r.best_offset;
			for (int i = lane; i < 80; i += 32)
			{
				const int rd = i / 40, k = i % 40;
				int idx = -1;
				if (rd == 0 && new_length + k < o1) idx = fq_base_index(b1[new_length + k]);
				if (rd == 1 && k < r.best_offset && new_length + k < o2) idx = fq_base_index(b2[new_length + k]);
				if (idx >= 0) atomicAdd(&s_acons[(rd * 40 + k) * 5 + idx], 1u);
				else if (idx == -2) A.plan->acons_error = 1;
			}
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 2 * 40 * 5; i += kFqOutPairs)
		if (s_acons[i]) atomicAdd(&A.acons[i], (unsigned long long)s_acons[i]);
}

} // namespace spg
