// spg_qc.cuh -- raw-read statistics (-qc) of a batch: the accumulators of StatisticsReads::update(FastqEntry, direction) of
// imgag/ngs-bits (src/cppNGS/StatisticsReads.cpp:26-81) that the paired-end qcML report needs, as one reduction kernel.
//
// Same skeleton as the trimming kernel (spg_kernel.cuh): persistent CTAs, one producer warp that streams tiles of whole rows into a
// shared-memory ring with 1-D bulk copies (TMA), CW consumer warps, one warp per read pair, lane l owns the cycles l, 32+l, ...
// The inner loop has no atomics: per base one table lookup that yields a one-hot 6-bit field (five of them packed in a register
// per owned cycle) and one lookup for the quality (value, >=20 and >=30 flags in separate bit fields, so that one add per base
// accumulates the read's quality sum and both counts). Every 31 pairs a warp spills its packed counters into the CTA's table in
// shared memory; per read a warp reduction gives the mean quality. At the end each CTA adds its table to the device-wide 64-bit
// accumulators.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/seqpurge_b200.h"
#include "spg_kernel.cuh"

namespace spg
{

// layout of the device-wide accumulator array (unsigned long long)
constexpr int kQcReadsF = 0, kQcReadsR = 1, kQcBases = 2, kQcReadQ20 = 3, kQcBaseQ20 = 4, kQcBaseQ30 = 5, kQcErrors = 6;
constexpr int kQcLen = 8;                          // [SPG_MAXLEN] read length histogram
constexpr int kQcPile = kQcLen + SPG_MAXLEN;       // [SPG_MAXLEN][5]
constexpr int kQcQf = kQcPile + 5 * SPG_MAXLEN;    // [SPG_MAXLEN] quality sums, forward reads
constexpr int kQcQr = kQcQf + SPG_MAXLEN;          // [SPG_MAXLEN] reverse reads
constexpr int kQcBaseQual = kQcQr + SPG_MAXLEN;    // [100] bases by quality
constexpr int kQcReadQual = kQcBaseQual + 100;     // [100] reads by rounded mean quality
constexpr int kQcDistF = kQcReadQual + 100;        // [60] Histogram(0,60,1) of the mean quality, forward reads
constexpr int kQcDistR = kQcDistF + 60;            // [60] reverse reads
constexpr int kQcWords = kQcDistR + 60;

struct QcArgs
{
	const uint8_t* b1;
	const uint8_t* q1;
	const uint8_t* b2;
	const uint8_t* q2;
	const uint16_t* len1;
	const uint16_t* len2;
	long long n_pairs;
	const int* n_dev; // if not null: the number of pairs is read from device memory (<= n_pairs)
	int stride;
	int tile_pairs; // multiple of 8
	int stages;     // <= kMaxStages
	unsigned long long* acc; // [kQcWords]
	int forward_only; // rows of read 2 are not read, nothing is counted as reverse read (single-end input)
	int strict;       // FastqEntry::validate: bases of exactly A,C,G,T,N, qualities of 33..74 only
	int plots;        // also fill the histograms behind the qcML plots (bases by quality, reads by mean quality): one shared-memory atomic per 32 bases
	int* bad_flag;    // if not null: set to 1 when this launch met a character that counts in `errors` (per-chunk report of the FASTQ stream)
	uint8_t bin_of_int[100]; // bin of Histogram(0,60,1) for an integral mean quality k (host-evaluated double expression)
};

// The two per-read bins of the plots from the read's integer quality sum and length (len > 0): index of read_qualities_
// (std::round(sum/len), ties away from zero) and bin of qscore_dist (floor(sum/len / 60 * 60) in double). Both double expressions
// only depend on the exact quotient unless it is an integer (a non-integral quotient with a denominator below 1000 is more than
// 1e-3 away from the next integer, far beyond the rounding of two double operations), so integers and the 100-entry table do.
__device__ __forceinline__ void qc_read_bins(const QcArgs& A, int total, int len, int& rq, int& bin)
{
	const int k = total / len, rem = total - k * len;
	rq = (2 * total + len) / (2 * len);
	bin = rem == 0 ? (int)A.bin_of_int[min(k, 99)] : min(k, 59);
}

constexpr uint32_t kQcBad = 0x80000000u;

__device__ __forceinline__ uint32_t qc_base_field(int c) // one-hot 6-bit field of a base; Pileup::inc (src/cppNGS/Pileup.cpp:17-32)
{
	switch (c)
	{
		case 'A': case 'a': return 1u;
		case 'C': case 'c': return 1u << 6;
		case 'G': case 'g': return 1u << 12;
		case 'T': case 't': return 1u << 18;
		case 'N': case 'n': return 1u << 24;
		case '-': case '~': return 0u; // deletion / ignored: no A,C,G,T,N count
		default: return kQcBad;        // the reference throws "Unknown base"
	}
}
// bits 0..6: q, bit 12: q >= 20, bit 20: q >= 30 (StatisticsReads.cpp:53-60); up to 15 of them can be added without a carry between fields
__device__ __forceinline__ uint32_t qc_qual_field(int byte)
{
	const int q = (int)(signed char)byte - 33;
	if (q < 0 || q >= 100) return kQcBad; // q >= 100 throws in the reference, q < 0 indexes out of bounds there
	return (uint32_t)q | (q >= 20 ? 0x1000u : 0u) | (q >= 30 ? 0x100000u : 0u);
}

// the additional checks of FastqEntry::validate (src/cppNGS/FastqFileStream.cpp:19-45, short reads)
__device__ __forceinline__ bool qc_strict_bad(int strict, int byte, bool base)
{
	if (!strict) return false;
	if (base) return !(byte == 'A' || byte == 'C' || byte == 'G' || byte == 'T' || byte == 'N');
	return byte < 33 || byte > 74;
}

template <int NW, int CW>
__global__ void __launch_bounds__((CW + 1) * 32) qc_kernel(const __grid_constant__ QcArgs A)
{
	constexpr int kThreads = (CW + 1) * 32;
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ __align__(8) uint64_t full_bar[kMaxStages];
	__shared__ __align__(8) uint64_t empty_bar[kMaxStages];
	__shared__ uint32_t lutb[256], lutq[256];
	__shared__ uint32_t s_len[SPG_MAXLEN];
	__shared__ uint32_t s_acc[7][NW * 32]; // [A,C,G,T,N,qsum_f,qsum_r][cycle] of this CTA
	__shared__ uint32_t s_hist[320];       // base_qualities[100] | read_qualities[100] | qscore_dist forward[60] | reverse[60]
	__shared__ unsigned long long s_scalar[8];

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int TP = A.tile_pairs;
	const uint32_t plane_bytes = (uint32_t)TP * (uint32_t)A.stride;
	const uint32_t stage_bytes = 4u * plane_bytes + 4u * (uint32_t)TP;
	const long long n_pairs = A.n_dev ? (long long)*A.n_dev : A.n_pairs;
	const long long n_tiles = (n_pairs + TP - 1) / TP;
	const uint32_t smem_base = smem_u32(smem);

	for (int i = threadIdx.x; i < 256; i += kThreads)
	{
		lutb[i] = qc_base_field(i) | (qc_strict_bad(A.strict, i, true) ? kQcBad : 0u);
		lutq[i] = qc_qual_field(i) | (qc_strict_bad(A.strict, i, false) ? kQcBad : 0u);
	}
	for (int i = threadIdx.x; i < SPG_MAXLEN; i += kThreads) s_len[i] = 0;
	for (int i = threadIdx.x; i < 7 * NW * 32; i += kThreads) (&s_acc[0][0])[i] = 0;
	for (int i = threadIdx.x; i < 320; i += kThreads) s_hist[i] = 0;
	if (threadIdx.x < 8) s_scalar[threadIdx.x] = 0;
	if (threadIdx.x == 0)
	{
		for (int s = 0; s < A.stages; ++s)
		{
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], CW);
		}
		fence_barrier_init();
	}
	__syncthreads();

	if (warp == CW)
	{
		// ===== producer: one lane issues the bulk copies of each tile (same tile layout as trim_kernel) =====
		if (lane == 0)
		{
			int it = 0;
			for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it)
			{
				const int s = it % A.stages;
				const uint32_t round = (uint32_t)(it / A.stages);
				if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1u);
				const long long first = t * TP;
				const int cnt = (int)min((long long)TP, n_pairs - first);
				const uint32_t row_bytes = ((uint32_t)cnt * (uint32_t)A.stride + 15u) & ~15u;
				const uint32_t len_bytes = (uint32_t)((cnt + 7) / 8) * 16u;
				const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
				mbar_arrive_expect_tx(&full_bar[s], 4 * row_bytes + 2 * len_bytes);
				const size_t goff = (size_t)first * A.stride;
				bulk_g2s(st + 0 * plane_bytes, A.b1 + goff, row_bytes, &full_bar[s]);
				bulk_g2s(st + 1 * plane_bytes, A.q1 + goff, row_bytes, &full_bar[s]);
				bulk_g2s(st + 2 * plane_bytes, A.b2 + goff, row_bytes, &full_bar[s]);
				bulk_g2s(st + 3 * plane_bytes, A.q2 + goff, row_bytes, &full_bar[s]);
				bulk_g2s(st + 4 * plane_bytes, A.len1 + first, len_bytes, &full_bar[s]);
				bulk_g2s(st + 4 * plane_bytes + 2u * (uint32_t)TP, A.len2 + first, len_bytes, &full_bar[s]);
			}
		}
	}
	else
	{
		// ===== consumers =====
		uint32_t pk[NW], qs[NW]; // per owned cycle: five 6-bit base counters; quality sums (forward | reverse << 16)
#pragma unroll
		for (int w = 0; w < NW; ++w) pk[w] = qs[w] = 0;
		uint32_t c20 = 0, c30 = 0, bad = 0;
		unsigned long long bases = 0;
		uint32_t reads = 0, rq20 = 0; // warp-uniform, reported by lane 0
		int since_flush = 0;
		// the per-read bins of the plots need two integer divisions: lane k keeps (quality sum, length, direction) of the k-th read
		// since the last round and all lanes divide at once every 32 reads
		int st_total = 0, st_len = 0, st_n = 0;
		auto read_bins = [&]() {
			if (lane < st_n && st_len > 0)
			{
				int rq, bin;
				qc_read_bins(A, st_total & 0x7FFFFFFF, st_len, rq, bin);
				if (rq < 100) atomicAdd(&s_hist[100 + rq], 1u);
				atomicAdd(&s_hist[(st_total < 0 ? 260 : 200) + bin], 1u); // bit 31 of the stashed sum: reverse read
			}
			st_n = 0;
		};

		auto flush = [&]() {
#pragma unroll
			for (int w = 0; w < NW; ++w)
			{
				const int cyc = 32 * w + lane;
#pragma unroll
				for (int k = 0; k < 5; ++k)
				{
					const uint32_t v = (pk[w] >> (6 * k)) & 63u;
					if (v) atomicAdd(&s_acc[k][cyc], v);
				}
				const uint32_t f = qs[w] & 0xFFFFu, r = qs[w] >> 16;
				if (f) atomicAdd(&s_acc[5][cyc], f);
				if (r) atomicAdd(&s_acc[6][cyc], r);
				pk[w] = qs[w] = 0;
			}
			since_flush = 0;
		};

		int it = 0;
		for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it)
		{
			const int s = it % A.stages;
			const uint32_t round = (uint32_t)(it / A.stages);
			mbar_wait(&full_bar[s], round & 1u);
			const long long first = t * TP;
			const int cnt = (int)min((long long)TP, n_pairs - first);
			const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
			const uint32_t lens = st + 4 * plane_bytes;
			for (int pr = warp; pr < cnt; pr += CW)
			{
				const uint32_t roff = (uint32_t)pr * (uint32_t)A.stride;
#pragma unroll
				for (int rd = 0; rd < 2; ++rd)
				{
					if (rd == 1 && A.forward_only) break;
					const uint32_t rb = st + (uint32_t)(2 * rd) * plane_bytes + roff, rq = rb + plane_bytes;
					int len = (int)lds_u16(lens + 2u * (uint32_t)(rd * TP + pr));
					bases += (unsigned long long)len;
					if (lane == 0 && len < SPG_MAXLEN) atomicAdd(&s_len[len], 1u);
					if (len > A.stride || len >= SPG_MAXLEN)
					{
						bad = kQcBad;
						len = min(len, A.stride); // stay inside the row
					}
					uint32_t racc = 0;
#pragma unroll
					for (int w = 0; w < NW; ++w)
					{
						const int pos = 32 * w + lane;
						if (pos < len)
						{
							const uint32_t vb = lutb[lds_u8(rb + pos)], vq = lutq[lds_u8(rq + pos)];
							if (A.plots && !(vq & kQcBad)) atomicAdd(&s_hist[vq & 0x7Fu], 1u); // base_qualities_[q]++
							bad |= vb | vq; // only bit 31 is looked at
							pk[w] += vb;    // a bad base adds bit 31, which no field uses
							racc += vq;
							qs[w] += (vq & 0x7Fu) << (16 * rd);
						}
					}
					c20 += (racc >> 12) & 0xFu;
					c30 += (racc >> 20) & 0xFu;
					const int total = __reduce_add_sync(kFull, (int)(racc & 0xFFFu));
					// mean_qscore = q_sum/cycles >= 20.0 (only if cycles > 0: 0/0 is not a valid float there)
					if (len > 0 && total >= 20 * len) ++rq20;
					if (A.plots)
					{
						if (lane == st_n)
						{
							st_total = total | (rd ? (int)0x80000000 : 0);
							st_len = len;
						}
						if (++st_n == 32) read_bins();
					}
				}
				++reads;
				if (++since_flush == 31) flush(); // 2 reads x 31 pairs = 62 < 64: the 6-bit fields cannot overflow
			}
			__syncwarp();
			if (lane == 0)
			{
				fence_proxy_async();
				mbar_arrive(&empty_bar[s]);
			}
		}
		flush();
		read_bins();
		const unsigned long long t20 = __reduce_add_sync(kFull, c20), t30 = __reduce_add_sync(kFull, c30);
		const bool any_bad = __any_sync(kFull, (bad & kQcBad) != 0);
		if (lane == 0)
		{
			atomicAdd(&s_scalar[kQcReadsF], (unsigned long long)reads);
			if (!A.forward_only) atomicAdd(&s_scalar[kQcReadsR], (unsigned long long)reads);
			atomicAdd(&s_scalar[kQcBases], bases);
			atomicAdd(&s_scalar[kQcReadQ20], (unsigned long long)rq20);
			atomicAdd(&s_scalar[kQcBaseQ20], t20);
			atomicAdd(&s_scalar[kQcBaseQ30], t30);
			if (any_bad) atomicAdd(&s_scalar[kQcErrors], 1ull);
			if (any_bad && A.bad_flag) atomicMax(A.bad_flag, 1);
		}
	}
	__syncthreads();

	// one 64-bit atomic per CTA and non-zero counter
	if (threadIdx.x < 8 && s_scalar[threadIdx.x]) atomicAdd(&A.acc[threadIdx.x], s_scalar[threadIdx.x]);
	for (int i = threadIdx.x; i < SPG_MAXLEN; i += kThreads)
		if (s_len[i]) atomicAdd(&A.acc[kQcLen + i], (unsigned long long)s_len[i]);
	for (int i = threadIdx.x; i < 320; i += kThreads)
		if (s_hist[i]) atomicAdd(&A.acc[kQcBaseQual + i], (unsigned long long)s_hist[i]); // the four histograms are laid out in the same order
	for (int i = threadIdx.x; i < 7 * NW * 32; i += kThreads)
	{
		const int k = i / (NW * 32), cycle = i % (NW * 32);
		const uint32_t v = s_acc[k][cycle];
		if (!v || cycle >= SPG_MAXLEN) continue;
		if (k < 5) atomicAdd(&A.acc[kQcPile + 5 * cycle + k], (unsigned long long)v);
		else atomicAdd(&A.acc[(k == 5 ? kQcQf : kQcQr) + cycle], (unsigned long long)v);
	}
}

// reads longer than the register path (rows > 320 bytes): plain loops with device-wide atomics
__global__ void __launch_bounds__(256) qc_kernel_generic(const __grid_constant__ QcArgs A)
{
	const int lane = threadIdx.x & 31;
	const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
	const long long n_pairs = A.n_dev ? (long long)*A.n_dev : A.n_pairs;
	for (long long r = gwarp; r < n_pairs; r += nwarps)
	{
		for (int rd = 0; rd < (A.forward_only ? 1 : 2); ++rd)
		{
			const uint8_t* brow = (rd ? A.b2 : A.b1) + (size_t)r * A.stride;
			const uint8_t* qrow = (rd ? A.q2 : A.q1) + (size_t)r * A.stride;
			int len = rd ? A.len2[r] : A.len1[r];
			bool bad = len > A.stride || len >= SPG_MAXLEN;
			if (bad) len = min(len, A.stride);
			int rsum = 0, n20 = 0, n30 = 0;
			for (int pos = lane; pos < len && !bad; pos += 32)
			{
				const uint32_t vb = qc_base_field(brow[pos]), vq = qc_qual_field(qrow[pos]);
				if (((vb | vq) & kQcBad) || qc_strict_bad(A.strict, brow[pos], true) || qc_strict_bad(A.strict, qrow[pos], false))
				{
					bad = true;
					break;
				}
				for (int k = 0; k < 5; ++k)
					if ((vb >> (6 * k)) & 1u) atomicAdd(&A.acc[kQcPile + 5 * pos + k], 1ull);
				const int qv = (int)(signed char)qrow[pos] - 33;
				atomicAdd(&A.acc[(rd ? kQcQr : kQcQf) + pos], (unsigned long long)qv);
				rsum += qv;
				if (A.plots) atomicAdd(&A.acc[kQcBaseQual + qv], 1ull);
				n20 += (vq >> 12) & 1u;
				n30 += (vq >> 20) & 1u;
			}
			const int total = __reduce_add_sync(0xffffffffu, rsum);
			const int t20 = __reduce_add_sync(0xffffffffu, n20), t30 = __reduce_add_sync(0xffffffffu, n30);
			const bool any_bad = __any_sync(0xffffffffu, bad);
			if (lane == 0)
			{
				atomicAdd(&A.acc[rd ? kQcReadsR : kQcReadsF], 1ull);
				atomicAdd(&A.acc[kQcBases], (unsigned long long)len);
				if (len < SPG_MAXLEN) atomicAdd(&A.acc[kQcLen + len], 1ull);
				if (len > 0 && total >= 20 * len) atomicAdd(&A.acc[kQcReadQ20], 1ull);
				if (A.plots && len > 0 && !any_bad)
				{
					int rq, bin;
					qc_read_bins(A, total, len, rq, bin);
					if (rq < 100) atomicAdd(&A.acc[kQcReadQual + rq], 1ull);
					atomicAdd(&A.acc[(rd ? kQcDistR : kQcDistF) + bin], 1ull);
				}
				atomicAdd(&A.acc[kQcBaseQ20], (unsigned long long)t20);
				atomicAdd(&A.acc[kQcBaseQ30], (unsigned long long)t30);
				if (any_bad) atomicAdd(&A.acc[kQcErrors], 1ull);
				if (any_bad && A.bad_flag) atomicMax(A.bad_flag, 1);
			}
		}
	}
}

} // namespace spg
