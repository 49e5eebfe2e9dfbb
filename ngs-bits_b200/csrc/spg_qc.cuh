// spg_qc.cuh -- raw-read statistics (-qc) of a batch: the accumulators of StatisticsReads::update(FastqEntry, direction) of
// imgag/ngs-bits (src/cppNGS/StatisticsReads.cpp:26-81) that the paired-end qcML report needs, as one reduction kernel.
//
// One warp per read pair (grid-stride), lane l owns the cycles l, 32+l, ...: the per-cycle base counts (A,C,G,T,N) and quality
// sums live in that lane's registers for the whole launch, so the inner loop has no atomics: per base one table lookup that
// yields a one-hot 6-bit field (five of them packed in a word, spilled into full counters every 31 pairs) and one for the
// quality (>=20 / >=30 flags). Per read a warp reduction gives the mean quality. Registers are combined per CTA in shared memory
// and added to the device-wide 64-bit accumulators at the end.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/seqpurge_b200.h"

namespace spg
{

// layout of the device-wide accumulator array (unsigned long long)
constexpr int kQcReadsF = 0, kQcReadsR = 1, kQcBases = 2, kQcReadQ20 = 3, kQcBaseQ20 = 4, kQcBaseQ30 = 5, kQcErrors = 6;
constexpr int kQcLen = 8;                          // [SPG_MAXLEN] read length histogram
constexpr int kQcPile = kQcLen + SPG_MAXLEN;       // [SPG_MAXLEN][5]
constexpr int kQcQf = kQcPile + 5 * SPG_MAXLEN;    // [SPG_MAXLEN] quality sums, forward reads
constexpr int kQcQr = kQcQf + SPG_MAXLEN;          // [SPG_MAXLEN] reverse reads
constexpr int kQcWords = kQcQr + SPG_MAXLEN;

struct QcArgs
{
	const uint8_t* b1;
	const uint8_t* q1;
	const uint8_t* b2;
	const uint8_t* q2;
	const uint16_t* len1;
	const uint16_t* len2;
	long long n_pairs;
	int stride;
	unsigned long long* acc; // [kQcWords]
};

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
__device__ __forceinline__ uint32_t qc_qual_field(int byte) // bit 0: q >= 20, bit 16: q >= 30; StatisticsReads.cpp:53-60
{
	const int q = (int)(signed char)byte - 33;
	if (q < 0 || q >= 100) return kQcBad; // q >= 100 throws in the reference, q < 0 indexes out of bounds there
	return (q >= 20 ? 1u : 0u) | (q >= 30 ? 0x10000u : 0u);
}

template <int NW>
__global__ void __launch_bounds__(256) qc_kernel(const __grid_constant__ QcArgs A)
{
	__shared__ uint32_t lutb[256], lutq[256];
	__shared__ uint32_t s_len[SPG_MAXLEN];
	__shared__ uint32_t s_acc[NW * 32 * 7]; // [cycle][A,C,G,T,N,qsum_f,qsum_r] of this CTA
	__shared__ unsigned long long s_scalar[8];

	for (int i = threadIdx.x; i < 256; i += blockDim.x)
	{
		lutb[i] = qc_base_field(i);
		lutq[i] = qc_qual_field(i);
	}
	for (int i = threadIdx.x; i < SPG_MAXLEN; i += blockDim.x) s_len[i] = 0;
	for (int i = threadIdx.x; i < NW * 32 * 7; i += blockDim.x) s_acc[i] = 0;
	if (threadIdx.x < 8) s_scalar[threadIdx.x] = 0;
	__syncthreads();

	const int lane = threadIdx.x & 31;
	const long long gwarp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;

	uint32_t cnt[NW][5], pk[NW], qsum[2][NW];
#pragma unroll
	for (int w = 0; w < NW; ++w)
	{
		pk[w] = 0;
		qsum[0][w] = qsum[1][w] = 0;
#pragma unroll
		for (int k = 0; k < 5; ++k) cnt[w][k] = 0;
	}
	uint32_t c20 = 0, c30 = 0, bad = 0;
	unsigned long long reads[2] = {0, 0}, bases = 0, rq20 = 0; // warp-uniform, kept by every lane, reported by lane 0
	int since_flush = 0;

	for (long long r = gwarp; r < A.n_pairs; r += nwarps)
	{
#pragma unroll
		for (int rd = 0; rd < 2; ++rd)
		{
			const uint8_t* brow = (rd ? A.b2 : A.b1) + (size_t)r * A.stride;
			const uint8_t* qrow = (rd ? A.q2 : A.q1) + (size_t)r * A.stride;
			int len = rd ? A.len2[r] : A.len1[r];
			if (len > A.stride || len >= SPG_MAXLEN)
			{
				bad = kQcBad;
				len = min(len, A.stride); // stay inside the row
			}
			uint32_t pq = 0;
			int rsum = 0;
#pragma unroll
			for (int w = 0; w < NW; ++w)
			{
				const int pos = 32 * w + lane;
				if (pos < len)
				{
					const uint32_t b = brow[pos], q = qrow[pos];
					const uint32_t vb = lutb[b], vq = lutq[q];
					bad |= vb | vq;
					pk[w] += vb & 0x3FFFFFFFu;
					pq += vq & 0x00010001u;
					const int qv = (int)(signed char)q - 33;
					qsum[rd][w] += (uint32_t)qv;
					rsum += qv;
				}
			}
			c20 += pq & 0xFFFFu;
			c30 += pq >> 16;
			const int total = __reduce_add_sync(0xffffffffu, rsum);
			// mean_qscore = q_sum/cycles >= 20.0 (only if cycles > 0: 0/0 is not a valid float there)
			if (len > 0 && total >= 20 * len) ++rq20;
			++reads[rd];
			bases += (unsigned long long)len;
			if (lane == 0 && len < SPG_MAXLEN) atomicAdd(&s_len[len], 1u);
		}
		if (++since_flush == 31) // 2 reads x 31 pairs = 62 < 64: the 6-bit fields cannot overflow
		{
#pragma unroll
			for (int w = 0; w < NW; ++w)
			{
#pragma unroll
				for (int k = 0; k < 5; ++k) cnt[w][k] += (pk[w] >> (6 * k)) & 63u;
				pk[w] = 0;
			}
			since_flush = 0;
		}
	}
#pragma unroll
	for (int w = 0; w < NW; ++w)
#pragma unroll
		for (int k = 0; k < 5; ++k) cnt[w][k] += (pk[w] >> (6 * k)) & 63u;

	// combine the warps of this CTA in shared memory, then one 64-bit atomic per CTA and counter
#pragma unroll
	for (int w = 0; w < NW; ++w)
	{
		uint32_t* row = &s_acc[(32 * w + lane) * 7];
#pragma unroll
		for (int k = 0; k < 5; ++k)
			if (cnt[w][k]) atomicAdd(&row[k], cnt[w][k]);
		if (qsum[0][w]) atomicAdd(&row[5], qsum[0][w]);
		if (qsum[1][w]) atomicAdd(&row[6], qsum[1][w]);
	}
	const unsigned long long t20 = __reduce_add_sync(0xffffffffu, c20), t30 = __reduce_add_sync(0xffffffffu, c30);
	const bool any_bad = __any_sync(0xffffffffu, (bad & kQcBad) != 0);
	if (lane == 0)
	{
		atomicAdd(&s_scalar[kQcReadsF], reads[0]);
		atomicAdd(&s_scalar[kQcReadsR], reads[1]);
		atomicAdd(&s_scalar[kQcBases], bases);
		atomicAdd(&s_scalar[kQcReadQ20], rq20);
		atomicAdd(&s_scalar[kQcBaseQ20], t20);
		atomicAdd(&s_scalar[kQcBaseQ30], t30);
		if (any_bad) atomicAdd(&s_scalar[kQcErrors], 1ull);
	}
	__syncthreads();
	if (threadIdx.x < 8 && s_scalar[threadIdx.x]) atomicAdd(&A.acc[threadIdx.x], s_scalar[threadIdx.x]);
	for (int i = threadIdx.x; i < SPG_MAXLEN; i += blockDim.x)
		if (s_len[i]) atomicAdd(&A.acc[kQcLen + i], (unsigned long long)s_len[i]);
	for (int i = threadIdx.x; i < NW * 32 * 7; i += blockDim.x)
	{
		const uint32_t v = s_acc[i];
		if (!v) continue;
		const int cycle = i / 7, k = i % 7;
		if (cycle >= SPG_MAXLEN) continue;
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
	for (long long r = gwarp; r < A.n_pairs; r += nwarps)
	{
		for (int rd = 0; rd < 2; ++rd)
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
				if ((vb | vq) & kQcBad)
				{
					bad = true;
					break;
				}
				for (int k = 0; k < 5; ++k)
					if ((vb >> (6 * k)) & 1u) atomicAdd(&A.acc[kQcPile + 5 * pos + k], 1ull);
				const int qv = (int)(signed char)qrow[pos] - 33;
				atomicAdd(&A.acc[(rd ? kQcQr : kQcQf) + pos], (unsigned long long)qv);
				rsum += qv;
				n20 += vq & 1u;
				n30 += (vq >> 16) & 1u;
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
				atomicAdd(&A.acc[kQcBaseQ20], (unsigned long long)t20);
				atomicAdd(&A.acc[kQcBaseQ30], (unsigned long long)t30);
				if (any_bad) atomicAdd(&A.acc[kQcErrors], 1ull);
			}
		}
	}
}

} // namespace spg
