// spg_qc_lanes.cuh -- raw-read statistics (-qc), second layout: one LANE per read pair, like spg_lanes.cuh.
//
// Same accumulators as spg::qc_kernel (StatisticsReads::update, src/cppNGS/StatisticsReads.cpp:26-81). A warp works on 32 pairs at a
// time; in step j every lane holds the four bases and four qualities of cycles 4j .. 4j+3 of ITS read, so the per-cycle sums over
// the 32 reads are warp reductions of packed byte fields (REDUX): per step one PRMT lookup per letter (count bytes 0/1 by the low
// three bits of a base, after an exact check against the canonical letter) and one reduction per letter, two reductions for the
// quality sums (even / odd bytes in 16-bit fields). The reduced words are kept by lane j mod 32 in packed registers and flushed to
// the CTA's per-cycle table every 7 units of a kind; per-read values (quality sum, Q20/Q30 counts, length) never leave the lane.
// About 135 warp instructions per pair instead of 335: one load per four bases, no per-base table lookups.
#pragma once
#include "spg_lanes.cuh"
#include "spg_qc.cuh"

namespace spg
{

// a stage holds ONE plane (the base rows or the quality rows) of ONE read of 32 pairs. A warp reads its stage during its whole pass, so
// the ring has to hold one stage per consumer warp plus the stages that are on their way: with whole reads per stage the ring was as deep
// as there are warps and nothing was ever prefetched (a third of the warps' time went into waiting for the producer).
constexpr int kQcLaneStagesMax = 48;
__host__ __device__ constexpr uint32_t qc_lane_stage_bytes(int stride) { return ((32u * (uint32_t)stride + 16u) + 127u) & ~127u; }

// count bytes (0/1) of one letter for the four bases of a word, by the low three bits of each byte (A 1, C 3, T 4, N 6, G 7; the same
// for lower case): PRMT lookup tables lo (indices 0-3) / hi (4-7)
__host__ __device__ constexpr uint32_t qc_lut_lo(int x) { return x == 0 ? 0x00000100u : x == 1 ? 0x01000000u : 0u; }                        // A C G T N
__host__ __device__ constexpr uint32_t qc_lut_hi(int x) { return x == 2 ? 0x01000000u : x == 3 ? 0x00000001u : x == 4 ? 0x00010000u : 0u; }
constexpr uint32_t kQcCanonLo = 0x43FF41FFu, kQcCanonHi = 0x474EFF54u;            // upper-case letter by index, 0xFF where there is none

// per-byte classification of a word that holds something else than letters (Pileup::inc: '-' and '~' count nothing, anything else is
// an error): count bytes of the five letters
__device__ __forceinline__ void qc_slow_bases(uint32_t w, uint32_t mbytes, uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4, uint32_t& bad)
{
	c0 = c1 = c2 = c3 = c4 = 0;
#pragma unroll 1
	for (int b = 0; b < 4; ++b)
	{
		if (!((mbytes >> (8 * b)) & 0xFFu)) continue;
		const uint32_t f = qc_base_field((int)((w >> (8 * b)) & 0xFFu));
		if (f & kQcBad) bad |= kQcBad;
		const uint32_t one = 1u << (8 * b);
		if (f & 1u) c0 |= one;
		if (f & (1u << 6)) c1 |= one;
		if (f & (1u << 12)) c2 |= one;
		if (f & (1u << 18)) c3 |= one;
		if (f & (1u << 24)) c4 |= one;
	}
}

template <int NW, int CW>
__global__ void __launch_bounds__((CW + 1) * 32, 1) qc_lanes_kernel(const __grid_constant__ QcArgs A)
{
	constexpr int kThreads = (CW + 1) * 32;
	constexpr int NS = (8 * NW + 31) / 32; // accumulator slots per lane: step j is kept by lane j % 32 in slot j / 32
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ __align__(8) uint64_t full_bar[kQcLaneStagesMax];
	__shared__ __align__(8) uint64_t empty_bar[kQcLaneStagesMax];
	__shared__ uint32_t s_len[SPG_MAXLEN];
	__shared__ uint32_t s_acc[7][NW * 32]; // [A,C,G,T,N,qsum_f,qsum_r][cycle] of this CTA
	__shared__ uint32_t s_hist[320];       // base_qualities[100] | read_qualities[100] | qscore_dist forward[60] | reverse[60]
	__shared__ unsigned long long s_scalar[8];
	__shared__ uint32_t next_it;
	__shared__ volatile uint32_t issued;

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int S = A.stages;
	const uint32_t stage_bytes = qc_lane_stage_bytes(A.stride);
	const uint32_t n_pairs = A.n_dev ? (uint32_t)*A.n_dev : (uint32_t)A.n_pairs;
	const uint32_t n_tiles = (n_pairs + 31u) / 32u;
	const uint32_t n_my = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1u) / gridDim.x : 0u;
	const uint32_t smem_base = smem_u32(smem);

	for (int i = threadIdx.x; i < SPG_MAXLEN; i += kThreads) s_len[i] = 0;
	for (int i = threadIdx.x; i < 7 * NW * 32; i += kThreads) (&s_acc[0][0])[i] = 0;
	for (int i = threadIdx.x; i < 320; i += kThreads) s_hist[i] = 0;
	if (threadIdx.x < 8) s_scalar[threadIdx.x] = 0;
	if (threadIdx.x == 0)
	{
		for (int s = 0; s < S; ++s)
		{
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], 1);
		}
		next_it = 0;
		issued = 0;
		fence_barrier_init();
	}
	__syncthreads();

	if (warp == CW)
	{
		// ===== producer: unit u = (tile, read, plane): the base rows or the quality rows of that read of 32 pairs =====
		if (lane == 0)
		{
			const uint32_t per_tile = A.forward_only ? 2u : 4u;
			for (uint32_t u = 0; u < n_my * per_tile; ++u)
			{
				const uint32_t it = u / per_tile, k = u % per_tile;
				const int s = (int)(u % (uint32_t)S);
				const uint32_t round = u / (uint32_t)S;
				if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1u);
				const uint32_t first = (blockIdx.x + it * gridDim.x) * 32u;
				const uint32_t cnt = min(32u, n_pairs - first);
				const uint32_t row_bytes = (cnt * (uint32_t)A.stride + 15u) & ~15u;
				const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
				mbar_arrive_expect_tx(&full_bar[s], row_bytes);
				const uint8_t* src = k == 0 ? A.b1 : k == 1 ? A.q1 : k == 2 ? A.b2 : A.q2;
				bulk_g2s(st, src + (size_t)first * A.stride, row_bytes, &full_bar[s]);
				__threadfence_block();
				issued = u + 1u;
			}
		}
	}
	else
	{
		// ===== consumers: a warp claims one unit (a tile's read 1 or read 2) at a time =====
		uint32_t pile[NS][5];    // [slot][letter]: count of the letter in the four cycles of step 32*slot + lane, one byte each (<= 224); both reads
		uint32_t qsum[2][NS][2]; // [read] quality sums of the even / odd cycles of that step, 16-bit fields (<= 224 * 94)
#pragma unroll
		for (int sl = 0; sl < NS; ++sl)
		{
#pragma unroll
			for (int x = 0; x < 5; ++x) pile[sl][x] = 0;
			qsum[0][sl][0] = qsum[0][sl][1] = qsum[1][sl][0] = qsum[1][sl][1] = 0;
		}
		uint32_t c20 = 0, c30 = 0, bad = 0, rq20 = 0, reads = 0, reads_r = 0;
		unsigned long long bases = 0;
		int since_flush = 0;

		auto flush = [&]() {
#pragma unroll
			for (int sl = 0; sl < NS; ++sl)
			{
				const int cyc0 = 4 * (32 * sl + lane);
				if (cyc0 < NW * 32)
				{
#pragma unroll
					for (int x = 0; x < 5; ++x)
					{
						const uint32_t v = pile[sl][x];
						if (v)
#pragma unroll
							for (int b = 0; b < 4; ++b)
								if ((v >> (8 * b)) & 0xFFu) atomicAdd(&s_acc[x][cyc0 + b], (v >> (8 * b)) & 0xFFu);
					}
#pragma unroll
					for (int rd = 0; rd < 2; ++rd)
					{
						const uint32_t e = qsum[rd][sl][0], o = qsum[rd][sl][1];
						if (e & 0xFFFFu) atomicAdd(&s_acc[5 + rd][cyc0 + 0], e & 0xFFFFu);
						if (o & 0xFFFFu) atomicAdd(&s_acc[5 + rd][cyc0 + 1], o & 0xFFFFu);
						if (e >> 16) atomicAdd(&s_acc[5 + rd][cyc0 + 2], e >> 16);
						if (o >> 16) atomicAdd(&s_acc[5 + rd][cyc0 + 3], o >> 16);
					}
				}
#pragma unroll
				for (int x = 0; x < 5; ++x) pile[sl][x] = 0;
				qsum[0][sl][0] = qsum[0][sl][1] = qsum[1][sl][0] = qsum[1][sl][1] = 0;
			}
			since_flush = 0;
		};

		const uint32_t per_tile = A.forward_only ? 2u : 4u;
		const uint32_t n_units = n_my * per_tile;
		int since_flush_q = 0;
		for (;;)
		{
			uint32_t u = 0;
			if (lane == 0) u = atomicAdd(&next_it, 1u);
			u = __shfl_sync(kFull, u, 0);
			if (u >= n_units) break;
			const uint32_t it = u / per_tile, k = u % per_tile;
			const bool rev = k >= 2u;        // read 2 of the tile
			const bool qual = (k & 1u) != 0; // the quality rows (else the base rows)
			const int s = (int)(u % (uint32_t)S);
			const uint32_t first = (blockIdx.x + it * gridDim.x) * 32u;
			const uint32_t p = first + (uint32_t)lane;
			const bool active = p < n_pairs;
			int len = 0;
			if (active) len = rev ? A.len2[p] : A.len1[p];
			if (active)
			{
				if (!qual) // the per-read counters are kept by the unit of the base rows
				{
					bases += (unsigned long long)len;
					if (len < SPG_MAXLEN) atomicAdd(&s_len[len], 1u);
					if (rev) ++reads_r;
					else ++reads;
				}
				if (len > A.stride || len >= SPG_MAXLEN)
				{
					bad = kQcBad;
					len = min(len, A.stride);
				}
			}
			// the stage of this unit: first wait until its copy has been issued (see trim_lanes_kernel), then for the data
			while (issued <= u) __nanosleep(100);
			mbar_wait(&full_bar[s], (u / (uint32_t)S) & 1u);
			const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
			const int nsteps = (__reduce_max_sync(kFull, len) + 3) >> 2; // warp-uniform
			const uint32_t row = st + (uint32_t)lane * (uint32_t)A.stride;
			const uint32_t ar = row & ~3u;
			const uint32_t sh = (row & 2u) ? 16u : 0u; // rows start on even addresses: a row in the middle of a word is realigned
			uint32_t w_next = lds_u32(ar);
			if (!qual)
			{
				static_for<NS>([&](auto slc) {
					constexpr int sl = decltype(slc)::value;
					const int jend = min(32, nsteps - 32 * sl);
#pragma unroll 2
					for (int jj = 0; jj < jend; ++jj)
					{
						const int j = 32 * sl + jj;
						const uint32_t w0 = w_next;
						w_next = lds_u32(ar + 4u * (uint32_t)(j + 1));
						uint32_t wb = __funnelshift_r(w0, w_next, sh);
						const uint32_t mbytes = low_bits(8 * (len - 4 * j)); // bytes of this word inside the read
						// ---- bases: letters by the low three bits after an exact check against the canonical letter (case ignored)
						wb &= mbytes; // 0x00 beyond the read: index 0, no letter
						uint32_t uu;
						asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(uu) : "r"(wb), "r"(wb >> 4), "r"(0x87878787u));
						const uint32_t sel = prmt(uu, 0u, 0x4420u);
						const uint32_t canon = prmt(kQcCanonLo, kQcCanonHi, sel);
						const uint32_t m01 = mbytes & 0x01010101u;
						uint32_t c0, c1, c2, c3, c4;
						if (((wb ^ canon) & 0xDFDFDFDFu & mbytes) == 0u)
						{
							c0 = prmt(qc_lut_lo(0), qc_lut_hi(0), sel) & m01;
							c1 = prmt(qc_lut_lo(1), qc_lut_hi(1), sel) & m01;
							c2 = prmt(qc_lut_lo(2), qc_lut_hi(2), sel) & m01;
							c3 = prmt(qc_lut_lo(3), qc_lut_hi(3), sel) & m01;
							c4 = prmt(qc_lut_lo(4), qc_lut_hi(4), sel) & m01;
							if (A.strict && ((wb ^ canon) & mbytes)) bad |= kQcBad; // FastqEntry::validate: upper case only
						}
						else
						{
							qc_slow_bases(wb, mbytes, c0, c1, c2, c3, c4, bad);
							if (A.strict) bad |= kQcBad;
						}
						// ---- sums over the 32 reads of the warp; lane jj keeps them
						const uint32_t r0 = __reduce_add_sync(kFull, c0), r1 = __reduce_add_sync(kFull, c1), r2 = __reduce_add_sync(kFull, c2);
						const uint32_t r3 = __reduce_add_sync(kFull, c3), r4 = __reduce_add_sync(kFull, c4);
						if (lane == jj)
						{
							pile[sl][0] += r0;
							pile[sl][1] += r1;
							pile[sl][2] += r2;
							pile[sl][3] += r3;
							pile[sl][4] += r4;
						}
					}
				});
			}
			else
			{
				int tot = 0;
				uint32_t n20 = 0, n30 = 0;
				static_for<NS>([&](auto slc) {
					constexpr int sl = decltype(slc)::value;
					const int jend = min(32, nsteps - 32 * sl);
#pragma unroll 2
					for (int jj = 0; jj < jend; ++jj)
					{
						const int j = 32 * sl + jj;
						const uint32_t w0 = w_next;
						w_next = lds_u32(ar + 4u * (uint32_t)(j + 1));
						uint32_t wq = __funnelshift_r(w0, w_next, sh);
						const uint32_t mbytes = low_bits(8 * (len - 4 * j)); // bytes of this word inside the read
						// ---- qualities: bytes 33 .. 127 are q = 0 .. 94 (anything else is an error: q >= 100 or a negative char)
						wq = (wq & mbytes) | (0x21212121u & ~mbytes); // '!' (q 0) beyond the read
						const uint32_t okq = ((wq & 0x7F7F7F7Fu) + 0x5F5F5F5Fu) & ~wq & 0x80808080u; // bit 7: 33 <= byte < 128
						if (okq != 0x80808080u) bad |= kQcBad;
						if (A.strict && ((wq + 0x35353535u) & 0x80808080u)) bad |= kQcBad; // byte > 74 ('J')
						const uint32_t qv = (wq - 0x21212121u) & 0x7F7F7F7Fu;                // q per byte
						tot = (int)__dp4a(qv, 0x01010101u, (unsigned)tot);
						n20 += __popc((qv + 0x6C6C6C6Cu) & 0x80808080u & mbytes);            // q + 108 >= 128 <=> q >= 20
						n30 += __popc((qv + 0x62626262u) & 0x80808080u & mbytes);            // q + 98 >= 128 <=> q >= 30
						if (A.plots && mbytes)
						{
							for (int b = 0; b < 4; ++b)
								if ((mbytes >> (8 * b)) & 0xFFu)
								{
									const uint32_t q = (qv >> (8 * b)) & 0x7Fu;
									if (q < 100) atomicAdd(&s_hist[q], 1u);
								}
						}
						const uint32_t re = __reduce_add_sync(kFull, qv & 0x00FF00FFu), ro = __reduce_add_sync(kFull, (qv >> 8) & 0x00FF00FFu);
						if (lane == jj)
						{
							qsum[0][sl][0] += rev ? 0u : re;
							qsum[0][sl][1] += rev ? 0u : ro;
							qsum[1][sl][0] += rev ? re : 0u;
							qsum[1][sl][1] += rev ? ro : 0u;
						}
					}
				});
				c20 += n20;
				c30 += n30;
				if (active)
				{
					// mean_qscore = q_sum/cycles >= 20.0 (only if cycles > 0: 0/0 is not a valid float there)
					if (len > 0 && tot >= 20 * len) ++rq20;
					if (A.plots && len > 0)
					{
						int rq, bin;
						qc_read_bins(A, tot, len, rq, bin);
						if (rq < 100) atomicAdd(&s_hist[100 + rq], 1u);
						atomicAdd(&s_hist[(rev ? 260 : 200) + bin], 1u);
					}
				}
			}
			__syncwarp();
			if (lane == 0) // the stage goes back to the producer
			{
				fence_proxy_async();
				mbar_arrive(&empty_bar[s]);
			}
			// 7 units x 32 reads = 224 < 256: the byte fields of the letter counts cannot overflow (nor the 16-bit quality sums: 224 x 94)
			if (qual) ++since_flush_q;
			else ++since_flush;
			if (since_flush == 7 || since_flush_q == 7)
			{
				flush();
				since_flush_q = 0;
			}
		}
		flush();
		const unsigned long long t20 = __reduce_add_sync(kFull, c20), t30 = __reduce_add_sync(kFull, c30);
		const unsigned long long trq = __reduce_add_sync(kFull, rq20), treads = __reduce_add_sync(kFull, reads), treads_r = __reduce_add_sync(kFull, reads_r);
		unsigned long long tb = bases;
#pragma unroll
		for (int d = 16; d > 0; d >>= 1) tb += __shfl_xor_sync(kFull, tb, d);
		const bool any_bad = __any_sync(kFull, (bad & kQcBad) != 0);
		if (lane == 0)
		{
			atomicAdd(&s_scalar[kQcReadsF], treads);
			atomicAdd(&s_scalar[kQcReadsR], treads_r);
			atomicAdd(&s_scalar[kQcBases], tb);
			atomicAdd(&s_scalar[kQcReadQ20], trq);
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
		if (s_hist[i]) atomicAdd(&A.acc[kQcBaseQual + i], (unsigned long long)s_hist[i]);
	for (int i = threadIdx.x; i < 7 * NW * 32; i += kThreads)
	{
		const int k = i / (NW * 32), cycle = i % (NW * 32);
		const uint32_t v = s_acc[k][cycle];
		if (!v || cycle >= SPG_MAXLEN) continue;
		if (k < 5) atomicAdd(&A.acc[kQcPile + 5 * cycle + k], (unsigned long long)v);
		else atomicAdd(&A.acc[(k == 5 ? kQcQf : kQcQr) + cycle], (unsigned long long)v);
	}
}

} // namespace spg
