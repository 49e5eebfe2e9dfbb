// spg_lanes.cuh -- sm_100a trimming kernel, second layout: one LANE per read pair (a warp works on 32 pairs at a time).
//
// Same computation as spg::trim_kernel (the per-pair body of AnalysisWorker::run, src/SeqPurge/AnalysisWorker.cpp:122-441), for the
// variants compiled for one read length FULL. The warp-per-pair kernel spends more than half of its instructions on moving data
// between lanes (ballots to pack, shuffles and votes to decide, one record per warp); with one pair per lane
//   * a read is packed from its own bytes with word loads and multiplications (4 bases per LDS.32: mask, multiply-gather the four
//     bits of a plane into a nibble, funnel-shift the nibble into the plane word); the check "only A/C/G/T" is a PRMT lookup of the
//     canonical byte by the low three bits of each byte, compared with the byte itself (exact for all 256 values),
//   * the insert sweep and the adapter scans are loops over the bit shift r = offset mod 32 with the word index as a compile-time
//     constant; no votes, no shuffles: a lane notes its own survivors / hits in divergent code that is rarely taken,
//   * quality trimming reads the last bytes of the quality rows straight from global memory (the quality rows are never staged:
//     only the sectors that hold a trimming point are fetched),
//   * every lane writes its own 8-byte record (coalesced).
// Only the base rows go through the TMA ring (2 bulk copies per 32 pairs). Pairs that are not "two reads of FULL bases made of
// A/C/G/T only" (N, other bytes, ragged lengths) and pairs with more pre-filter survivors than the per-lane queue holds are handed
// to the warp-cooperative general path of spg_kernel.cuh (process_pair on a copy of the four rows), so results are identical by
// construction and are cross-checked in the tests against the oracle, the byte-wise path and the warp-per-pair kernel.
#pragma once
#include "spg_kernel.cuh"

namespace spg
{

constexpr int kLaneStagesMax = 8;
constexpr int kLaneQCap = 6; // pre-filter survivors a lane can queue per pair; more: general path

template <int OFF>
__device__ __forceinline__ uint32_t lds_u32_at(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
	return v;
}
// a load from a table that is written once at kernel start (behind a __syncthreads): not volatile, so that the compiler may move it
// ahead of the arithmetic that precedes its use
template <int OFF>
__device__ __forceinline__ int lds_s16_const_at(uint32_t a)
{
	int v;
	asm("ld.shared.s16 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
	return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t a)
{
	uint2 v;
	asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
	return v;
}
__device__ __forceinline__ void sts_v2(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
__host__ __device__ constexpr uint32_t low_mask_const(int n) { return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u)); }
__host__ __device__ constexpr uint64_t low_mask64_const(int n) { return n >= 64 ? ~0ull : (n <= 0 ? 0ull : ((1ull << n) - 1ull)); }
// Words of an overlap of at most T bases that the pre-filter of the insert sweep looks at: the lo-plane differences inside the
// first c words are a lower bound of the mismatch count just like those of the whole overlap, and for unrelated reads (half of
// the lo bits differ) 32c positions exceed a limit of T/5 mismatches by 3.7 standard deviations when 16c - 3.7 sqrt(8c) >= T/5.
// A weaker filter only means more offsets for the exact count (with a much laxer -match_perc), never a different result.
__host__ __device__ constexpr int sweep_words(int T) { return T <= 27 ? 1 : T <= 86 ? 2 : T <= 149 ? 3 : T <= 215 ? 4 : T <= 283 ? 5 : 6; }

// ---- per-warp shared memory of the lane path -----------------------------------------------------------------------------------------
// plane copies [4][NW+1][32 lanes] (read 1 forward hi/lo, read 2 reversed hi/lo; word NW is a zero pad), the quality window
// [8][32 lanes], the survivor queue [kLaneQCap][32 lanes] u16 and the list of pairs that wait for the general path. The copy area
// doubles as the store of the packing loop's accumulators and as the row buffer of the general path.
template <int NW>
struct LaneSmem
{
	static constexpr uint32_t kCopyBytes = 4u * (NW + 1) * 128u;
	static constexpr uint32_t kQualBytes = 0u; // the general quality search runs when the copy area is free: its window lives there (7 words per lane)
	static constexpr uint32_t kQueueBytes = kLaneQCap * 64u;
	static constexpr uint32_t kRareBytes = 32u * 8u; // pairs waiting for the general path: pair index, len1 | len2 << 16
	static constexpr uint32_t kWarpBytes = kCopyBytes + kQualBytes + kQueueBytes + kRareBytes;
};
__host__ __device__ constexpr uint32_t lane_stage_bytes(int stride) { return ((2u * 32u * (uint32_t)stride + 16u) + 127u) & ~127u; }

// ---- packing: one read per lane ------------------------------------------------------------------------------------------------------
// Multipliers that gather bit 1 (lo plane) / bit 2 (hi plane) of the four bytes of a word into the top nibble, first base at bit 31:
// byte i's bit sits at 8i+1 (8i+2) and goes to 31-i; all other partial products land below bit 28 on distinct positions (no carries).
constexpr uint32_t kMulLo = (1u << 30) | (1u << 21) | (1u << 12) | (1u << 3);
constexpr uint32_t kMulHi = (1u << 29) | (1u << 20) | (1u << 11) | (1u << 2);
// canonical byte by the low three bits of a base: 1 'A', 3 'C', 7 'G', 4 'T'; every other index gives 0xFF, whose own index is 7
// ('G'), so a byte equals its lookup exactly if it is one of A/C/G/T. Bits 7 / 3 of a byte leak into bit 3 of its selector nibble
// (PRMT's sign-replicate mode, result 0x00 or 0xFF): such a byte is never equal to its lookup either.
constexpr uint32_t kLutLo = 0x43FF41FFu, kLutHi = 0x47FFFF54u;

// One word of a row: four bases into the top nibbles of the two accumulators, and the difference to the canonical bytes.
__device__ __forceinline__ uint32_t lane_pack_word(uint32_t w, uint32_t& acch, uint32_t& accl)
{
	accl = __funnelshift_l((w & 0x02020202u) * kMulLo, accl, 4);
	acch = __funnelshift_l((w & 0x04040404u) * kMulHi, acch, 4);
	uint32_t u; // byte k: index of byte k | index of byte k+1 << 4 = bit select (w & c) | ((w >> 4) & ~c), one LOP3
	asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(u) : "r"(w), "r"(w >> 4), "r"(0x07070707u));
	return prmt(kLutLo, kLutHi, prmt(u, 0u, 0x4420u)) ^ w;          // canonical bytes ^ bytes
}

// Packs the two rows of a pair (shared addresses, 2-byte aligned, FULL bases each). The accumulator words go to the lane's plane
// copy area: word (2*read + plane) * NW + g at acc + 128 * that index, plane 0 = hi, 1 = lo. acc[g] = brev(natural word g), natural
// bit b of word g = base 32g + b - odd, odd = row & 2 (a row that starts in the middle of a word drags two foreign bytes in front).
// One loop body serves all full groups of 8 words of both reads (the code stays in the instruction cache); only the last group,
// where the read ends, is separate. Returns != 0 if a byte of one of the reads is not one of A/C/G/T.
template <int NW, int FULL>
__device__ __forceinline__ uint32_t lane_pack_rows(uint32_t row1, uint32_t row2, uint32_t acc)
{
	constexpr int NWRD = (FULL + 2 + 3) / 4, NG = (NWRD + 7) / 8, NLAST = NWRD - 8 * (NG - 1);
	static_assert(4 * NWRD <= 32 * NW, "the words of a row must fit the planes");
	static_assert(NG >= 2 && 4 * (8 * (NG - 1)) + 3 < FULL, "words that reach beyond the read must lie in the last group");
	uint32_t bad = 0;
#pragma unroll 1
	for (int rd = 0; rd < 2; ++rd)
	{
		const uint32_t row = rd ? row2 : row1;
		const uint32_t odd = row & 2u;
		uint32_t a = row & ~3u;
		uint32_t dst = acc + (uint32_t)rd * (2u * NW * 128u);
		uint32_t first_mask = odd ? 0xFFFF0000u : 0xFFFFFFFFu; // the two bytes in front of the read
#pragma unroll 1
		for (int g = 0; g < NG - 1; ++g)
		{
			uint32_t ah = 0, al = 0;
			static_for<8>([&](auto kc) {
				constexpr int k = decltype(kc)::value;
				uint32_t d = lane_pack_word(lds_u32_at<4 * k>(a), ah, al);
				if constexpr (k == 0) d &= first_mask;
				bad |= d;
			});
			first_mask = 0xFFFFFFFFu;
			sts_u32(dst, ah);
			sts_u32(dst + NW * 128u, al);
			dst += 128u;
			a += 32u;
		}
		{
			uint32_t ah = 0, al = 0;
			static_for<NLAST>([&](auto kc) {
				constexpr int k = decltype(kc)::value;
				constexpr int j = 8 * (NG - 1) + k;
				uint32_t d = lane_pack_word(lds_u32_at<4 * k>(a), ah, al);
				if constexpr (4 * j + 3 >= FULL) // bytes behind the read
				{
					const int nvalid = FULL + (int)odd - 4 * j;
					d &= nvalid >= 4 ? 0xFFFFFFFFu : (nvalid <= 0 ? 0u : ((1u << (8 * nvalid)) - 1u));
				}
				bad |= d;
			});
			if constexpr (NLAST < 8)
			{
				ah <<= 4 * (8 - NLAST);
				al <<= 4 * (8 - NLAST);
			}
			sts_u32(dst, ah);
			sts_u32(dst + NW * 128u, al);
#pragma unroll
			for (int g = NG; g < NW; ++g) // plane words beyond the row
			{
				sts_u32(dst + 128u * (uint32_t)(g - NG + 1), 0u);
				sts_u32(dst + NW * 128u + 128u * (uint32_t)(g - NG + 1), 0u);
			}
		}
	}
	return bad;
}
template <int NW>
__device__ __forceinline__ void lane_load_acc(uint32_t acc, int rd, int plane, uint32_t (&v)[NW])
{
#pragma unroll
	for (int g = 0; g < NW; ++g) v[g] = lds_u32(acc + (uint32_t)((2 * rd + plane) * NW + g) * 128u);
}

// natural, left-aligned plane words (bit b of word w = base 32w+b) from the reversed accumulators
template <int NW, int FULL>
__device__ __forceinline__ void lane_forward(const uint32_t (&acc)[NW], uint32_t odd, uint32_t (&f)[NW])
{
	uint32_t n[NW + 1];
#pragma unroll
	for (int w = 0; w < NW; ++w) n[w] = __brev(acc[w]);
	n[NW] = 0;
#pragma unroll
	for (int w = 0; w < NW; ++w) f[w] = __funnelshift_r(n[w], n[w + 1], odd) & low_mask_const(FULL - 32 * w);
}
// reversed read, left aligned: bit j = base FULL-1-j (the lo plane of revcomp(read); its hi plane is the complement of this one's)
template <int NW, int FULL>
__device__ __forceinline__ void lane_reversed(const uint32_t (&acc)[NW], uint32_t odd, uint32_t (&r)[NW])
{
	constexpr int D0 = 32 * NW - FULL, WO = D0 >> 5, BS = D0 & 31;
	static_assert(BS >= 2, "the alignment shift must stay inside a word");
	const uint32_t bs = (uint32_t)BS - odd;
#pragma unroll
	for (int w = 0; w < NW; ++w)
	{
		const int i0 = w + WO, i1 = w + WO + 1; // words of the reversed string: rev[i] = acc[NW-1-i]
		const uint32_t lo = i0 < NW ? acc[NW - 1 - (i0 < NW ? i0 : 0)] : 0u;
		const uint32_t hi = i1 < NW ? acc[NW - 1 - (i1 < NW ? i1 : 0)] : 0u;
		r[w] = __funnelshift_r(lo, hi, bs) & low_mask_const(FULL - 32 * w);
	}
}

// the FULL-bit string of a plane the other way round (forward <-> reversed read), both left aligned
template <int NW, int FULL>
__device__ __forceinline__ void lane_flip(const uint32_t (&in)[NW], uint32_t (&out)[NW])
{
	uint32_t acc[NW];
#pragma unroll
	for (int w = 0; w < NW; ++w) acc[w] = __brev(in[w]);
	lane_reversed<NW, FULL>(acc, 0u, out);
}

// 32 bits of a plane copy starting at bit `pos` (0 <= pos < 32*NW); plane: shared address of word 0 of this lane
__device__ __forceinline__ uint32_t lane_extract(uint32_t plane, int pos)
{
	const uint32_t a = plane + 128u * (uint32_t)(pos >> 5);
	return __funnelshift_r(lds_u32(a), lds_u32(a + 128u), pos);
}

// ---- survivors of the pre-filter, per lane ----------------------------------------------------------------------------------------------
// exact mismatch count of insert offset o (both planes; AnalysisWorker.cpp:151-168 for reads without N)
template <int NW, int FULL>
__device__ __forceinline__ int lane_exact_mm(uint32_t copy, int o, const uint32_t (&f1h)[NW], const uint32_t (&f1l)[NW])
{
	const uint32_t r2h = copy + 2u * (NW + 1) * 128u, r2l = copy + 3u * (NW + 1) * 128u;
	int mm = 0;
#pragma unroll
	for (int k = 0; k < NW; ++k)
	{
		const int n = FULL - o - 32 * k; // compared positions in word k of read 1
		if (n > 0)
		{
			const uint32_t sh = lane_extract(r2h, o + 32 * k), sl = lane_extract(r2l, o + 32 * k);
			mm += __popc((~(sh ^ f1h[k]) | (sl ^ f1l[k])) & low_bits(n)); // hi plane of revcomp = complement of the reversed hi plane
		}
	}
	return mm;
}

// probability rank + adapter-presence check of a candidate (AnalysisWorker.cpp:178-259) for a pair of two FULL-length reads without N;
// (rank << 16) | offset, or kNoKey. Same decisions as candidate_key_warp, evaluated by one lane on the plane copies.
template <int NW, int FULL>
__device__ __forceinline__ uint32_t lane_candidate_key(const KArgs& A, uint32_t copy, int o, int m, int mm)
{
	int n = m, mis = mm, cnt = m + mm;
	while (cnt >= kRankDim)
	{
		n >>= 1;
		mis >>= 1;
		cnt = n + mis;
	}
	const uint32_t rank = __ldg(&A.ranktab[cnt * kRankDim + n]);
	if (rank == 0xFFFFu) return kNoKey;
	const int alen = min(A.ao, o); // both fragments: seq1.mid(len2-o, ao) and R2[len2-o ..) hold min(ao, o) bases when len1 == len2
	const int pos = FULL - o;
	const uint32_t v1h = lane_extract(copy, pos), v1l = lane_extract(copy + (NW + 1) * 128u, pos);
	// R2[pos + i] = reversed[o - 1 - i]: take reversed bits [o-alen, o) and turn them around
	const uint32_t e2h = lane_extract(copy + 2u * (NW + 1) * 128u, o - alen), e2l = lane_extract(copy + 3u * (NW + 1) * 128u, o - alen);
	const uint32_t v2h = __brev(e2h) >> (32 - alen), v2l = __brev(e2l) >> (32 - alen);
	const uint32_t val1 = low_bits(alen) & ~A.a1n, val2 = low_bits(alen) & ~A.a2n;
	const int mm1 = __popc(((v1h ^ A.a1h) | (v1l ^ A.a1l)) & val1), m1 = __popc(val1) - mm1;
	const int mm2 = __popc(((v2h ^ A.a2h) | (v2l ^ A.a2l)) & val2), m2 = __popc(val2) - mm2;
	if (o < 10)
	{
		const int max_mm = o < 3 ? 0 : (o < 6 ? 1 : 2);
		if (!(mm1 <= max_mm || mm2 <= max_mm)) return kNoKey;
	}
	else
	{
		const double p1 = __ldg(&A.psmall[(m1 + mm1) * (A.ao + 1) + m1]);
		const double p2 = __ldg(&A.psmall[(m2 + mm2) * (A.ao + 1) + m2]);
		if (__dmul_rn(p1, p2) > A.mep) return kNoKey;
	}
	return (rank << 16) | (uint32_t)o;
}

// ---- adapter-only scan of one read per lane (AnalysisWorker.cpp:307-353 / :355-407): first passing offset or -1 ---------------------------
// fh/fl: forward planes of the read; ah/al: planes of the first a_size adapter bases. Full windows are isolated by a multiplication
// (a_size bits to the top) and pass with at most maxmm mismatches; windows cut by the read end take multiplier and limit from
// FullTab::r1tail (indexed by round and bit shift).
template <int NW, int FULL>
__device__ __forceinline__ int lane_adapter_scan(const KArgs& A, const FullTab<NW, FULL>& F, const uint32_t (&fh)[NW], const uint32_t (&fl)[NW], uint32_t ah, uint32_t al,
                                                 int maxmm)
{
	constexpr int QF = FullTab<NW, FULL>::QF;
	const uint32_t amul = 1u << (32 - A.a_size);
	uint32_t best = 0xFFFFFFFFu;
#pragma unroll 1
	for (int r = 0; r < 32; ++r)
	{
		int mmq[NW];
		bool hit = false;
		const uint32_t tail_addr = smem_u32(F.r1tail) + 8u * (uint32_t)r;
		static_for<NW>([&](auto qc) {
			constexpr int q = decltype(qc)::value;
			const uint32_t sh = __funnelshift_r(fh[q], (q + 1 < NW) ? fh[q + 1 < NW ? q + 1 : 0] : 0u, r);
			const uint32_t sl = __funnelshift_r(fl[q], (q + 1 < NW) ? fl[q + 1 < NW ? q + 1 : 0] : 0u, r);
			const uint32_t x = (sh ^ ah) | (sl ^ al);
			if constexpr (q < QF)
			{
				mmq[q] = __popc(x * amul);
				hit |= mmq[q] <= maxmm;
			}
			else
			{
				const uint2 t = lds_v2_at<256 * (q - QF)>(tail_addr); // F.r1tail[q - QF][r]
				mmq[q] = __popc(x * t.x);
				hit |= mmq[q] <= (int)t.y;
			}
		});
		if (hit) // rare
		{
#pragma unroll
			for (int q = NW - 1; q >= 0; --q)
			{
				const int lim = q < QF ? maxmm : F.r1tail[q < QF ? 0 : q - QF][r].y;
				if (mmq[q] <= lim) best = min(best, (uint32_t)(32 * q + r));
			}
		}
	}
	return (int)best; // 0xFFFFFFFF -> -1
}

// The same scan with a bit-parallel filter in front (used when KArgs::seed_ok): a window that passes has fewer mismatches than it
// holds complete 4-base blocks of the adapter (checked on the host for every window length), so one of those blocks matches the
// read exactly (pigeonhole). "Block j of the adapter occurs at offset o" is evaluated for 32 offsets per instruction: the read's four
// base indicators (bit p = base p is A / C / G / T) are kept in shared memory (`ind`, [4][NW+1] words per lane, which of them an
// adapter position needs comes from KArgs::a?off), shifted by the position inside the adapter and ANDed. Only the offsets that
// survive (about 2 % of them) get the exact count of lane_adapter_scan, with the same multipliers and limits.
// HASN: the read may hold N (plane fn; N is packed like G): N positions match no adapter base, and every offset whose window holds
// an N is a candidate as well (the pigeonhole argument only covers windows of A/C/G/T); candidates are then decided with the
// general rule (pass table by number of compared bases), which equals the limit rule for windows without N.
template <int NW, int FULL, bool HASN = false>
__device__ __forceinline__ int lane_adapter_scan_seeds(const KArgs& A, const SmemTables& T, const FullTab<NW, FULL>& F, const uint32_t (&fh)[NW], const uint32_t (&fl)[NW],
                                                       const uint32_t (&fn)[NW], uint32_t ind, const uint16_t* aoff, uint32_t ah, uint32_t al, uint32_t an, int maxmm)
{
	constexpr int QF = FullTab<NW, FULL>::QF;
#pragma unroll
	for (int w = 0; w < NW; ++w)
	{
		sts_u32(ind + (0u * (NW + 1) + w) * 128u, ~(fh[w] | fl[w]) & low_mask_const(FULL - 32 * w)); // A = 00 (also what lies behind the read)
		sts_u32(ind + (1u * (NW + 1) + w) * 128u, ~fh[w] & fl[w]);                                     // C = 01
		sts_u32(ind + (2u * (NW + 1) + w) * 128u, fh[w] & fl[w] & (HASN ? ~fn[w] : kFull));            // G = 11
		sts_u32(ind + (3u * (NW + 1) + w) * 128u, fh[w] & ~fl[w]);                                     // T = 10
	}
#pragma unroll
	for (int k = 0; k < 4; ++k) sts_u32(ind + ((uint32_t)k * (NW + 1) + NW) * 128u, 0u);
	uint32_t cand[NW];
#pragma unroll
	for (int w = 0; w < NW; ++w) cand[w] = 0;
	if (HASN)
	{
		// offsets o with an N in [o, o+32): the N plane smeared towards lower positions
		int n_count = 0;
#pragma unroll
		for (int w = 0; w < NW; ++w)
		{
			cand[w] = fn[w];
			n_count += __popc(fn[w]);
		}
#pragma unroll
		for (int d = 1; d < 32; d <<= 1)
		{
			uint32_t sh[NW];
#pragma unroll
			for (int w = 0; w < NW; ++w) sh[w] = __funnelshift_r(cand[w], (w + 1 < NW) ? cand[w + 1 < NW ? w + 1 : 0] : 0u, d);
#pragma unroll
			for (int w = 0; w < NW; ++w) cand[w] |= sh[w];
		}
		// A full window with ONE N still has a clean exact block when the parameters say so (KArgs::seed_n1_ok: fewer mismatches
		// allowed on a_size-1 bases than blocks that the N leaves intact): then only the windows cut by the read's end need the
		// explicit check. Reads with several N keep every window that holds one.
		if (n_count == 1 && A.seed_n1_ok)
		{
			const int first_cut = FULL - A.a_size + 1; // first offset whose window is cut
#pragma unroll
			for (int w = 0; w < NW; ++w) cand[w] &= ~low_bits(first_cut - 32 * w);
		}
	}
	const int nblk = A.a_size >> 2;
#pragma unroll 1
	for (int j = 0; j < nblk; ++j)
	{
		uint32_t acc[NW];
#pragma unroll
		for (int w = 0; w < NW; ++w) acc[w] = kFull;
#pragma unroll
		for (int t = 0; t < 4; ++t)
		{
			const int i = 4 * j + t;
			const uint32_t base = ind + (uint32_t)aoff[i];
			uint32_t iw[NW + 1];
#pragma unroll
			for (int w = 0; w <= NW; ++w) iw[w] = lds_u32(base + 128u * (uint32_t)w);
#pragma unroll
			for (int w = 0; w < NW; ++w) acc[w] &= __funnelshift_r(iw[w], iw[w + 1], i); // bit o: read base o+i equals adapter base i
		}
#pragma unroll
		for (int w = 0; w < NW; ++w) cand[w] |= acc[w];
	}
	const uint32_t amul = 1u << (32 - A.a_size);
	uint32_t best = 0xFFFFFFFFu;
	static_for<NW>([&](auto wc) {
		constexpr int w = decltype(wc)::value;
		uint32_t c = cand[w];
		while (c != 0 && best == 0xFFFFFFFFu)
		{
			const int b = __ffs((int)c) - 1;
			c &= c - 1;
			const uint32_t sh = __funnelshift_r(fh[w], (w + 1 < NW) ? fh[w + 1 < NW ? w + 1 : 0] : 0u, b);
			const uint32_t sl = __funnelshift_r(fl[w], (w + 1 < NW) ? fl[w + 1 < NW ? w + 1 : 0] : 0u, b);
			const uint32_t x = (sh ^ ah) | (sl ^ al);
			if constexpr (HASN)
			{
				const int cnt = min(A.a_size, FULL - 32 * w - b);
				const uint32_t sn = __funnelshift_r(fn[w], (w + 1 < NW) ? fn[w + 1 < NW ? w + 1 : 0] : 0u, b);
				const uint32_t valid = low_bits(cnt) & ~an & ~sn;
				if (cnt > 0 && ((T.passM[__popc(valid)] >> __popc(x & valid)) & 1u)) best = (uint32_t)(32 * w + b);
			}
			else
			{
				uint32_t mul = amul;
				int lim = maxmm;
				if constexpr (w >= QF)
				{
					const int2 t = F.r1tail[w - QF][b];
					mul = (uint32_t)t.x;
					lim = t.y;
				}
				if (__popc(x * mul) <= lim) best = (uint32_t)(32 * w + b);
			}
		}
	});
	return (int)best;
}

// ---- FastqEntry::trimQuality (src/cppNGS/FastqFileStream.cpp:52-87), one read per lane ----------------------------------------------------
// Window 5 (the default) in registers, on the raw bytes (thresholds shifted by the quality offset), which is the reference's
// signed-char arithmetic as long as all bytes are below 0x80 (anything else: the general search).
// (1) lane_last_good: the highest position below n whose quality reaches the cutoff, 16 aligned bytes per step from the 3' end. A
//     window that starts above it holds only bases below the cutoff and cannot pass, so the search proper starts there -- long
//     low-quality tails cost one load and a few SWAR instructions per 16 bases instead of one window sum per base.
// (2) lane_quality5_block: 16 qualities [lo, lo+16), the window starts lo .. e-5 among them from the top: twelve sliding sums (sign
//     bits of sum - threshold collected by funnel shifts), "quality >= cutoff" of the 16 bases by one SWAR addition per word.
// Only the sectors that hold the end of the read are ever fetched (the quality planes may even live in pinned host memory).

// bit j of the result: byte j of the 16 (v[j/4], byte j%4) is >= cutb (1..127); all bytes must be below 0x80
__device__ __forceinline__ uint32_t lane_ge16(const uint32_t (&v)[4], int cutb)
{
	// byte + 128 - cutb carries into bit 7; the four bits 7 of a word are gathered into the top nibble by a multiplication
	uint32_t ge = 0;
#pragma unroll
	for (int k = 3; k >= 0; --k) ge = __funnelshift_l(((v[k] + (uint32_t)(0x80 - cutb) * 0x01010101u) & 0x80808080u) * 0x00204081u, ge, 4);
	return ge;
}

// highest position p < n with quality >= cutoff; -1: none; -2: a byte >= 0x80 was met
__device__ __forceinline__ int lane_last_good(const uint8_t* qrow, int n, int cutb)
{
	int end = n;
	while (end > 0)
	{
		const uintptr_t a = ((uintptr_t)qrow + (uintptr_t)(end - 1)) & ~(uintptr_t)15; // aligned chunk that holds position end-1
		const int p0 = (int)((intptr_t)a - (intptr_t)(uintptr_t)qrow);                    // position of its first byte (may lie in front of the row)
		const uint4 c = __ldg(reinterpret_cast<const uint4*>(a));
		const uint32_t v[4] = {c.x, c.y, c.z, c.w};
		const uint32_t in = low_bits(end - p0) & ~low_bits(-p0); // bytes of the chunk that are positions [max(p0,0), end)
		// a byte >= 0x80 inside the read: the general search decides (bit 7 of byte j -> bit j by the same multiplication)
		uint32_t hb = 0;
#pragma unroll
		for (int k = 3; k >= 0; --k) hb = __funnelshift_l((v[k] & 0x80808080u) * 0x00204081u, hb, 4);
		if (hb & in) return -2;
		uint32_t w[4];
#pragma unroll
		for (int k = 0; k < 4; ++k) w[k] = v[k] & 0x7F7F7F7Fu; // bytes outside the read may be anything
		const uint32_t ge = lane_ge16(w, cutb) & in;
		if (ge) return p0 + 31 - __clz(ge);
		end = p0;
	}
	return -1;
}

// windows of 5 starting at lo .. e-5 (lo <= e-5 is not required), bytes [lo, lo+16) of the row. Returns the new length if one of them
// passes (highest start first, then the trailing bases below the cutoff are dropped), -1 if none does, -2 for a byte >= 0x80.
__device__ __forceinline__ int lane_quality5_eval(const uint32_t (&v)[4], int lo, int e, int cutb, int thrb); // v: the 16 bytes [lo, lo+16)
__device__ __forceinline__ int lane_quality5_block(const uint8_t* qrow, int lo, int e, int cutb, int thrb)
{
	// two aligned 16-byte loads cover the 16 bytes wherever they start; the second one is not needed for an aligned start
	const uintptr_t a = (uintptr_t)qrow + (uintptr_t)lo;
	const uint32_t off = (uint32_t)(a & 15u);
	const uint4* wp = reinterpret_cast<const uint4*>(a - off);
	const uint4 c0 = __ldg(wp);
	uint4 c1 = make_uint4(0u, 0u, 0u, 0u);
	if (off) c1 = __ldg(wp + 1);
	// words off/4 .. off/4+4 of the eight, then the byte shift inside a word
	const bool s2 = off & 8u, s1 = off & 4u;
	const uint32_t x0 = s2 ? c0.z : c0.x, x1 = s2 ? c0.w : c0.y, x2 = s2 ? c1.x : c0.z, x3 = s2 ? c1.y : c0.w, x4 = s2 ? c1.z : c1.x, x5 = s2 ? c1.w : c1.y;
	uint32_t w[5];
	w[0] = s1 ? x1 : x0;
	w[1] = s1 ? x2 : x1;
	w[2] = s1 ? x3 : x2;
	w[3] = s1 ? x4 : x3;
	w[4] = s1 ? x5 : x4;
	const uint32_t boff = off & 3u;
	uint32_t v[4];
#pragma unroll
	for (int k = 0; k < 4; ++k) v[k] = __funnelshift_r(w[k], w[k + 1], 8u * boff); // bytes of positions lo+4k ..
	return lane_quality5_eval(v, lo, e, cutb, thrb);
}
__device__ __forceinline__ int lane_quality5_eval(const uint32_t (&v)[4], int lo, int e, int cutb, int thrb)
{
	if ((v[0] | v[1] | v[2] | v[3]) & 0x80808080u) return -2;
	const uint32_t ge = lane_ge16(v, cutb); // bit j: quality of position lo+j reaches the cutoff
	auto byte_at = [&](int j) -> int { return (int)prmt(v[j >> 2], 0u, 0x4440u | (uint32_t)(j & 3)); };
	int b[16];
#pragma unroll
	for (int j = 0; j < 16; ++j) b[j] = byte_at(j);
	int sd = b[11] + b[12] + b[13] + b[14] + b[15] - thrb; // window start j = 11: sum - threshold, negative = below the cutoff
	uint32_t fail = (uint32_t)sd >> 31;
#pragma unroll
	for (int j = 10; j >= 0; --j)
	{
		sd += b[j] - b[j + 5];
		fail = __funnelshift_l((uint32_t)sd, fail, 1); // the sign of the newest window enters at bit 0: in the end bit j = window start lo+j
	}
	const uint32_t pass = ~fail & low_bits(min(e - lo - 4, 12)); // window starts lo .. e-5
	if (pass == 0) return -1;
	const int t = 31 - __clz(pass);             // highest passing window start (relative to lo)
	const uint32_t keep = ge & low_bits(t + 5); // bases below the window's end that reach the cutoff (not empty: the window's mean does)
	return lo + 32 - __clz(keep);               // one past the last of them: trailing bases below the cutoff are dropped
}

// the whole search for one read; -2: the general search has to decide (a byte >= 0x80, thresholds outside 1..127)
__device__ __forceinline__ int lane_trim_quality5(const KArgs& A, const uint8_t* qrow, int n)
{
	const int cutb = A.qcut + A.qoff, thrb = A.qthr + 5 * A.qoff;
	if (cutb < 1 || cutb > 127) return -2;
	if (n < 5) return n; // shorter than the window: not trimmed
	// the twelve windows at the 3' end first (most reads end there); only then the jump over a long low-quality tail
	int e = n;
	{
		const int lo = max(e - 16, 0);
		const int r = lane_quality5_block(qrow, lo, e, cutb, thrb);
		if (r != -1) return r;
		if (lo == 0) return 0; // every window failed: the read is emptied
		e = lo + 4;            // what is left: the window starts 0 .. lo-1
	}
	const int g = lane_last_good(qrow, e, cutb);
	if (g == -2) return -2;
	if (g < 0) return 0;  // no base reaches the cutoff, so no window does
	e = min(e, g + 5);    // windows that start above g hold only bases below the cutoff
	for (;;)
	{
		const int lo = max(e - 16, 0);
		const int r = lane_quality5_block(qrow, lo, e, cutb, thrb);
		if (r != -1) return r;
		if (lo == 0) return 0;
		e = lo + 4;
	}
}

// The same with the read's last 16 qualities at hand (`qtail`, device memory, or null; SPG_OPT_QUAL_TAILS: the caller ships them with
// the bases while the quality rows stay in the pinned slot). A read that the adapter steps left alone (n == len >= 16) is decided from
// them whenever one of the twelve windows at its 3' end passes -- most reads; every other read goes to its quality row as before.
__device__ __forceinline__ int lane_trim_quality5_tail(const KArgs& A, const uint8_t* qtail, const uint8_t* qrow, int n, int len)
{
	if (qtail != nullptr && n == len && n >= 16)
	{
		const int cutb = A.qcut + A.qoff, thrb = A.qthr + 5 * A.qoff;
		if (cutb >= 1 && cutb <= 127)
		{
			const uint4 c = __ldg(reinterpret_cast<const uint4*>(qtail));
			const uint32_t v[4] = {c.x, c.y, c.z, c.w};
			const int r = lane_quality5_eval(v, n - 16, n, cutb, thrb);
			if (r >= 0) return r; // -1 (no window of the twelve passes) and -2 (a byte >= 0x80): the search in the row decides
		}
	}
	return lane_trim_quality5(A, qrow, n);
}

// General search: any window <= 8, any trimming point.
// All lanes of the warp call it together (uniform loops, inactive lanes idle). The qualities are read from global memory in blocks
// of 16 window starts: 7 words that cover the positions [P, P + window + 16] go through the lane's shared-memory window `scr`
// (word k at scr + 128k), from which single bytes are picked. Returns the new length.
__device__ __noinline__ int lane_trim_quality(const KArgs& A, const uint8_t* qrow, int n, bool live, uint32_t scr, int done)
{
	const int win = A.qwin;
	int res = live ? n : done; // lanes that are not live keep the result they already have
	live = live && n >= win;   // a read shorter than the window is not trimmed
	int i = n - win; // current window start
	int s = 0;
	for (int b = 0;; ++b)
	{
		if (!__any_sync(kFull, live)) break;
		const int P = n - win - 16 * b - 15; // lowest position this block can touch (may be negative: never read there)
		const uintptr_t a = (uintptr_t)qrow + (intptr_t)P;
		const int boff = (int)(a & 3u);
		const uint32_t* wp = reinterpret_cast<const uint32_t*>(a - (uintptr_t)boff);
		if (live)
		{
#pragma unroll
			for (int k = 0; k < 7; ++k)
			{
				const int pos0 = P - boff + 4 * k; // position of the word's first byte
				uint32_t v = 0;
				if (pos0 + 3 >= 0 && pos0 < n) v = __ldg(wp + k);
				sts_u32(scr + 128u * k, v);
			}
		}
		__syncwarp();
		auto q = [&](int x) -> int {
			const int off = x - P + boff;
			return lds_s8(scr + 128u * (uint32_t)(off >> 2) + (uint32_t)(off & 3)) - A.qoff;
		};
		if (b == 0 && live)
		{
			for (int j = 0; j < win; ++j) s += q(i + j);
		}
		for (int kk = 0; kk < 16; ++kk)
		{
			if (live)
			{
				if (b > 0 || kk > 0) s += q(i) - q(i + win);
				if (s >= A.qthr)
				{
					int nn = i + win; // then drop trailing bases below the cutoff (ends inside the window: its mean reaches the cutoff)
					while (nn > i && q(nn - 1) < A.qcut) --nn;
					res = nn;
					live = false;
				}
				else if (i == 0)
				{
					res = 0; // no window reaches the cutoff: the read is emptied
					live = false;
				}
				else --i;
			}
			if (!__any_sync(kFull, live)) break;
		}
		__syncwarp();
	}
	return res;
}

// FastqEntry::trimN (src/cppNGS/FastqFileStream.cpp:89-117) on the forward N plane of one read: first run of num_n consecutive N
// that lies inside the first `count` bases; returns the new length
template <int NW>
__device__ __forceinline__ int lane_trim_n(const uint32_t* npl, int count, int num_n)
{
	if (count < num_n) return count;
	uint32_t run[NW]; // bit s: positions s .. s+k are all N
	bool any = false;
#pragma unroll
	for (int w = 0; w < NW; ++w)
	{
		run[w] = npl[w];
		any |= run[w] != 0;
	}
	if (!any) return count;
	for (int k = 1; k < num_n; ++k) // AND with the plane shifted right by k bits
	{
		const int wo = k >> 5;
#pragma unroll
		for (int w = 0; w < NW; ++w)
		{
			uint32_t lo = 0, hi = 0;
#pragma unroll
			for (int u = 0; u < NW; ++u)
			{
				if (u == w + wo) lo = npl[u];
				if (u == w + wo + 1) hi = npl[u];
			}
			run[w] &= __funnelshift_r(lo, hi, k);
		}
	}
	int s = -1;
#pragma unroll
	for (int w = NW - 1; w >= 0; --w)
		if (run[w]) s = 32 * w + __ffs((int)run[w]) - 1;
	return (s >= 0 && s <= count - num_n) ? s : count;
}

// ---- lengths from the three steps, quality trimming, [N trimming,] the record: the end of a pair's work, one pair per lane --------------------
// key: best insert candidate or kNoKey; fwd / rev: adapter-only offsets or -1. n1pl / n2pl: forward N planes of the two reads (lanes that may
// hold N), or null. All lanes of the warp call it together; only lanes with `plain` write a record.
template <int FULL, int NW = 1>
__device__ __forceinline__ void lane_finish(const KArgs& A, uint32_t p, bool plain, uint32_t key, int fwd, int rev, const uint32_t* n1pl, const uint32_t* n2pl, uint32_t scr)
{
	int n1 = FULL, n2 = FULL, best_offset = -1;
	uint32_t flags = 0;
	if (key != kNoKey) // insert hit (AnalysisWorker.cpp:269-302)
	{
		best_offset = (int)(key & 0xFFFFu);
		n1 = n2 = FULL - best_offset;
		flags |= SPG_F_INSERT;
	}
	else if (fwd >= 0 || rev >= 0) // adapter-only hit (:410-426)
	{
		flags |= SPG_F_ADAPTER;
		if (fwd >= 0) n1 = fwd;
		if (rev >= 0) n2 = rev;
		if (fwd < 0) n1 = min(n1, rev);
		if (rev < 0) n2 = min(n2, fwd);
	}
	if (A.qcut > 0) // :430-434
	{
		const size_t goff = (size_t)p * A.stride;
		int t1 = -2, t2 = -2; // -2: not decided yet
		if (A.qwin == 5 && plain)
		{
			t1 = lane_trim_quality5_tail(A, A.qt1 ? A.qt1 + 16 * (size_t)p : nullptr, A.q1 + goff, n1, FULL);
			t2 = lane_trim_quality5_tail(A, A.qt2 ? A.qt2 + 16 * (size_t)p : nullptr, A.q2 + goff, n2, FULL);
		}
		if (__any_sync(kFull, plain && t1 < 0)) t1 = lane_trim_quality(A, A.q1 + goff, n1, plain && t1 < 0, scr, t1);
		if (__any_sync(kFull, plain && t2 < 0)) t2 = lane_trim_quality(A, A.q2 + goff, n2, plain && t2 < 0, scr, t2);
		if (t1 < n1) flags |= SPG_F_Q1;
		if (t2 < n2) flags |= SPG_F_Q2;
		n1 = t1;
		n2 = t2;
	}
	if (A.ncut > 0 && n1pl != nullptr && plain) // :437-441, FastqEntry::trimN: the first run of ncut N inside the (trimmed) read cuts it there
	{
		const int t1 = lane_trim_n<NW>(n1pl, n1, A.ncut), t2 = lane_trim_n<NW>(n2pl, n2, A.ncut);
		if (t1 < n1) flags |= SPG_F_N1;
		if (t2 < n2) flags |= SPG_F_N2;
		n1 = t1;
		n2 = t2;
	}
	if (plain)
	{
		uint2 rec;
		rec.x = (uint32_t)n1 | ((uint32_t)n2 << 16);
		rec.y = ((uint32_t)best_offset & 0xFFFFu) | (flags << 16);
		*reinterpret_cast<uint2*>(A.out + p) = rec;
	}
}

// =========================================================================================================================================
// Pairs with N. About 3 % of the pairs of a typical run hold an N somewhere; the warp-cooperative general path costs six times a
// lane's work for each of them. They are collected instead (with everything else that left the main path) and 32 at a time go through
// the same lane-per-pair steps in an N-aware form: rows read from global memory (they are no longer staged), a third plane for N,
// N positions masked out of every comparison, thresholds that hold for any number of remaining positions. A pair that is not
// "two reads of FULL bases made of A/C/G/T/N" still ends in the general path.
// =========================================================================================================================================
constexpr uint32_t kMulN = (1u << 28) | (1u << 19) | (1u << 10) | (1u << 1); // bit 3 of the four bytes ('N' is the only one of ACGTN that has it) -> top nibble
constexpr uint32_t kLutHiN = 0x474EFF54u;                                     // as kLutHi, with index 6 -> 'N'

// 32 bits of a plane held in registers, from bit pos (0 <= pos < 32 NW)
template <int NW>
__device__ __forceinline__ uint32_t lane_reg_extract(const uint32_t (&pl)[NW], int pos)
{
	const int wi = pos >> 5;
	uint32_t lo = 0, hi = 0;
#pragma unroll
	for (int w = 0; w < NW; ++w)
		if (wi == w)
		{
			lo = pl[w];
			hi = (w + 1 < NW) ? pl[w + 1 < NW ? w + 1 : 0] : 0u;
		}
	return __funnelshift_r(lo, hi, pos);
}

// Packs one read from its row in GLOBAL memory (2-byte aligned, FULL bases) into the reversed accumulators of the hi, lo and N plane
// (see lane_pack_rows); returns != 0 if a byte is not one of A/C/G/T/N.
template <int NW, int FULL>
__device__ __noinline__ uint32_t lane_pack_global(const uint8_t* row, uint32_t* acch, uint32_t* accl, uint32_t* accn)
{
	constexpr int NWRD = (FULL + 2 + 3) / 4;
	const uintptr_t ra = (uintptr_t)row;
	const int odd = (int)(ra & 2u);
	const uint32_t* wp = reinterpret_cast<const uint32_t*>(ra & ~(uintptr_t)3);
	uint32_t bad = 0;
#pragma unroll 1
	for (int g = 0; g < NW; ++g)
	{
		uint32_t ah = 0, al = 0, an = 0;
		uint32_t w[8];
#pragma unroll
		for (int k = 0; k < 8; ++k) w[k] = (8 * g + k < NWRD) ? __ldg(wp + 8 * g + k) : 0u; // words behind the row add empty nibbles
#pragma unroll
		for (int k = 0; k < 8; ++k)
		{
			const int lo = 4 * (8 * g + k) - odd; // position of the word's first byte
			al = __funnelshift_l((w[k] & 0x02020202u) * kMulLo, al, 4);
			ah = __funnelshift_l((w[k] & 0x04040404u) * kMulHi, ah, 4);
			an = __funnelshift_l((w[k] & 0x08080808u) * kMulN, an, 4);
			// canonical byte by the low three bits (N included); the selector's spare bit takes the byte's own bit 7, see kLutLo
			uint32_t u;
			asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(u) : "r"(w[k]), "r"(w[k] >> 4), "r"(0x87878787u));
			uint32_t d = prmt(kLutLo, kLutHiN, prmt(u, 0u, 0x4420u)) ^ w[k];
			const int first = max(0, -lo), last = min(4, FULL - lo); // bytes [first, last) of this word belong to the read
			uint32_t m = 0;
			if (last > first) m = (last >= 4 ? 0xFFFFFFFFu : ((1u << (8 * last)) - 1u)) & ~((1u << (8 * first)) - 1u);
			bad |= d & m;
		}
		acch[g] = ah;
		accl[g] = al;
		accn[g] = an;
	}
	return bad;
}

// exact counts of insert offset o with N positions left out (AnalysisWorker.cpp:151-168): mismatches and compared positions
template <int NW, int FULL>
__device__ __forceinline__ void lane_exact_mm_n(int o, const uint32_t (&f1h)[NW], const uint32_t (&f1l)[NW], const uint32_t (&n1)[NW], const uint32_t (&r2h)[NW],
                                                const uint32_t (&r2l)[NW], const uint32_t (&n2r)[NW], int& mm, int& tot)
{
	mm = 0;
	tot = 0;
#pragma unroll
	for (int k = 0; k < NW; ++k)
	{
		const int n = FULL - o - 32 * k;
		if (n > 0)
		{
			const uint32_t sh = lane_reg_extract<NW>(r2h, o + 32 * k), sl = lane_reg_extract<NW>(r2l, o + 32 * k), sn = lane_reg_extract<NW>(n2r, o + 32 * k);
			const uint32_t valid = low_bits(n) & ~(sn | n1[k]);
			mm += __popc((~(sh ^ f1h[k]) | (sl ^ f1l[k])) & valid);
			tot += __popc(valid);
		}
	}
}

// lane_candidate_key with N positions of the reads left out of the adapter fragments
template <int NW, int FULL>
__device__ __forceinline__ uint32_t lane_candidate_key_n(const KArgs& A, int o, int m, int mm, const uint32_t (&f1h)[NW], const uint32_t (&f1l)[NW], const uint32_t (&n1)[NW],
                                                         const uint32_t (&r2h)[NW], const uint32_t (&r2l)[NW], const uint32_t (&n2r)[NW])
{
	int n = m, mis = mm, cnt = m + mm;
	while (cnt >= kRankDim)
	{
		n >>= 1;
		mis >>= 1;
		cnt = n + mis;
	}
	const uint32_t rank = __ldg(&A.ranktab[cnt * kRankDim + n]);
	if (rank == 0xFFFFu) return kNoKey;
	const int alen = min(A.ao, o);
	const int pos = FULL - o;
	const uint32_t v1h = lane_reg_extract<NW>(f1h, pos), v1l = lane_reg_extract<NW>(f1l, pos), v1n = lane_reg_extract<NW>(n1, pos);
	const uint32_t e2h = lane_reg_extract<NW>(r2h, o - alen), e2l = lane_reg_extract<NW>(r2l, o - alen), e2n = lane_reg_extract<NW>(n2r, o - alen);
	const uint32_t v2h = __brev(e2h) >> (32 - alen), v2l = __brev(e2l) >> (32 - alen), v2n = __brev(e2n) >> (32 - alen);
	const uint32_t val1 = low_bits(alen) & ~A.a1n & ~v1n, val2 = low_bits(alen) & ~A.a2n & ~v2n;
	const int mm1 = __popc(((v1h ^ A.a1h) | (v1l ^ A.a1l)) & val1), m1 = __popc(val1) - mm1;
	const int mm2 = __popc(((v2h ^ A.a2h) | (v2l ^ A.a2l)) & val2), m2 = __popc(val2) - mm2;
	if (o < 10)
	{
		const int max_mm = o < 3 ? 0 : (o < 6 ? 1 : 2);
		if (!(mm1 <= max_mm || mm2 <= max_mm)) return kNoKey;
	}
	else
	{
		const double p1 = __ldg(&A.psmall[(m1 + mm1) * (A.ao + 1) + m1]);
		const double p2 = __ldg(&A.psmall[(m2 + mm2) * (A.ao + 1) + m2]);
		if (__dmul_rn(p1, p2) > A.mep) return kNoKey;
	}
	return (rank << 16) | (uint32_t)o;
}

// adapter-only scan of a read that may hold N: every offset, N positions left out of the window (the pass table is indexed by the
// number of compared bases, as in adapter_scan_rounds)
template <int NW, int FULL>
__device__ __forceinline__ int lane_adapter_scan_n(const KArgs& A, const SmemTables& T, const uint32_t (&fh)[NW], const uint32_t (&fl)[NW], const uint32_t (&fn)[NW], uint32_t ah,
                                                   uint32_t al, uint32_t an)
{
	uint32_t best = 0xFFFFFFFFu;
#pragma unroll 1
	for (int r = 0; r < 32; ++r)
	{
#pragma unroll
		for (int q = NW - 1; q >= 0; --q)
		{
			const int cnt = min(A.a_size, FULL - 32 * q - r);
			if (cnt > 0)
			{
				const uint32_t sh = __funnelshift_r(fh[q], (q + 1 < NW) ? fh[q + 1 < NW ? q + 1 : 0] : 0u, r);
				const uint32_t sl = __funnelshift_r(fl[q], (q + 1 < NW) ? fl[q + 1 < NW ? q + 1 : 0] : 0u, r);
				const uint32_t sn = __funnelshift_r(fn[q], (q + 1 < NW) ? fn[q + 1 < NW ? q + 1 : 0] : 0u, r);
				const uint32_t valid = low_bits(cnt) & ~an & ~sn;
				const uint32_t x = (sh ^ ah) | (sl ^ al);
				if ((T.passM[__popc(valid)] >> __popc(x & valid)) & 1u) best = min(best, (uint32_t)(32 * q + r));
			}
		}
	}
	return (int)best;
}

// up to 32 collected pairs (entry e: pair index, len1 | len2 << 16 at list + 8e), one per lane. Returns the lanes whose pair still
// needs the general path.
template <int NW, int FULL>
__device__ __noinline__ uint32_t lane_tile_n(const KArgs& A, const SmemTables& T, const FullTab<NW, FULL>& F, uint32_t list, int n_entries, uint32_t queue, uint32_t scr, int lane)
// (scr: this lane's word 0 of the warp's copy area: indicator planes of the seed scans, window of the general quality search)
{
	uint32_t p = 0;
	bool plain = false;
	if (lane < n_entries)
	{
		const uint2 ent = lds_v2(list + 8u * (uint32_t)lane);
		p = ent.x;
		plain = (ent.y & 0xFFFFu) == (uint32_t)FULL && (ent.y >> 16) == (uint32_t)FULL;
	}
	const bool member = lane < n_entries;
	uint32_t f1h[NW], f1l[NW], n1[NW], r2h[NW], r2l[NW], n2r[NW];
	{
		uint32_t acch[NW], accl[NW], accn[NW];
		const size_t goff = (size_t)p * A.stride;
		const uint32_t odd1 = (uint32_t)((uintptr_t)(A.b1 + goff) & 2u), odd2 = (uint32_t)((uintptr_t)(A.b2 + goff) & 2u);
		uint32_t bad = 0;
		if (plain) bad = lane_pack_global<NW, FULL>(A.b1 + goff, acch, accl, accn);
		lane_forward<NW, FULL>(acch, odd1, f1h);
		lane_forward<NW, FULL>(accl, odd1, f1l);
		lane_forward<NW, FULL>(accn, odd1, n1);
		if (plain) bad |= lane_pack_global<NW, FULL>(A.b2 + goff, acch, accl, accn);
		lane_reversed<NW, FULL>(acch, odd2, r2h);
		lane_reversed<NW, FULL>(accl, odd2, r2l);
		lane_reversed<NW, FULL>(accn, odd2, n2r);
		plain = plain && bad == 0;
	}
	// ---- step 1: the pre-filter counts lo-plane differences at positions where neither read holds an N and compares with the
	// largest limit of any overlap of at most that many bases (F.thr_env): the N positions only shrink the set of compared positions
	int nq = 0;
	{
		uint32_t v1[NW]; // lo plane and "not N" of read 1
#pragma unroll
		for (int w = 0; w < NW; ++w) v1[w] = ~n1[w];
#pragma unroll 1
		for (int r = 0; r < 32; ++r)
		{
			uint32_t sv[NW], svn[NW];
#pragma unroll
			for (int w = 0; w < NW; ++w)
			{
				sv[w] = __funnelshift_r(r2l[w], (w + 1 < NW) ? r2l[w + 1 < NW ? w + 1 : 0] : 0u, r);
				svn[w] = __funnelshift_r(n2r[w], (w + 1 < NW) ? n2r[w + 1 < NW ? w + 1 : 0] : 0u, r);
			}
#pragma unroll
			for (int q = 0; q < NW; ++q)
			{
				int mml = 0;
				const int kmax = min(NW - q, sweep_words(FULL - 32 * q)); // folded after unrolling; see sweep_words
#pragma unroll
				for (int k = 0; k < NW - q; ++k)
				{
					const int w = q + k;
					if (k < kmax) mml += __popc((sv[w] ^ f1l[k]) & v1[k] & ~svn[w] & low_bits(FULL - 32 * w - r));
				}
				if (plain && mml <= (int)F.thr_env[32 * q + r])
				{
					if (nq < kLaneQCap) sts_u16(queue + 64u * (uint32_t)nq, (uint32_t)(32 * q + r));
					++nq;
				}
			}
		}
	}
	if (nq > kLaneQCap) plain = false;
	uint32_t key = kNoKey;
	for (int c = 0; c < kLaneQCap; ++c)
	{
		const bool mine = plain && c < nq;
		if (!__any_sync(kFull, mine)) break;
		if (mine)
		{
			const int o = (int)lds_u16(queue + 64u * (uint32_t)c);
			int mm, tot;
			lane_exact_mm_n<NW, FULL>(o, f1h, f1l, n1, r2h, r2l, n2r, mm, tot);
			if (tot > 0 && tot - mm >= (int)T.mmin[tot]) key = min(key, lane_candidate_key_n<NW, FULL>(A, o, tot - mm, mm, f1h, f1l, n1, r2h, r2l, n2r));
		}
	}
	int fwd = -1, rev = -1;
	uint32_t n2[NW]; // forward N plane of read 2 (scans, trimN)
	lane_flip<NW, FULL>(n2r, n2);
	if (__any_sync(kFull, plain && key == kNoKey))
	{
		uint32_t f2h[NW], f2l[NW];
		lane_flip<NW, FULL>(r2h, f2h);
		lane_flip<NW, FULL>(r2l, f2l);
		if (A.seed_ok)
		{
			fwd = lane_adapter_scan_seeds<NW, FULL, true>(A, T, F, f1h, f1l, n1, scr, A.a1off, A.a1h, A.a1l, A.a1n, A.a1maxmm);
			rev = lane_adapter_scan_seeds<NW, FULL, true>(A, T, F, f2h, f2l, n2, scr, A.a2off, A.a2h, A.a2l, A.a2n, A.a2maxmm);
		}
		else
		{
			fwd = lane_adapter_scan_n<NW, FULL>(A, T, f1h, f1l, n1, A.a1h, A.a1l, A.a1n);
			rev = lane_adapter_scan_n<NW, FULL>(A, T, f2h, f2l, n2, A.a2h, A.a2l, A.a2n);
		}
	}
	lane_finish<FULL, NW>(A, p, plain, key, fwd, rev, n1, n2, scr);
	return __ballot_sync(kFull, member && !plain);
}

// general path for one pair of the tile: the four rows are copied from global memory into the warp's buffer and handed to the
// warp-cooperative code of spg_kernel.cuh
template <int NW>
__device__ __noinline__ void lane_general_pair(const KArgs& A, const SmemTables& T, uint32_t buf, uint32_t p, int len1, int len2, int lane)
{
	const uint32_t rb = ((uint32_t)A.stride + 3u) & ~3u; // rows of the buffer start on word boundaries
	const size_t goff = (size_t)p * A.stride;
	const int halves = A.stride / 2;
	for (int v0 = 0; v0 < halves; v0 += 128) // up to 16 loads of a lane in flight before the first store
	{
		uint32_t x[4][4];
#pragma unroll
		for (int u = 0; u < 4; ++u)
		{
			const int v = v0 + 32 * u + lane;
			const bool in = v < halves;
			x[u][0] = in ? reinterpret_cast<const uint16_t*>(A.b1 + goff)[v] : 0u;
			x[u][1] = in ? reinterpret_cast<const uint16_t*>(A.q1 + goff)[v] : 0u;
			x[u][2] = in ? reinterpret_cast<const uint16_t*>(A.b2 + goff)[v] : 0u;
			x[u][3] = in ? reinterpret_cast<const uint16_t*>(A.q2 + goff)[v] : 0u;
		}
#pragma unroll
		for (int u = 0; u < 4; ++u)
		{
			const int v = v0 + 32 * u + lane;
			if (v < halves)
			{
#pragma unroll
				for (int k = 0; k < 4; ++k) sts_u16(buf + (uint32_t)k * rb + 2u * (uint32_t)v, x[u][k]);
			}
		}
	}
	__syncwarp();
	Pair P;
	P.r1 = buf;
	P.q1 = buf + rb;
	P.r2 = buf + 2 * rb;
	P.q2 = buf + 3 * rb;
	P.len1 = len1;
	P.len2 = len2;
	bool edited = false;
	// FULL = 0 selects the general code, which never touches the per-length tables: any object serves as the reference
	process_pair<NW, 0>(A, T, *reinterpret_cast<const FullTab<NW, 1>*>(&T), P, lane, A.out + p, edited);
	__syncwarp();
}

// ---- the kernel ------------------------------------------------------------------------------------------------------------------------------
// dynamic shared memory: [stages][ read-1 rows | read-2 rows : 32*stride each ] [CW][LaneSmem<NW>]
template <int NW, int FULL, int CW, int MINB>
__global__ void __launch_bounds__((CW + 1) * 32, MINB) trim_lanes_kernel(const __grid_constant__ KArgs A)
{
	static_assert(NW > 0 && FULL >= 52 && FULL <= 32 * NW, "FULL must fit the plane words");
	constexpr int kThreads = (CW + 1) * 32;
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ __align__(8) uint64_t full_bar[kLaneStagesMax];
	__shared__ __align__(8) uint64_t empty_bar[kLaneStagesMax];
	__shared__ SmemTables T;
	__shared__ FullTab<NW, FULL> F;
	__shared__ uint32_t next_it;         // next tile of this CTA to be claimed by a consumer warp
	__shared__ volatile uint32_t issued; // tiles whose copies the producer has issued

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int S = A.stages;
	const uint32_t stage_bytes = lane_stage_bytes(A.stride);
	const uint32_t n_pairs = A.n_dev ? (uint32_t)*A.n_dev : (uint32_t)A.n_pairs;
	const uint32_t n_tiles = (n_pairs + 31u) / 32u;
	const uint32_t n_my = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1u) / gridDim.x : 0u; // tiles blockIdx.x, + gridDim.x, ...
	const uint32_t smem_base = smem_u32(smem);

	for (int i = threadIdx.x; i < SPG_MAXLEN; i += kThreads) T.mmin[i] = A.mmin[i];
	for (int i = threadIdx.x; i < 256; i += kThreads) T.not_acgt[i] = (i == 'A' || i == 'C' || i == 'G' || i == 'T') ? 0 : 1;
	if (threadIdx.x < 21)
	{
		T.passA[threadIdx.x] = A.passA[threadIdx.x];
		uint32_t bm = 0;
		for (int j = 0; j <= (int)threadIdx.x; ++j) bm |= ((A.passA[threadIdx.x] >> (threadIdx.x - j)) & 1u) << j;
		T.passM[threadIdx.x] = bm;
	}
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
	full_tab_init(A, T, F, (int)threadIdx.x, kThreads);
	__syncthreads();

	if (warp == CW)
	{
		// ===== producer: the base rows of 32 pairs per stage, two bulk copies =====
		if (lane == 0)
		{
			for (uint32_t it = 0; it < n_my; ++it)
			{
				const int s = (int)(it % (uint32_t)S);
				const uint32_t round = it / (uint32_t)S;
				if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1u);
				const uint32_t first = (blockIdx.x + it * gridDim.x) * 32u;
				const uint32_t cnt = min(32u, n_pairs - first);
				const uint32_t row_bytes = (cnt * (uint32_t)A.stride + 15u) & ~15u; // see trim_kernel: the planes are readable up to a multiple of 8 rows
				const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
				mbar_arrive_expect_tx(&full_bar[s], 2 * row_bytes);
				const size_t goff = (size_t)first * A.stride;
				bulk_g2s(st, A.b1 + goff, row_bytes, &full_bar[s]);
				bulk_g2s(st + 32u * (uint32_t)A.stride, A.b2 + goff, row_bytes, &full_bar[s]);
				__threadfence_block();
				issued = it + 1u;
			}
		}
		return;
	}

	// ===== consumers: a warp claims the next staged tile of its CTA =====
	const uint32_t wbase = smem_base + (uint32_t)S * stage_bytes + (uint32_t)warp * LaneSmem<NW>::kWarpBytes;
	const uint32_t copy = wbase + 4u * (uint32_t)lane;                                      // word 0 of plane 0 of this lane
	const uint32_t scr = copy;                                                              // quality window of the general search (after the scans)
	const uint32_t queue = wbase + LaneSmem<NW>::kCopyBytes + LaneSmem<NW>::kQualBytes + 2u * (uint32_t)lane; // entry c at + 64c
	const uint32_t rare = wbase + LaneSmem<NW>::kCopyBytes + LaneSmem<NW>::kQualBytes + LaneSmem<NW>::kQueueBytes;
	const uint32_t thr_addr = smem_u32(F.thr);
	int n_rare = 0; // warp-uniform
	// Pairs that are not two full-length reads of A/C/G/T are collected and handed to the general path 32 at a time (its code is
	// large and would otherwise be fetched anew for one pair in every tile)
	auto flush_rare = [&]() {
		uint32_t rest = n_rare >= 32 ? 0xFFFFFFFFu : ((1u << n_rare) - 1u);
		if (n_rare > 0 && A.n_lanes) rest = lane_tile_n<NW, FULL>(A, T, F, rare, n_rare, queue, scr, lane); // pairs with N: one per lane
		while (rest)
		{
			const int e = __ffs((int)rest) - 1;
			rest &= rest - 1;
			const uint2 ent = lds_v2(rare + 8u * (uint32_t)e);
			lane_general_pair<NW>(A, T, wbase, ent.x, (int)(ent.y & 0xFFFFu), (int)(ent.y >> 16), lane);
		}
		n_rare = 0;
	};

	for (;;)
	{
		uint32_t it = 0;
		if (lane == 0) it = atomicAdd(&next_it, 1u);
		it = __shfl_sync(kFull, it, 0);
		if (it >= n_my) break;
		const int s = (int)(it % (uint32_t)S);
		const uint32_t first = (blockIdx.x + it * gridDim.x) * 32u;
		const uint32_t p = first + (uint32_t)lane;
		const bool active = p < n_pairs;
		int len1 = 0, len2 = 0;
		if (active)
		{
			len1 = A.len1[p];
			len2 = A.len2[p];
		}
		bool plain = active && len1 == FULL && len2 == FULL;
		if (active && A.qcut > 0 && !A.quals_on_host) // the last qualities of both reads will be wanted at the end of this tile: start their way into L2 now
		{
			const size_t qoff = (size_t)p * A.stride + (size_t)(FULL - 16);
			asm volatile("prefetch.global.L2 [%0];" ::"l"(A.q1 + qoff));
			asm volatile("prefetch.global.L2 [%0];" ::"l"(A.q2 + qoff));
		}

		// ---- pack both reads from the staged rows, then hand the stage back ----
		uint32_t f1h[NW], f1l[NW], r2l[NW];
		// A parity wait is only meaningful while the barrier is at most one phase behind: claims can run further ahead of the
		// producer than the ring is deep (8 warps, 2-4 stages), so a warp first waits until its tile's copies have been issued --
		// from then on the stage's barrier is in this tile's phase or has just completed it.
		while (issued <= it) __nanosleep(64);
		mbar_wait(&full_bar[s], (it / (uint32_t)S) & 1u);
		{
			const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
			const uint32_t row1 = st + (uint32_t)lane * (uint32_t)A.stride;
			const uint32_t row2 = row1 + 32u * (uint32_t)A.stride;
			const uint32_t bad = lane_pack_rows<NW, FULL>(row1, row2, copy);
			__syncwarp();
			if (lane == 0) // the stage goes back to the producer
			{
				fence_proxy_async();
				mbar_arrive(&empty_bar[s]);
			}
			uint32_t acch[NW], accl[NW], r2h[NW];
			lane_load_acc<NW>(copy, 0, 0, acch);
			lane_load_acc<NW>(copy, 0, 1, accl);
			lane_forward<NW, FULL>(acch, row1 & 2u, f1h);
			lane_forward<NW, FULL>(accl, row1 & 2u, f1l);
			lane_load_acc<NW>(copy, 1, 0, acch);
			lane_load_acc<NW>(copy, 1, 1, accl);
			lane_reversed<NW, FULL>(acch, row2 & 2u, r2h);
			lane_reversed<NW, FULL>(accl, row2 & 2u, r2l);
			plain = plain && bad == 0;
#pragma unroll
			for (int w = 0; w < NW; ++w)
			{
				sts_u32(copy + (0u * (NW + 1) + w) * 128u, f1h[w]);
				sts_u32(copy + (1u * (NW + 1) + w) * 128u, f1l[w]);
				sts_u32(copy + (2u * (NW + 1) + w) * 128u, r2h[w]);
				sts_u32(copy + (3u * (NW + 1) + w) * 128u, r2l[w]);
			}
#pragma unroll
			for (int k = 0; k < 4; ++k) sts_u32(copy + ((uint32_t)k * (NW + 1) + NW) * 128u, 0u);
		}

		// ---- step 1: insert sweep. Pre-filter on the lo plane (a lower bound of the mismatch count), all offsets 32q + r ----
		int nq = 0; // queued survivors
		{
#pragma unroll 1
			for (int r = 0; r < 32; ++r)
			{
				uint32_t sv[NW];
#pragma unroll
				for (int w = 0; w < NW; ++w) sv[w] = __funnelshift_r(r2l[w], (w + 1 < NW) ? r2l[w + 1 < NW ? w + 1 : 0] : 0u, r);
				int mmlq[NW];
				bool any = false;
				const uint32_t ta = thr_addr + 2u * (uint32_t)r;
				static_for<NW>([&](auto qc) {
					constexpr int q = decltype(qc)::value;
					constexpr int KMAX = (NW - q) < sweep_words(FULL - 32 * q) ? (NW - q) : sweep_words(FULL - 32 * q);
					int mml = 0;
#pragma unroll
					for (int k = 0; k < KMAX; ++k)
					{
						const int w = q + k;
						uint32_t x;
						if (32 * w + 62 < FULL) x = sv[w] ^ f1l[k];
						else
						{
							// compared positions: i < FULL - o, i.e. the low FULL - 32w - r bits of this word: a 64-bit constant shifted by r
							const uint64_t M = low_mask64_const(FULL - 32 * w); // folded after unrolling
							x = xor_and(sv[w], f1l[k], __funnelshift_r((uint32_t)M, (uint32_t)(M >> 32), r));
						}
						mml += __popc(x);
					}
					mmlq[q] = mml;
					any |= mml <= lds_s16_const_at<64 * q>(ta); // F.thr[32q + r]
				});
				if (any && plain) // rare: queue the offsets for the exact count
				{
#pragma unroll
					for (int q = 0; q < NW; ++q)
					{
						if (mmlq[q] <= (int)F.thr[32 * q + r])
						{
							if (nq < kLaneQCap) sts_u16(queue + 64u * (uint32_t)nq, (uint32_t)(32 * q + r));
							++nq;
						}
					}
				}
			}
		}
		if (nq > kLaneQCap) plain = false; // low-complexity reads: the general path takes any number of candidates
		uint32_t key = kNoKey;
		for (int c = 0; c < kLaneQCap; ++c)
		{
			const bool mine = plain && c < nq;
			if (!__any_sync(kFull, mine)) break;
			if (mine)
			{
				const int o = (int)lds_u16(queue + 64u * (uint32_t)c);
				const int mm = lane_exact_mm<NW, FULL>(copy, o, f1h, f1l);
				if (mm <= (int)F.thr[o]) key = min(key, lane_candidate_key<NW, FULL>(A, copy, o, FULL - o - mm, mm));
			}
		}

		// ---- steps 2/3: adapter-only scans for pairs without an insert match ----
		int fwd = -1, rev = -1;
		if (__any_sync(kFull, plain && key == kNoKey))
		{
			// read 2 in its original orientation (steps 2/3 scan it forward): the reversed planes the other way round. Formed only
			// now -- batches in which every pair has an insert match never need them, and the sweep has that many registers more
			uint32_t f2h[NW], f2l[NW];
			{
				uint32_t r2h[NW];
#pragma unroll
				for (int w = 0; w < NW; ++w) r2h[w] = lds_u32(copy + (2u * (NW + 1) + w) * 128u);
				lane_flip<NW, FULL>(r2h, f2h);
				lane_flip<NW, FULL>(r2l, f2l);
			}
			if (A.seed_ok)
			{
				fwd = lane_adapter_scan_seeds<NW, FULL>(A, T, F, f1h, f1l, f1h, copy, A.a1off, A.a1h, A.a1l, A.a1n, A.a1maxmm);
				rev = lane_adapter_scan_seeds<NW, FULL>(A, T, F, f2h, f2l, f2h, copy, A.a2off, A.a2h, A.a2l, A.a2n, A.a2maxmm);
			}
			else
			{
				fwd = lane_adapter_scan<NW, FULL>(A, F, f1h, f1l, A.a1h, A.a1l, A.a1maxmm);
				rev = lane_adapter_scan<NW, FULL>(A, F, f2h, f2l, A.a2h, A.a2l, A.a2maxmm);
			}
		}

		// ---- lengths, quality trimming, record ----
		lane_finish<FULL>(A, p, plain, key, fwd, rev, nullptr, nullptr, scr);

		// ---- everything else waits for the general path ----
		const uint32_t rest = __ballot_sync(kFull, active && !plain);
		if (rest)
		{
			if (n_rare + __popc(rest) > 32) flush_rare();
			if (active && !plain) sts_v2(rare + 8u * (uint32_t)(n_rare + __popc(rest & ((1u << lane) - 1u))), p, (uint32_t)len1 | ((uint32_t)len2 << 16));
			n_rare += __popc(rest);
			__syncwarp();
		}
	}
	flush_rare();
}

} // namespace spg
