// spg_kernel.cuh -- sm_100a trimming kernel: one warp per read pair, TMA-staged tiles, bit-plane offset sweep.
//
// What it computes is the per-pair body of AnalysisWorker::run of imgag/ngs-bits
// (src/SeqPurge/AnalysisWorker.cpp:122-441); the numbered steps in the comments are the reference's.
// Nothing here is derived from the reference's code structure: the reference walks bytes offset by offset, this kernel
//   * stages tiles of pairs (ASCII rows, as FASTQ delivers them) into shared memory with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier, a dedicated producer warp, NS-deep ring),
//   * packs every read into bit planes (hi bit, lo bit, N) with warp ballots,
//   * evaluates 32 insert offsets per round -- lane l owns the offsets o with o mod 32 == l, so the funnel-shift amount
//     is the lane id and all word indices are compile-time constants (planes live in registers),
//   * decides with host-built integer tables (minimum matches per overlap length, dense ranks of the match
//     probabilities), so no floating-point function is evaluated on the device and every decision is bit-exact,
//   * handles unusual input (bytes outside ACGTN in read 1, reads longer than the plane path) in a byte-wise path
//     that mirrors the specification directly and doubles as the on-device cross-check of the plane path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/seqpurge_b200.h"

namespace spg
{

constexpr int kConsumerWarps = 8;
constexpr int kThreads = (kConsumerWarps + 1) * 32; // + 1 producer warp
constexpr int kMaxStages = 4;
constexpr uint32_t kNoKey = 0xFFFFFFFFu;
constexpr int kRankDim = 171; // factorial cache holds 0..170 (BasicStatistics.cpp:249-262)

// kernel arguments, passed by value
struct KArgs
{
	// batch (device pointers)
	uint8_t* b1;
	uint8_t* q1;
	uint8_t* b2;
	uint8_t* q2;
	const uint16_t* len1;
	const uint16_t* len2;
	spg_result* out;
	long long n_pairs;
	int stride;     // bytes per row
	int tile_pairs; // pairs per staged tile (multiple of 8)
	int stages;
	// decision tables (device pointers)
	const uint16_t* mmin;    // [1000] minimum #matches for an overlap of T compared bases to pass -match_perc
	const uint16_t* ranktab; // [171*171] dense rank of matchProbability(0.25,n,count), 0xFFFF if > mep
	const double* psmall;    // [(ao+1)*(ao+1)] matchProbability(0.25,n,count) for count<=adapter_overlap
	unsigned long long* ec_m1; // -ec histograms
	unsigned long long* ec_m2;
	unsigned long long* ec_epr;
	// run constants
	double mep;
	int a_size;
	int ao; // adapter_overlap
	int qcut, qwin, qoff, qthr; // qthr: smallest window sum s with (double)s/qwin >= qcut
	int ncut;
	int ec;
	int force_bytewise;
	uint32_t a1h, a1l, a1n; // planes of the first a_size adapter bases
	uint32_t a2h, a2l, a2n;
	uint32_t passA[21]; // [T] bit m: adapter-only hit with m matches out of T compared bases passes (steps 2/3)
	uint8_t a1[32];     // adapter bytes (first 32)
	uint8_t a2[32];
};

// ---- PTX helpers: mbarrier + 1-D bulk copy (TMA) ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	uint32_t ok;
	do
	{
		asm volatile(
		    "{\n\t.reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(ok)
		    : "r"(smem_u32(bar)), "r"(parity)
		    : "memory");
	} while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes),
	             "r"(smem_u32(bar))
	             : "memory");
}

// ---- per-pair view of the staged tile ----------------------------------------------------------------------------------------------
struct Pair
{
	uint8_t* r1; // shared memory rows
	uint8_t* q1;
	uint8_t* r2;
	uint8_t* q2;
	int len1, len2;
};

// per-CTA tables in shared memory
struct SmemTables
{
	uint16_t mmin[SPG_MAXLEN];
	uint32_t passA[21];
};

__device__ __forceinline__ bool is_acgtn(uint32_t c)
{
	// A=0x41 C=0x43 G=0x47 N=0x4E T=0x54: same high bits 010, membership of the low 5 bits by one shift
	const uint32_t M = (1u << 1) | (1u << 3) | (1u << 7) | (1u << 14) | (1u << 20);
	return ((c & 0xE0u) == 0x40u) && ((M >> (c & 31u)) & 1u);
}
__device__ __forceinline__ uint32_t comp_base(uint32_t c) // Sequence::complement for a byte known to be ACGTN
{
	// A<->T, C<->G, N->N
	return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'N';
}

// three-way comparison of the reference: N on either side is "invalid", else match / mismatch
#define SPG_CMP3(b1, b2, m, mm)                 \
	do                                           \
	{                                            \
		if ((b1) != 'N' && (b2) != 'N')          \
		{                                        \
			if ((b1) == (b2)) ++(m);             \
			else ++(mm);                         \
		}                                        \
	} while (0)

// A candidate insert offset that passed the -match_perc filter: probability rank (rejects p > mep) and the
// adapter-presence check of step 1 (AnalysisWorker.cpp:178-259). Rare path (a few % of offsets): byte-wise from the tile.
// Returns (rank << 16) | offset, or kNoKey.
__device__ __noinline__ uint32_t candidate_key(const KArgs& A, const Pair& P, int o, int m, int mm)
{
	// BasicStatistics::matchProbability halves (n, mismatches) until count! fits a double, i.e. count <= 170
	int n = m, mis = mm, cnt = m + mm;
	while (cnt >= kRankDim)
	{
		n >>= 1;
		mis >>= 1;
		cnt = n + mis;
	}
	uint32_t rank = __ldg(&A.ranktab[cnt * kRankDim + n]);
	if (rank == 0xFFFFu) return kNoKey; // p > mep

	int m1 = 0, mm1 = 0;
	{
		int pos = P.len2 - o; // seq1.mid(len2-offset, adapter_overlap)
		int alen = pos < P.len1 ? min(A.ao, P.len1 - pos) : 0;
		for (int i = 0; i < alen; ++i)
		{
			uint32_t x = P.r1[pos + i], y = A.a1[i];
			SPG_CMP3(x, y, m1, mm1);
		}
	}
	int m2 = 0, mm2 = 0;
	{
		int alen = min(o, A.ao); // seq2.left(offset).toReverseComplement().left(adapter_overlap) == R2[len2-offset ..)
		for (int i = 0; i < alen; ++i)
		{
			uint32_t x = P.r2[P.len2 - o + i], y = A.a2[i];
			SPG_CMP3(x, y, m2, mm2);
		}
	}
	if (o < 10)
	{
		int max_mm = o < 3 ? 0 : (o < 6 ? 1 : 2);
		if (!(mm1 <= max_mm || mm2 <= max_mm)) return kNoKey;
	}
	else
	{
		double p1 = __ldg(&A.psmall[(m1 + mm1) * (A.ao + 1) + m1]);
		double p2 = __ldg(&A.psmall[(m2 + mm2) * (A.ao + 1) + m2]);
		if (__dmul_rn(p1, p2) > A.mep) return kNoKey;
	}
	return (rank << 16) | (uint32_t)o;
}

// ---- FastqEntry::trimQuality (src/cppNGS/FastqFileStream.cpp:52-87), warp-parallel over window positions ------------------------------
__device__ __forceinline__ int qual_at(const uint8_t* q, int i, int qoff) { return (int)(signed char)q[i] - qoff; }

__device__ int trim_quality_warp(const KArgs& A, const uint8_t* q, int count, int lane)
{
	const int window = A.qwin;
	if (count < window) return count;
	// highest i in [0, count-window] whose window sum reaches the threshold
	int top = count - window;
	int found = -1;
	for (int base = top & ~31; base >= 0; base -= 32)
	{
		int i = base + lane;
		bool ok = false;
		if (i <= top)
		{
			int s = 0;
			for (int w = 0; w < window; ++w) s += qual_at(q, i + w, A.qoff);
			ok = s >= A.qthr;
		}
		uint32_t b = __ballot_sync(0xffffffffu, ok);
		if (b)
		{
			found = base + 31 - __clz(b);
			break;
		}
	}
	if (found < 0) return 0; // no window reaches the cutoff: read is emptied
	int count_new = found + window;
	// drop trailing bases below the cutoff
	while (count_new > 0)
	{
		int i = count_new - 1 - lane;
		bool low = (i >= 0) && (qual_at(q, i, A.qoff) < A.qcut);
		uint32_t b = __ballot_sync(0xffffffffu, low);
		int run = __ffs(~b) - 1; // number of consecutive low bases from the end; -1 if all 32
		if (run < 0)
		{
			count_new -= 32;
			continue;
		}
		count_new -= run;
		break;
	}
	return max(count_new, 0);
}

// ---- FastqEntry::trimN (src/cppNGS/FastqFileStream.cpp:89-117), warp-parallel over run starts -------------------------------------------
__device__ int trim_n_warp(const uint8_t* r, int count, int num_n, int lane)
{
	if (count < num_n) return count;
	int top = count - num_n;
	for (int base = 0; base <= top; base += 32)
	{
		int s = base + lane;
		bool run = s <= top;
		if (run)
		{
			for (int k = 0; k < num_n; ++k)
			{
				if (r[s + k] != 'N')
				{
					run = false;
					break;
				}
			}
		}
		uint32_t b = __ballot_sync(0xffffffffu, run);
		if (b) return base + __ffs(b) - 1;
	}
	return count;
}

// ---- AnalysisWorker::correctErrors (AnalysisWorker.cpp:19-77), warp-parallel: index i touches r1[i] and r2[count-1-i] only --------------
// returns false if a read-1 byte had to be complemented that the reference cannot complement
__device__ bool correct_errors_warp(const KArgs& A, const Pair& P, int n1, int n2, int lane, bool& newN1, bool& newN2)
{
	const int count = min(n1, n2);
	int mm_count = 0;
	bool bad = false;
	for (int base = 0; base < count; base += 32)
	{
		int i = base + lane;
		bool mism = false;
		if (i < count)
		{
			int i2 = count - 1 - i;
			uint32_t a = P.r1[i], b = P.r2[i2];
			uint32_t cb = comp_base(b);
			if (a != cb)
			{
				mism = true;
				int qa = qual_at(P.q1, i, A.qoff), qb = qual_at(P.q2, i2, A.qoff);
				if (qa > qb)
				{
					if (!is_acgtn(a)) bad = true;
					else
					{
						uint32_t rep = comp_base(a);
						P.r2[i2] = (uint8_t)rep;
						P.q2[i2] = P.q1[i];
						if (rep == 'N') newN2 = true;
						atomicAdd(&A.ec_m2[i2], 1ull);
					}
				}
				else if (qa < qb)
				{
					P.r1[i] = (uint8_t)cb;
					P.q1[i] = P.q2[i2];
					if (cb == 'N') newN1 = true;
					atomicAdd(&A.ec_m1[i], 1ull);
				}
			}
		}
		mm_count += __popc(__ballot_sync(0xffffffffu, mism));
	}
	bad = __any_sync(0xffffffffu, bad);
	newN1 = __any_sync(0xffffffffu, newN1);
	newN2 = __any_sync(0xffffffffu, newN2);
	if (!bad && mm_count > 0 && lane == 0) atomicAdd(&A.ec_epr[mm_count], 1ull);
	__syncwarp();
	return !bad;
}

// ---- byte-wise path: steps 1-3 straight from the staged ASCII rows (any byte values, any length < 1000) ------------------------------------
__device__ uint32_t step1_bytewise(const KArgs& A, const SmemTables& T, const Pair& P, int lane)
{
	const int L = min(P.len1, P.len2);
	uint32_t key = kNoKey;
	for (int o = lane; o < L; o += 32)
	{
		if (o == 0) continue;
		int m = 0, mm = 0;
		for (int j = o; j < L; ++j)
		{
			uint32_t x = P.r1[j - o];
			uint32_t y = comp_base(P.r2[P.len2 - 1 - j]); // seq2[j] of the reference
			SPG_CMP3(x, y, m, mm);
		}
		int tot = m + mm;
		if (tot > 0 && m >= T.mmin[tot]) key = min(key, candidate_key(A, P, o, m, mm));
	}
	return key;
}

__device__ int adapter_scan_bytewise(const KArgs& A, const SmemTables& T, const uint8_t* r, int len, const uint8_t* adapter, int lane)
{
	for (int base = 0; base < len; base += 32)
	{
		int o = base + lane;
		bool pass = false;
		if (o < len)
		{
			int m = 0, mm = 0;
			int cnt = min(A.a_size, len - o);
			for (int i = 0; i < cnt; ++i)
			{
				uint32_t x = r[o + i], y = adapter[i];
				SPG_CMP3(x, y, m, mm);
			}
			pass = (T.passA[m + mm] >> m) & 1u;
		}
		uint32_t b = __ballot_sync(0xffffffffu, pass);
		if (b) return base + __ffs(b) - 1;
	}
	return -1;
}

// ---- bit-plane path ----------------------------------------------------------------------------------------------------------------------
// base code: bit1 of the ASCII byte -> lo plane, bit2 -> hi plane (A=00 C=01 G=11 T=10); complement flips the hi bit only.
template <int NW>
struct Planes
{
	uint32_t h[NW], l[NW], n[NW];
};

// forward planes of one read; reports N presence and bytes outside ACGTN
template <int NW>
__device__ __forceinline__ void pack_forward(const uint8_t* row, int len, int lane, Planes<NW>& pl, bool& hasN, bool& other)
{
	uint32_t anyN = 0, anyOther = 0;
#pragma unroll
	for (int w = 0; w < NW; ++w)
	{
		int pos = 32 * w + lane;
		uint32_t c = pos < len ? row[pos] : (uint32_t)'A';
		pl.h[w] = __ballot_sync(0xffffffffu, c & 4u);
		pl.l[w] = __ballot_sync(0xffffffffu, c & 2u);
		uint32_t special = __ballot_sync(0xffffffffu, (c == 'N') || !is_acgtn(c));
		pl.n[w] = 0;
		if (special) // rare, warp-uniform
		{
			pl.n[w] = __ballot_sync(0xffffffffu, c == 'N');
			anyN |= pl.n[w];
			anyOther |= special & ~pl.n[w];
		}
	}
	hasN = anyN != 0;
	other = anyOther != 0;
}

// planes of revcomp(read 2): position j holds complement(R2[len-1-j])
template <int NW, bool HASN>
__device__ __forceinline__ void pack_revcomp(const uint8_t* row, int len, int lane, Planes<NW>& pl)
{
#pragma unroll
	for (int w = 0; w < NW; ++w)
	{
		int pos = 32 * w + lane;
		uint32_t c = pos < len ? row[len - 1 - pos] : (uint32_t)'T';
		pl.h[w] = __ballot_sync(0xffffffffu, !(c & 4u));
		pl.l[w] = __ballot_sync(0xffffffffu, c & 2u);
		pl.n[w] = HASN ? __ballot_sync(0xffffffffu, c == 'N') : 0u;
	}
}

// step 1 on planes (AnalysisWorker.cpp:137-266)
template <int NW, bool HASN>
__device__ __forceinline__ uint32_t step1_planes(const KArgs& A, const SmemTables& T, const Pair& P, const Planes<NW>& s1, const Planes<NW>& s2, int lane)
{
	const int L = min(P.len1, P.len2);
	uint32_t key = kNoKey;
#pragma unroll
	for (int q = 0; q < NW; ++q)
	{
		if (32 * q < L) // warp-uniform
		{
			const int o = 32 * q + lane;
			const int rem = L - o; // compared positions i in [0, rem): s1[i] vs s2[i+o]
			int mm = 0, nv = 0;
#pragma unroll
			for (int k = 0; k < NW - q; ++k)
			{
				if (32 * (q + k) < L) // warp-uniform: lane 0 still has bases in word k
				{
					const uint32_t hh = (q + k + 1 < NW) ? s2.h[q + k + 1] : 0u;
					const uint32_t hl = (q + k + 1 < NW) ? s2.l[q + k + 1] : 0u;
					const uint32_t xh = __funnelshift_r(s2.h[q + k], hh, lane) ^ s1.h[k];
					const uint32_t xl = __funnelshift_r(s2.l[q + k], hl, lane) ^ s1.l[k];
					const int nb = min(max(rem - 32 * k, 0), 32);
					uint32_t mask = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
					if (HASN)
					{
						const uint32_t hn = (q + k + 1 < NW) ? s2.n[q + k + 1] : 0u;
						mask &= ~(__funnelshift_r(s2.n[q + k], hn, lane) | s1.n[k]);
						nv += __popc(mask);
					}
					mm += __popc((xh | xl) & mask);
				}
			}
			const int tot = HASN ? nv : max(rem, 0);
			const int m = tot - mm;
			if (o >= 1 && rem > 0 && tot > 0 && m >= (int)T.mmin[tot]) key = min(key, candidate_key(A, P, o, m, mm));
		}
	}
	return key;
}

// steps 2/3 on forward planes (AnalysisWorker.cpp:307-353, :355-407): first offset at which the adapter matches
template <int NW, bool HASN>
__device__ __forceinline__ int adapter_scan_planes(const KArgs& A, const SmemTables& T, const Planes<NW>& p, int len, uint32_t ah, uint32_t al, uint32_t an, int lane)
{
#pragma unroll
	for (int q = 0; q < NW; ++q)
	{
		if (32 * q < len) // warp-uniform
		{
			const int o = 32 * q + lane;
			const int cnt = min(A.a_size, len - o); // compared bases (read end cuts the window)
			const uint32_t hh = (q + 1 < NW) ? p.h[q + 1] : 0u;
			const uint32_t hl = (q + 1 < NW) ? p.l[q + 1] : 0u;
			const uint32_t x = (__funnelshift_r(p.h[q], hh, lane) ^ ah) | (__funnelshift_r(p.l[q], hl, lane) ^ al);
			uint32_t valid = cnt > 0 ? (((1u << cnt) - 1u) & ~an) : 0u;
			if (HASN)
			{
				const uint32_t hn = (q + 1 < NW) ? p.n[q + 1] : 0u;
				valid &= ~__funnelshift_r(p.n[q], hn, lane);
			}
			const int tot = __popc(valid);
			const int m = tot - __popc(x & valid);
			const bool pass = cnt > 0 && ((T.passA[tot] >> m) & 1u);
			const uint32_t b = __ballot_sync(0xffffffffu, pass);
			if (b) return 32 * q + __ffs(b) - 1;
		}
	}
	return -1;
}

struct Step123
{
	int best_offset; // -1 none
	int fwd, rev;    // adapter-only offsets, -1 none
};

template <int NW, bool HASN>
__device__ __forceinline__ Step123 steps_planes(const KArgs& A, const SmemTables& T, const Pair& P, const Planes<NW>& f1, const Planes<NW>& f2, int lane)
{
	Step123 r;
	r.fwd = r.rev = -1;
	uint32_t key;
	{
		Planes<NW> s2;
		pack_revcomp<NW, HASN>(P.r2, P.len2, lane, s2);
		key = step1_planes<NW, HASN>(A, T, P, f1, s2, lane);
	}
	key = __reduce_min_sync(0xffffffffu, key);
	r.best_offset = key == kNoKey ? -1 : (int)(key & 0xFFFFu);
	if (r.best_offset < 0)
	{
		r.fwd = adapter_scan_planes<NW, HASN>(A, T, f1, P.len1, A.a1h, A.a1l, A.a1n, lane);
		r.rev = adapter_scan_planes<NW, HASN>(A, T, f2, P.len2, A.a2h, A.a2l, A.a2n, lane);
	}
	return r;
}

// ---- one read pair, one warp ------------------------------------------------------------------------------------------------------------
template <int NW>
__device__ void process_pair(const KArgs& A, const SmemTables& T, const Pair& P, int lane, spg_result* out, bool& edited)
{
	spg_result res;
	res.len1 = res.len2 = 0;
	res.best_offset = -1;
	res.flags = 0;
	res.status = SPG_PAIR_OK;

	const int len1 = P.len1, len2 = P.len2;
	const int maxlen = max(len1, len2);
	bool hasN1 = false, hasN2 = false;
	Step123 st;
	st.best_offset = st.fwd = st.rev = -1;

	bool done = false;
	if (NW > 0 && maxlen <= 32 * NW && !A.force_bytewise)
	{
		Planes<(NW > 0 ? NW : 1)> f1, f2;
		bool other1, other2;
		pack_forward<(NW > 0 ? NW : 1)>(P.r1, len1, lane, f1, hasN1, other1);
		pack_forward<(NW > 0 ? NW : 1)>(P.r2, len2, lane, f2, hasN2, other2);
		if (other2)
		{
			res.status = SPG_PAIR_BAD_BASE_R2;
			done = true;
		}
		else if (!other1)
		{
			if (hasN1 || hasN2) st = steps_planes<(NW > 0 ? NW : 1), true>(A, T, P, f1, f2, lane);
			else st = steps_planes<(NW > 0 ? NW : 1), false>(A, T, P, f1, f2, lane);
			done = true;
		}
	}
	if (!done) // byte-wise path
	{
		bool bad2 = false;
		for (int i = lane; i < len2 && i < A.stride; i += 32)
		{
			uint32_t c = P.r2[i];
			bad2 |= !is_acgtn(c);
			hasN2 |= (c == 'N');
		}
		for (int i = lane; i < len1 && i < A.stride; i += 32) hasN1 |= (P.r1[i] == 'N');
		bad2 = __any_sync(0xffffffffu, bad2);
		hasN1 = __any_sync(0xffffffffu, hasN1);
		hasN2 = __any_sync(0xffffffffu, hasN2);
		if (bad2) res.status = SPG_PAIR_BAD_BASE_R2;
		else if (maxlen >= SPG_MAXLEN || maxlen > A.stride) res.status = SPG_PAIR_TOO_LONG;
		else
		{
			uint32_t key = __reduce_min_sync(0xffffffffu, step1_bytewise(A, T, P, lane));
			st.best_offset = key == kNoKey ? -1 : (int)(key & 0xFFFFu);
			if (st.best_offset < 0)
			{
				st.fwd = adapter_scan_bytewise(A, T, P.r1, len1, A.a1, lane);
				st.rev = adapter_scan_bytewise(A, T, P.r2, len2, A.a2, lane);
			}
		}
	}

	if (res.status == SPG_PAIR_OK)
	{
		int n1 = len1, n2 = len2;
		uint32_t flags = 0;
		if (st.best_offset >= 0) // insert hit (AnalysisWorker.cpp:269-302)
		{
			const int new_length = len2 - st.best_offset;
			n1 = min(n1, new_length);
			n2 = min(n2, new_length);
			flags |= SPG_F_INSERT;
			if (A.ec)
			{
				bool nn1 = false, nn2 = false;
				if (!correct_errors_warp(A, P, n1, n2, lane, nn1, nn2)) res.status = SPG_PAIR_BAD_BASE_EC;
				hasN1 |= nn1;
				hasN2 |= nn2;
				edited = true;
			}
		}
		else if (st.fwd >= 0 || st.rev >= 0) // adapter-only hit (:410-426)
		{
			flags |= SPG_F_ADAPTER;
			if (st.fwd >= 0) n1 = st.fwd;
			if (st.rev >= 0) n2 = st.rev;
			if (st.fwd < 0) n1 = min(n1, st.rev);
			if (st.rev < 0) n2 = min(n2, st.fwd);
		}
		if (res.status == SPG_PAIR_OK)
		{
			if (A.qcut > 0) // :430-434
			{
				int t1 = trim_quality_warp(A, P.q1, n1, lane);
				int t2 = trim_quality_warp(A, P.q2, n2, lane);
				if (t1 < n1) flags |= SPG_F_Q1;
				if (t2 < n2) flags |= SPG_F_Q2;
				n1 = t1;
				n2 = t2;
			}
			if (A.ncut > 0) // :437-441 (a read without any N cannot be cut)
			{
				if (hasN1)
				{
					int t1 = trim_n_warp(P.r1, n1, A.ncut, lane);
					if (t1 < n1) flags |= SPG_F_N1;
					n1 = t1;
				}
				if (hasN2)
				{
					int t2 = trim_n_warp(P.r2, n2, A.ncut, lane);
					if (t2 < n2) flags |= SPG_F_N2;
					n2 = t2;
				}
			}
			res.len1 = (uint16_t)n1;
			res.len2 = (uint16_t)n2;
			res.best_offset = (int16_t)st.best_offset;
			res.flags = (uint8_t)flags;
		}
	}
	if (lane == 0) *out = res;
}

// ---- the kernel: persistent CTAs, producer warp + 8 consumer warps, NS-stage TMA ring ---------------------------------------------------------
// dynamic shared memory: [stages][ b1 | q1 | b2 | q2 : tile_pairs*stride each ][ len1 | len2 : tile_pairs u16 each ]
template <int NW>
__global__ void __launch_bounds__(kThreads) trim_kernel(const __grid_constant__ KArgs A)
{
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ __align__(8) uint64_t full_bar[kMaxStages];
	__shared__ __align__(8) uint64_t empty_bar[kMaxStages];
	__shared__ SmemTables T;

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int TP = A.tile_pairs;
	const size_t plane_bytes = (size_t)TP * A.stride;
	const size_t stage_bytes = 4 * plane_bytes + 4 * (size_t)TP;
	const long long n_tiles = (A.n_pairs + TP - 1) / TP;

	for (int i = threadIdx.x; i < SPG_MAXLEN; i += kThreads) T.mmin[i] = A.mmin[i];
	if (threadIdx.x < 21) T.passA[threadIdx.x] = A.passA[threadIdx.x];
	if (threadIdx.x == 0)
	{
		for (int s = 0; s < A.stages; ++s)
		{
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], kConsumerWarps);
		}
		fence_barrier_init();
	}
	__syncthreads();

	if (warp == kConsumerWarps)
	{
		// ===== producer: one lane issues the bulk copies of each tile =====
		if (lane == 0)
		{
			int it = 0;
			for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it)
			{
				const int s = it % A.stages;
				const uint32_t round = (uint32_t)(it / A.stages);
				if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1u);
				const long long first = t * TP;
				const int cnt = (int)min((long long)TP, A.n_pairs - first);
				const uint32_t row_bytes = (uint32_t)cnt * (uint32_t)A.stride;
				const uint32_t len_bytes = (uint32_t)((cnt + 7) / 8) * 16u;
				uint8_t* st = smem + (size_t)s * stage_bytes;
				mbar_arrive_expect_tx(&full_bar[s], 4 * row_bytes + 2 * len_bytes);
				const size_t goff = (size_t)first * A.stride;
				bulk_g2s(st + 0 * plane_bytes, A.b1 + goff, row_bytes, &full_bar[s]);
				bulk_g2s(st + 1 * plane_bytes, A.q1 + goff, row_bytes, &full_bar[s]);
				bulk_g2s(st + 2 * plane_bytes, A.b2 + goff, row_bytes, &full_bar[s]);
				bulk_g2s(st + 3 * plane_bytes, A.q2 + goff, row_bytes, &full_bar[s]);
				bulk_g2s(st + 4 * plane_bytes, A.len1 + first, len_bytes, &full_bar[s]);
				bulk_g2s(st + 4 * plane_bytes + 2 * (size_t)TP, A.len2 + first, len_bytes, &full_bar[s]);
			}
		}
	}
	else
	{
		// ===== consumers: warp w takes pairs w, w+8, ... of each tile =====
		int it = 0;
		for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it)
		{
			const int s = it % A.stages;
			const uint32_t round = (uint32_t)(it / A.stages);
			mbar_wait(&full_bar[s], round & 1u);
			const long long first = t * TP;
			const int cnt = (int)min((long long)TP, A.n_pairs - first);
			uint8_t* st = smem + (size_t)s * stage_bytes;
			const uint16_t* l1 = reinterpret_cast<const uint16_t*>(st + 4 * plane_bytes);
			const uint16_t* l2 = l1 + TP;
			for (int pr = warp; pr < cnt; pr += kConsumerWarps)
			{
				Pair P;
				P.r1 = st + 0 * plane_bytes + (size_t)pr * A.stride;
				P.q1 = st + 1 * plane_bytes + (size_t)pr * A.stride;
				P.r2 = st + 2 * plane_bytes + (size_t)pr * A.stride;
				P.q2 = st + 3 * plane_bytes + (size_t)pr * A.stride;
				P.len1 = l1[pr];
				P.len2 = l2[pr];
				bool edited = false;
				process_pair<NW>(A, T, P, lane, A.out + first + pr, edited);
				if (edited) // -ec: write the edited rows back (16-byte vectors; rows are 16-byte aligned)
				{
					__syncwarp();
					const size_t goff = (size_t)(first + pr) * A.stride;
					const int vecs = A.stride / 16;
					for (int v = lane; v < vecs; v += 32)
					{
						reinterpret_cast<uint4*>(A.b1 + goff)[v] = reinterpret_cast<const uint4*>(P.r1)[v];
						reinterpret_cast<uint4*>(A.q1 + goff)[v] = reinterpret_cast<const uint4*>(P.q1)[v];
						reinterpret_cast<uint4*>(A.b2 + goff)[v] = reinterpret_cast<const uint4*>(P.r2)[v];
						reinterpret_cast<uint4*>(A.q2 + goff)[v] = reinterpret_cast<const uint4*>(P.q2)[v];
					}
				}
			}
			__syncwarp();
			if (lane == 0)
			{
				fence_proxy_async(); // order this warp's generic-proxy accesses before the next async-proxy refill
				mbar_arrive(&empty_bar[s]);
			}
		}
	}
}

} // namespace spg
