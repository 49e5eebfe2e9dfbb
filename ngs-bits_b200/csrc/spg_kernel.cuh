// spg_kernel.cuh -- sm_100a trimming kernel: one warp per read pair, TMA-staged tiles, bit-plane offset sweep.
//
// What it computes is the per-pair body of AnalysisWorker::run of imgag/ngs-bits
// (src/SeqPurge/AnalysisWorker.cpp:122-441); the numbered steps in the comments are the reference's.
// Nothing here is derived from the reference's code structure: the reference walks bytes offset by offset, this kernel
//   * stages tiles of pairs (ASCII rows, as FASTQ delivers them) into shared memory with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier, a dedicated producer warp, NS-deep ring); the pairs of a staged tile are dealt to the
//     consumer warps round robin,
//   * packs every read into two bit planes (hi/lo bit of a 2-bit base code) with warp ballots; every lane then holds the
//     whole read in registers. Read 2 is packed right-aligned, so that revcomp(read 2) is just the reversed bit string
//     (BREV per word) with the hi plane complemented -- no second pass over its bytes,
//   * evaluates 32 insert offsets per round: lane l owns the offsets o with o mod 32 == l, so the view of revcomp(read 2)
//     it needs is "planes shifted right by l bits" -- NW funnel shifts per plane, formed once -- and word indices are
//     compile-time constants,
//   * decides with host-built integer tables (minimum matches per overlap length with the -mep test folded in, dense ranks
//     of the match probabilities, pass bits of the adapter scans), so no floating-point function is evaluated on the device
//     and every decision is bit-exact,
//   * keeps everything unusual (N bases, bytes outside ACGTN, reads longer than the plane path, -ec, N trimming) in
//     non-inlined functions so that the common path stays small enough for the instruction cache; the byte-wise path
//     mirrors the specification directly and doubles as the on-device cross-check of the plane path,
//   * is additionally compiled for the read length of the run (template parameter FULL): pairs of two full-length reads take
//     steps_full, where position tests are compile-time or per-lane constants, the sweep first bounds the mismatch count from
//     below with one bit plane, and the adapter windows are isolated on the FMA pipe.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/seqpurge_b200.h"


namespace spg
{

constexpr int kMaxStages = 4;
constexpr uint32_t kNoKey = 0xFFFFFFFFu;
constexpr int kRankDim = 171; // factorial cache holds 0..170 (BasicStatistics.cpp:249-262)
constexpr uint32_t kFull = 0xffffffffu;

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
	const int* n_dev; // if not null: the number of pairs is read from device memory (<= n_pairs, which then only sizes the grid)
	int stride;     // bytes per row
	int tile_pairs; // pairs per staged tile (multiple of 8)
	int stages;
	// decision tables (device pointers)
	const uint16_t* mmin;    // [1000] minimum #matches for an overlap of T compared bases to survive the pre-filter
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
	uint32_t a1mask, a2mask; // adapter positions < a_size that are not N
	uint32_t a1pass, a2pass; // bit j: a full-window comparison (all of amask) with j mismatches passes
	int a1maxmm, a2maxmm;    // the same as a limit: most mismatches with which a full window passes (-1: never); only valid if full_ok
	int full_ok;             // adapters without N in their first a_size bases and every pass set of steps 2/3 an interval 0..k of mismatches
	int quals_on_host;       // lane kernel: q1 / q2 point into mapped pinned host memory (zero copy): no speculative prefetches over PCIe
	const uint8_t* qt1;      // lane kernel, or null: [n][16] the last 16 qualities of every read 1 (positions len-16 .. len-1), device memory
	const uint8_t* qt2;
	int n_lanes;             // lane kernel: pairs with N take the N-aware lane path (0: the general path, SPG_OPT_N_LANES)
	int seed_n1_ok;          // lane kernel: the same for a full window that holds one N (a_size-1 compared bases, one block lost)
	int seed_ok;             // lane kernel: every passing window of steps 2/3 has fewer mismatches than complete 4-base adapter blocks (and full_ok)
	uint16_t a1off[20];      // lane kernel: byte offset of the base-indicator plane (A,C,G,T -> 0..3 times (NW+1)*128) of adapter position i
	uint16_t a2off[20];
	uint32_t passA[21]; // [T] bit m: adapter-only hit with m matches out of T compared bases passes (steps 2/3)
	uint8_t a1[32];     // adapter bytes (first 32)
	uint8_t a2[32];
};

// f(std::integral_constant<int, 0>{}), ..., f(std::integral_constant<int, N-1>{}): a loop whose index is a compile-time constant in the
// body (offsets folded into instructions)
template <int N, int I = 0, typename F>
__device__ __forceinline__ void static_for(F&& f)
{
	if constexpr (I < N)
	{
		f(std::integral_constant<int, I>{});
		static_for<N, I + 1>(f);
	}
}

// ---- PTX helpers: shared-window loads/stores, mbarrier, 1-D bulk copy (TMA) --------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}
__device__ __forceinline__ int lds_s8(uint32_t a)
{
	int v;
	asm volatile("ld.shared.s8 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
	return v;
}
// the same with a compile-time byte offset folded into the instruction (no address arithmetic per load)
template <int OFF>
__device__ __forceinline__ uint32_t lds_u8_at(uint32_t a)
{
	uint32_t v;
	asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
	return v;
}
template <int OFF>
__device__ __forceinline__ int lds_s16_at(uint32_t a)
{
	int v;
	asm volatile("ld.shared.s16 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
	return v;
}
template <int OFF>
__device__ __forceinline__ uint2 lds_v2_at(uint32_t a)
{
	uint2 v;
	asm volatile("ld.shared.v2.u32 {%0,%1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(a), "n"(OFF));
	return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

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
	// try_wait suspends the warp in hardware until the phase completes or the time hint expires, so waiting warps do
	// not burn issue slots
	uint32_t ok;
	do
	{
		asm volatile(
		    "{\n\t.reg .pred p;\n\t"
		    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
		    "selp.u32 %0, 1, 0, p;\n\t}"
		    : "=r"(ok)
		    : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
		    : "memory");
	} while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src_gmem), "r"(bytes),
	             "r"(smem_u32(bar))
	             : "memory");
}

// 8-byte store by lane 0 only (predicated, no divergence)
__device__ __forceinline__ void stg_v2_lane0(void* p, uint32_t x, uint32_t y, int lane)
{
	asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %3, 0;\n\t@p st.global.v2.u32 [%0], {%1, %2};\n\t}" ::"l"(p), "r"(x), "r"(y), "r"(lane) : "memory");
}

// ---- per-pair view of the staged tile: shared-window byte addresses of the four rows ---------------------------------------------------
struct Pair
{
	uint32_t r1, q1, r2, q2;
	int len1, len2;
};

// per-CTA tables in shared memory
struct SmemTables
{
	uint16_t mmin[SPG_MAXLEN];
	uint32_t passA[21];
	uint32_t passM[21]; // the same by mismatches: [T] bit j: j mismatches out of T compared bases pass
	uint8_t not_acgt[256]; // 1 for every byte value except 'A','C','G','T' (validity of a base by one shared-memory load)
};

__device__ __forceinline__ bool is_acgtn(uint32_t c)
{
	// A=0x41 C=0x43 G=0x47 N=0x4E T=0x54: same high bits 010, membership of the low 5 bits by one shift
	const uint32_t M = (1u << 1) | (1u << 3) | (1u << 7) | (1u << 14) | (1u << 20);
	return ((c & 0xE0u) == 0x40u) && ((M >> (c & 31u)) & 1u);
}
__device__ __forceinline__ uint32_t comp_base(uint32_t c) // Sequence::complement for a byte known to be ACGTN
{
	return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'N';
}
// ballot of (v & bits) != 0 over the full warp (and + setp fuse into one LOP3 with predicate output)
__device__ __forceinline__ uint32_t ballot_bits(uint32_t v, uint32_t bits)
{
	uint32_t r;
	asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\tvote.sync.ballot.b32 %0, p, 0xffffffff;\n\t}" : "=r"(r) : "r"(v), "r"(bits));
	return r;
}
// (a ^ b) & c in one LOP3
__device__ __forceinline__ uint32_t xor_and(uint32_t a, uint32_t b, uint32_t c)
{
	uint32_t d;
	asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
	return d;
}
// low `nb` bits set, for any int nb (<=0 -> 0, >=32 -> all)
__device__ __forceinline__ uint32_t low_bits(int nb) { return __funnelshift_rc(kFull, 0u, (uint32_t)max(32 - nb, 0)); }

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

// ---- candidates of step 1 ----------------------------------------------------------------------------------------------------------------
// An insert offset that passed the pre-filter: probability rank (rejects p > mep) and the adapter-presence check
// (AnalysisWorker.cpp:178-259). Rare (about one per pair with a true insert match). Returns (rank << 16) | offset, or kNoKey.
// Warp-cooperative: all lanes evaluate the same candidate (o, m, mm); the adapter fragments (at most 32 bytes each) are
// compared one byte per lane and counted with ballots.
__device__ __noinline__ uint32_t candidate_key_warp(const KArgs& A, Pair P, int o, int m, int mm, int lane)
{
	// BasicStatistics::matchProbability halves (n, mismatches) until count! fits a double, i.e. count <= 170
	int n = m, mis = mm, cnt = m + mm;
	while (cnt >= kRankDim)
	{
		n >>= 1;
		mis >>= 1;
		cnt = n + mis;
	}
	const uint32_t rank = __ldg(&A.ranktab[cnt * kRankDim + n]);
	if (rank == 0xFFFFu) return kNoKey; // p > mep

	const int pos = P.len2 - o;                                          // seq1.mid(len2-offset, adapter_overlap)
	const int alen1 = pos < P.len1 ? min(A.ao, P.len1 - pos) : 0;
	const int alen2 = min(o, A.ao);                                      // seq2.left(offset).toReverseComplement().left(..) == R2[len2-offset ..)
	uint32_t x1 = 'N', x2 = 'N';
	if (lane < alen1) x1 = lds_u8(P.r1 + pos + lane);
	if (lane < alen2) x2 = lds_u8(P.r2 + pos + lane);
	const uint32_t y1 = A.a1[lane], y2 = A.a2[lane];
	const bool v1 = lane < alen1 && x1 != 'N' && y1 != 'N';
	const bool v2 = lane < alen2 && x2 != 'N' && y2 != 'N';
	const int m1 = __popc(__ballot_sync(kFull, v1 && x1 == y1));
	const int mm1 = __popc(__ballot_sync(kFull, v1 && x1 != y1));
	const int m2 = __popc(__ballot_sync(kFull, v2 && x2 == y2));
	const int mm2 = __popc(__ballot_sync(kFull, v2 && x2 != y2));
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

// ---- FastqEntry::trimQuality (src/cppNGS/FastqFileStream.cpp:52-87) ---------------------------------------------------------------------
__device__ __forceinline__ int qual_at(uint32_t q, int i, int qoff) { return lds_s8(q + i) - qoff; }

// general form: any window size, one window start per lane
__device__ __noinline__ int trim_quality_slow(const KArgs& A, uint32_t q, int count, int lane)
{
	const int window = A.qwin;
	if (count < window) return count;
	const int top = count - window;
	int found = -1;
	for (int base = top & ~31; base >= 0; base -= 32)
	{
		const int i = base + lane;
		bool ok = false;
		if (i <= top)
		{
			int s = 0;
			for (int w = 0; w < window; ++w) s += qual_at(q, i + w, A.qoff);
			ok = s >= A.qthr;
		}
		const uint32_t b = __ballot_sync(kFull, ok);
		if (b)
		{
			found = base + 31 - __clz(b);
			break;
		}
	}
	return found < 0 ? -1 : found + window;
}

// the complete algorithm for any position of the trimming point (window <= 32 by warp scans, else one window start per lane);
// returns the new length of the read
__device__ __noinline__ int trim_quality_general(const KArgs& A, uint32_t q, int count, int lane)
{
	const int window = A.qwin;
	if (count < window) return count;
	int count_new;
	if (window <= 32)
	{
		// blocks of 32 positions from the 3' end: lane l holds q[base+l]; an inclusive warp scan gives every window sum as a
		// difference of two prefix sums, so the window starts base .. base+32-window are decided per block with one load per lane
		count_new = -1;
		for (int base = count - 32; base + 32 - window >= 0; base -= 33 - window)
		{
			const int i = base + lane;
			const int v = i >= 0 ? qual_at(q, i, A.qoff) : 0;
			int p = v;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) // p += (lane >= d) ? shfl_up(p, d) : 0
				asm volatile("{\n\t.reg .pred g;\n\t.reg .s32 t;\n\tshfl.sync.up.b32 t|g, %0, %1, 0, 0xffffffff;\n\t@g add.s32 %0, %0, t;\n\t}" : "+r"(p) : "r"(d));
			const int s = __shfl_sync(kFull, p, lane + window - 1) - p + v; // q[i] + ... + q[i+window-1] for lane <= 32-window
			const bool ok = i >= 0 && lane <= 32 - window && s >= A.qthr;
			const uint32_t b = __ballot_sync(kFull, ok);
			if (b)
			{
				count_new = base + 31 - __clz(b) + window;
				break;
			}
		}
	}
	else count_new = trim_quality_slow(A, q, count, lane);
	if (count_new < 0) return 0; // no window reaches the cutoff: the read is emptied
	// drop trailing bases below the cutoff
	while (count_new > 0)
	{
		const int i = count_new - 1 - lane;
		const bool low = (i >= 0) && (qual_at(q, i, A.qoff) < A.qcut);
		const uint32_t b = __ballot_sync(kFull, low);
		const int run = __ffs(~b) - 1; // number of consecutive low bases from the end; -1 if all 32
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

// Common case inline: the trimming point lies within the last 32 bases of the read. One load per lane (q[count-32+lane]) serves
// both the window search (warp scan) and the removal of the trailing low-quality bases (bit tricks on one ballot); anything
// else goes to trim_quality_general. Returns the new length of the read (== count if nothing is trimmed).
__device__ __forceinline__ int trim_quality_warp(const KArgs& A, uint32_t q, int count, int lane)
{
	const int window = A.qwin;
	if (count < window) return count;
	if (window > 32) return trim_quality_general(A, q, count, lane);
	const int base = count - 32;
	const int i = base + lane;
	int v = 0;
	if (i >= 0) v = qual_at(q, i, A.qoff);
	int p = v;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
		asm volatile("{\n\t.reg .pred g;\n\t.reg .s32 t;\n\tshfl.sync.up.b32 t|g, %0, %1, 0, 0xffffffff;\n\t@g add.s32 %0, %0, t;\n\t}" : "+r"(p) : "r"(d));
	const int s = __shfl_sync(kFull, p, lane + window - 1) - p + v;
	const uint32_t okm = __ballot_sync(kFull, i >= 0 && lane <= 32 - window && s >= A.qthr);
	const uint32_t low = __ballot_sync(kFull, i >= 0 && v < A.qcut); // bit l: position base+l is below the cutoff
	if (okm == 0) return trim_quality_general(A, q, count, lane);     // trimming point further left (or nowhere): rare
	const int t = 31 - __clz(okm) + window - 1;                       // lane of the last base of the highest passing window
	const uint32_t x = ~low << (31 - t);                              // bit 31 = "base at lane t is not low", then downwards
	if (x == 0) return trim_quality_general(A, q, count, lane);       // low bases all the way to the start of this block: rare
	return base + t + 1 - __clz(x);
}

// Both reads of a pair in one pass: lanes 0-15 hold the last 16 qualities of read 1, lanes 16-31 those of read 2 (segmented
// warp scan of width 16). Covers trimming points within the last 17-window bases of each read; a read that needs more (or a
// window > 8, or a read shorter than the window) takes trim_quality_warp / trim_quality_general.
// The core is free of branches for any input (results are only meaningful for window <= 8 <= n1, n2), so that it can also be issued
// ahead of time for the untrimmed lengths and overlap with the packing of the bases. r1 / r2: new length, or -1 = not decided here.
template <bool SCAN>
__device__ __forceinline__ void trim_quality_pair_core(const KArgs& A, const Pair& P, int n1, int n2, int lane, int& r1, int& r2)
{
	const int window = A.qwin;
	const bool second = lane >= 16;
	const int hl = lane & 15;
	const int n_r = second ? n2 : n1; // the read this half warp works on
	const int base = n_r - 16;
	const int i = base + hl;
	int v = 0;
	if (i >= 0) v = qual_at(second ? P.q2 : P.q1, i, A.qoff);
	int s; // window sum q[i] + ... + q[i+window-1], valid for hl <= 16-window
	if (!SCAN) // window == 5, the default: four independent shuffles instead of a scan (no serial chain)
	{
		const int d1 = __shfl_down_sync(kFull, v, 1, 16), d2 = __shfl_down_sync(kFull, v, 2, 16);
		const int d3 = __shfl_down_sync(kFull, v, 3, 16), d4 = __shfl_down_sync(kFull, v, 4, 16);
		s = (v + d1 + d2) + (d3 + d4);
	}
	else
	{
		int p = v;
#pragma unroll
		for (int d = 1; d < 16; d <<= 1) // inclusive scan inside each half warp (c = (32-16)<<8: segment width 16)
			asm volatile("{\n\t.reg .pred g;\n\t.reg .s32 t;\n\tshfl.sync.up.b32 t|g, %0, %1, 0x1000, 0xffffffff;\n\t@g add.s32 %0, %0, t;\n\t}" : "+r"(p) : "r"(d));
		s = __shfl_sync(kFull, p, hl + window - 1, 16) - p + v;
	}
	const uint32_t okm = __ballot_sync(kFull, i >= 0 && hl <= 16 - window && s >= A.qthr);
	const uint32_t low = __ballot_sync(kFull, i >= 0 && v < A.qcut);
	// every lane decodes the votes of its own half (one instruction stream for both reads), lanes 0 and 16 hold the results
	const uint32_t sh = second ? 16u : 0u;
	const uint32_t ok_r = (okm >> sh) & 0xFFFFu;
	const uint32_t low_r = low >> sh;
	const int t = (31 - __clz(ok_r | 1u) + window - 1) & 31; // index (0..15) of the last base of the highest passing window
	const uint32_t x = ~low_r << (31 - t);                   // bit 31 = "base t is not low", then downwards; bits above t fall out
	int res = n_r - 16 + t + 1 - __clz(x | 1u);
	if (ok_r == 0 || x == 0) res = -1;                       // trimming point further left: rare
	r1 = __shfl_sync(kFull, res, 0);
	r2 = __shfl_sync(kFull, res, 16);
}

__device__ __forceinline__ void trim_quality_pair(const KArgs& A, const Pair& P, int n1, int n2, int lane, int& t1, int& t2)
{
	const int window = A.qwin;
	if (window > 8 || n1 < window || n2 < window)
	{
		t1 = trim_quality_warp(A, P.q1, n1, lane);
		t2 = trim_quality_warp(A, P.q2, n2, lane);
		return;
	}
	if (window == 5) trim_quality_pair_core<false>(A, P, n1, n2, lane, t1, t2);
	else trim_quality_pair_core<true>(A, P, n1, n2, lane, t1, t2);
	if (t1 < 0) t1 = trim_quality_general(A, P.q1, n1, lane);
	if (t2 < 0) t2 = trim_quality_general(A, P.q2, n2, lane);
}

// ---- FastqEntry::trimN (src/cppNGS/FastqFileStream.cpp:89-117), warp-parallel over run starts; only reads that hold an N get here ----------
__device__ __noinline__ int trim_n_warp(uint32_t r, int count, int num_n, int lane)
{
	if (count < num_n) return count;
	const int top = count - num_n;
	for (int base = 0; base <= top; base += 32)
	{
		const int s = base + lane;
		bool run = s <= top;
		if (run)
		{
			for (int k = 0; k < num_n; ++k)
			{
				if (lds_u8(r + s + k) != 'N')
				{
					run = false;
					break;
				}
			}
		}
		const uint32_t b = __ballot_sync(kFull, run);
		if (b) return base + __ffs(b) - 1;
	}
	return count;
}

// ---- AnalysisWorker::correctErrors (AnalysisWorker.cpp:19-77), warp-parallel: index i touches r1[i] and r2[count-1-i] only --------------
// returns false if a read-1 byte had to be complemented that the reference cannot complement
__device__ __noinline__ bool correct_errors_warp(const KArgs& A, Pair P, int n1, int n2, int lane, bool& newN1, bool& newN2)
{
	const int count = min(n1, n2);
	int mm_count = 0;
	bool bad = false, nn1 = false, nn2 = false;
	for (int base = 0; base < count; base += 32)
	{
		const int i = base + lane;
		bool mism = false;
		if (i < count)
		{
			const int i2 = count - 1 - i;
			const uint32_t a = lds_u8(P.r1 + i), b = lds_u8(P.r2 + i2);
			const uint32_t cb = comp_base(b);
			if (a != cb)
			{
				mism = true;
				const int qa = qual_at(P.q1, i, A.qoff), qb = qual_at(P.q2, i2, A.qoff);
				if (qa > qb)
				{
					if (!is_acgtn(a)) bad = true;
					else
					{
						const uint32_t rep = comp_base(a);
						sts_u8(P.r2 + i2, rep);
						sts_u8(P.q2 + i2, lds_u8(P.q1 + i));
						if (rep == 'N') nn2 = true;
						atomicAdd(&A.ec_m2[i2], 1ull);
					}
				}
				else if (qa < qb)
				{
					sts_u8(P.r1 + i, cb);
					sts_u8(P.q1 + i, lds_u8(P.q2 + i2));
					if (cb == 'N') nn1 = true;
					atomicAdd(&A.ec_m1[i], 1ull);
				}
			}
		}
		mm_count += __popc(__ballot_sync(kFull, mism));
	}
	bad = __any_sync(kFull, bad);
	newN1 = __any_sync(kFull, nn1);
	newN2 = __any_sync(kFull, nn2);
	if (!bad && mm_count > 0 && lane == 0) atomicAdd(&A.ec_epr[mm_count], 1ull);
	__syncwarp();
	return !bad;
}

struct Step123
{
	int best_offset; // insert-match offset, -1 none
	int fwd, rev;    // adapter-only offsets, -1 none
};
// result of the out-of-line paths (returned by value so that the common path keeps everything in registers)
struct RareSteps
{
	Step123 st;
	int status; // SPG_PAIR_*
	int hasN;   // bit 0: read 1 holds an N, bit 1: read 2
};

// ---- byte-wise path: steps 1-3 straight from the staged ASCII rows (any byte values, any length < 1000) ------------------------------------
__device__ __forceinline__ uint32_t candidate_key_lane(const KArgs& A, const Pair& P, int o, int m, int mm)
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
	int m1 = 0, mm1 = 0, m2 = 0, mm2 = 0;
	const int pos = P.len2 - o;
	const int alen1 = pos < P.len1 ? min(A.ao, P.len1 - pos) : 0;
	for (int i = 0; i < alen1; ++i)
	{
		const uint32_t x = lds_u8(P.r1 + pos + i), y = A.a1[i];
		SPG_CMP3(x, y, m1, mm1);
	}
	const int alen2 = min(o, A.ao);
	for (int i = 0; i < alen2; ++i)
	{
		const uint32_t x = lds_u8(P.r2 + pos + i), y = A.a2[i];
		SPG_CMP3(x, y, m2, mm2);
	}
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

__device__ int adapter_scan_bytewise(const KArgs& A, const SmemTables& T, uint32_t r, int len, const uint8_t* adapter, int lane)
{
	for (int base = 0; base < len; base += 32)
	{
		const int o = base + lane;
		bool pass = false;
		if (o < len)
		{
			int m = 0, mm = 0;
			const int cnt = min(A.a_size, len - o);
			for (int i = 0; i < cnt; ++i)
			{
				const uint32_t x = lds_u8(r + o + i), y = adapter[i];
				SPG_CMP3(x, y, m, mm);
			}
			pass = (T.passA[m + mm] >> m) & 1u;
		}
		const uint32_t b = __ballot_sync(kFull, pass);
		if (b) return base + __ffs(b) - 1;
	}
	return -1;
}

// read 2 is known to be ACGTN here
__device__ __noinline__ Step123 steps_bytewise(const KArgs& A, const SmemTables& T, Pair P, int lane)
{
	Step123 st;
	st.fwd = st.rev = -1;
	const int L = min(P.len1, P.len2);
	uint32_t key = kNoKey;
	for (int o = lane; o < L; o += 32)
	{
		if (o == 0) continue;
		int m = 0, mm = 0;
		for (int j = o; j < L; ++j)
		{
			const uint32_t x = lds_u8(P.r1 + j - o);
			const uint32_t y = comp_base(lds_u8(P.r2 + P.len2 - 1 - j)); // seq2[j] of the reference
			SPG_CMP3(x, y, m, mm);
		}
		const int tot = m + mm;
		if (tot > 0 && m >= T.mmin[tot]) key = min(key, candidate_key_lane(A, P, o, m, mm));
	}
	key = __reduce_min_sync(kFull, key);
	st.best_offset = key == kNoKey ? -1 : (int)(key & 0xFFFFu);
	if (st.best_offset < 0)
	{
		st.fwd = adapter_scan_bytewise(A, T, P.r1, P.len1, A.a1, lane);
		st.rev = adapter_scan_bytewise(A, T, P.r2, P.len2, A.a2, lane);
	}
	return st;
}

// ---- bit-plane path ----------------------------------------------------------------------------------------------------------------------
// base code: bit1 of the ASCII byte -> lo plane, bit2 -> hi plane (A=00 C=01 G=11 T=10); complement flips the hi bit only.
// Bit b of word w is position 32*w+b. Positions beyond the read are 0 in every plane.
template <int NW>
struct Planes
{
	uint32_t h[NW], l[NW];
};

// forward planes of one read; returns (per lane, in the low byte) whether one of its bytes is not A/C/G/T.
// D = 0: left aligned (bit b of word w is position 32*w+b). D = 32*NW-len: right aligned (position p sits at bit p+D, the read
// ends at the top of the last word) -- used for read 2, see planes_revcomp_shifted and adapter_scan_right.
template <int NW>
__device__ __forceinline__ uint32_t pack_forward(const SmemTables& T, uint32_t row, int len, int D, int lane, Planes<NW>& pl)
{
	uint32_t bad = 0; // != 0: one of this lane's bytes is not A/C/G/T
	const uint32_t lut = smem_u32(T.not_acgt);
#pragma unroll
	for (int w = 0; w < NW; ++w)
	{
		uint32_t c = 'A';
		const int pos = 32 * w + lane - D;
		if ((unsigned)pos < (unsigned)len) c = lds_u8(row + pos);
		pl.h[w] = ballot_bits(c, 4u);
		pl.l[w] = ballot_bits(c, 2u);
		bad |= lds_u8(lut + c); // table lookup instead of arithmetic: keeps the ALU pipe for the sweep
	}
	return bad;
}

// N plane of one read (same alignment rule as pack_forward) and whether it holds bytes outside ACGTN (only for pairs in which
// pack_forward saw something unusual)
template <int NW>
__device__ __forceinline__ void pack_special(uint32_t row, int len, int D, int lane, uint32_t (&n)[NW], bool& hasN, bool& other)
{
	uint32_t anyN = 0, anyOther = 0;
#pragma unroll
	for (int w = 0; w < NW; ++w)
	{
		uint32_t c = 'A';
		const int pos = 32 * w + lane - D;
		if ((unsigned)pos < (unsigned)len) c = lds_u8(row + pos);
		n[w] = __ballot_sync(kFull, c == 'N');
		anyN |= n[w];
		anyOther |= __ballot_sync(kFull, !is_acgtn(c));
	}
	hasN = anyN != 0;
	other = anyOther != 0;
}

// Planes of revcomp(read 2) from the RIGHT-ALIGNED forward planes of read 2, without touching the bytes again: reversing the
// whole NW*32-bit string (word order + BREV) puts complement-position j = len-1-p at bit j exactly because the read ends at the
// top bit; the complement flips the hi plane. Returned already shifted right by `lane` bits: word w holds positions 32*w+lane ..
// 32*w+lane+31 -- lane l owns every offset o = 32*q + l, so these NW words are all the windows of revcomp(read 2) it ever needs.
// Positions >= len hold padding (hi plane 1): they are never compared (masks of step 1 stop at min(len1,len2)).
template <int NW>
__device__ __forceinline__ void planes_revcomp_shifted(const Planes<NW>& f2r, int lane, Planes<NW>& s)
{
	uint32_t h[NW], l[NW];
#pragma unroll
	for (int w = 0; w < NW; ++w)
	{
		h[w] = ~__brev(f2r.h[NW - 1 - w]);
		l[w] = __brev(f2r.l[NW - 1 - w]);
	}
#pragma unroll
	for (int w = 0; w < NW; ++w)
	{
		s.h[w] = __funnelshift_r(h[w], (w + 1 < NW) ? h[w + 1] : 0u, lane);
		s.l[w] = __funnelshift_r(l[w], (w + 1 < NW) ? l[w + 1] : 0u, lane);
	}
}
template <int NW>
__device__ __forceinline__ void n_revcomp_shifted(const uint32_t (&n2r)[NW], int lane, uint32_t (&sn)[NW])
{
	uint32_t n[NW];
#pragma unroll
	for (int w = 0; w < NW; ++w) n[w] = __brev(n2r[NW - 1 - w]);
#pragma unroll
	for (int w = 0; w < NW; ++w) sn[w] = __funnelshift_r(n[w], (w + 1 < NW) ? n[w + 1] : 0u, lane);
}

template <int NW>
__device__ __forceinline__ void shift_words(const uint32_t (&p)[NW], int lane, uint32_t (&s)[NW])
{
#pragma unroll
	for (int w = 0; w < NW; ++w) s[w] = __funnelshift_r(p[w], (w + 1 < NW) ? p[w + 1] : 0u, lane);
}

// step 1 on planes (AnalysisWorker.cpp:137-266). s1: forward planes of read 1; s2s: planes of revcomp(read 2) shifted by lane.
// For offset o = 32*q+lane and word k of read 1 the partner word of read 2 is s2s[q+k], and the mask of compared positions
// (i < L-o) is low_bits(L - 32*(q+k) - lane): both depend on q+k only, so they are formed once per pair.
// n1 / n2s: N planes (read 1 / revcomp(read 2) shifted), only read when HASN.
template <int NW, bool HASN>
__device__ __forceinline__ int step1_planes(const KArgs& A, const SmemTables& T, const Pair& P, const Planes<NW>& s1, const Planes<NW>& s2s,
                                            const uint32_t (&n1)[NW], const uint32_t (&n2s)[NW], int lane)
{
	const int L = min(P.len1, P.len2);
	// mw[w] = low_bits(L - 32*w - lane), formed as the warp-uniform "position < L" bit string funnel-shifted by lane
	uint32_t mw[NW];
#pragma unroll
	for (int w = 0; w < NW; ++w) mw[w] = __funnelshift_r(low_bits(L - 32 * w), (w + 1 < NW) ? low_bits(L - 32 * (w + 1)) : 0u, lane);
	// survivors of the pre-filter are rare: every lane notes its own (bit q of smask) and the warp votes once after the sweep
	uint32_t smask = 0;
	int mmq[NW], totq[NW];
#pragma unroll
	for (int q = 0; q < NW; ++q)
	{
		int mm = 0, nv = 0;
#pragma unroll
		for (int k = 0; k < NW - q; ++k)
		{
			uint32_t valid = mw[q + k];
			if (HASN)
			{
				valid &= ~(n2s[q + k] | n1[k]);
				nv += __popc(valid);
			}
			mm += __popc(((s2s.h[q + k] ^ s1.h[k]) | (s2s.l[q + k] ^ s1.l[k])) & valid);
		}
		const int o = 32 * q + lane;
		const int tot = HASN ? nv : max(L - o, 0);
		mmq[q] = mm;
		if (HASN) totq[q] = tot;
		if (o >= 1 && tot > 0 && tot - mm >= (int)T.mmin[tot]) smask |= 1u << q;
	}
	if (ballot_bits(smask, kFull) == 0) return -1; // the common case for pairs without an insert match
	uint32_t key = kNoKey; // warp-uniform
#pragma unroll
	for (int q = 0; q < NW; ++q)
	{
		uint32_t b = __ballot_sync(kFull, (smask >> q) & 1u);
		while (b)
		{
			const int src = __ffs(b) - 1;
			b &= b - 1;
			const int mm = __shfl_sync(kFull, mmq[q], src);
			const int tot = HASN ? __shfl_sync(kFull, totq[q], src) : max(L - 32 * q - src, 0);
			key = min(key, candidate_key_warp(A, P, 32 * q + src, tot - mm, mm, lane));
		}
	}
	return key == kNoKey ? -1 : (int)(key & 0xFFFFu);
}

// steps 2/3 on forward planes (AnalysisWorker.cpp:307-353, :355-407): first offset at which the adapter matches.
// sh/sl(/sn): left-aligned forward planes of the read shifted by lane. Each lane notes its passing rounds; one vote per read.
// QF = number of leading rounds in which every lane has the full adapter window inside the read (32*q+31+a_size <= len): those
// need no per-lane mask. The caller dispatches on QF once per read, so the rounds themselves are free of branches.
// amask: the a_size adapter positions that are not N; pass_by_mm: bit j set = a full window with j mismatches passes.
template <int NW, bool HASN, int QF>
__device__ __forceinline__ uint32_t adapter_scan_rounds(const KArgs& A, const SmemTables& T, const uint32_t (&sh)[NW], const uint32_t (&sl)[NW], const uint32_t (&sn)[NW],
                                                        int len, uint32_t ah, uint32_t al, uint32_t an, uint32_t amask, uint32_t pass_by_mm, int lane)
{
	uint32_t pm = 0; // bit q: this lane's offset 32*q+lane passes
#pragma unroll
	for (int q = NW - 1; q >= 0; --q)
	{
		uint32_t bit;
		const uint32_t x = (sh[q] ^ ah) | (sl[q] ^ al);
		if (q < QF && !HASN) bit = pass_by_mm >> __popc(x & amask);
		else
		{
			const int cnt = min(A.a_size, len - 32 * q - lane); // compared bases (the read end cuts the window); <= 0: none
			uint32_t valid = low_bits(cnt) & ~an;
			if (HASN) valid &= ~sn[q];
			bit = cnt > 0 ? T.passM[__popc(valid)] >> __popc(x & valid) : 0u;
		}
		pm = pm * 2u + (bit & 1u);
	}
	return pm;
}

template <int NW, bool HASN>
__device__ __forceinline__ int adapter_scan_planes(const KArgs& A, const SmemTables& T, const uint32_t (&sh)[NW], const uint32_t (&sl)[NW], const uint32_t (&sn)[NW],
                                                   int len, uint32_t ah, uint32_t al, uint32_t an, uint32_t amask, uint32_t pass_by_mm, int lane)
{
	const int full = HASN ? 0 : (len - 31 - A.a_size >= 0 ? ((len - 31 - A.a_size) >> 5) + 1 : 0); // warp-uniform
	uint32_t pm;
#define SPG_SCAN(QF) pm = adapter_scan_rounds<NW, HASN, (QF) <= NW ? (QF) : NW>(A, T, sh, sl, sn, len, ah, al, an, amask, pass_by_mm, lane)
	if (NW > 10) SPG_SCAN(0); // long reads (16 / 32 plane words): every round takes the masked form, no dispatch on the read length
	else switch (full)
	{
		case 0: SPG_SCAN(0); break;
		case 1: SPG_SCAN(1); break;
		case 2: SPG_SCAN(2); break;
		case 3: SPG_SCAN(3); break;
		case 4: SPG_SCAN(4); break;
		case 5: SPG_SCAN(5); break;
		case 6: SPG_SCAN(6); break;
		case 7: SPG_SCAN(7); break;
		case 8: SPG_SCAN(8); break;
		case 9: SPG_SCAN(9); break;
		default: SPG_SCAN(10); break;
	}
#undef SPG_SCAN
	if (ballot_bits(pm, kFull) == 0) return -1; // the common case: one vote per read
	// lowest passing offset over all lanes
	const uint32_t mine = pm ? (uint32_t)(32 * (__ffs(pm) - 1) + lane) : 0xFFFFFFFFu;
	return (int)__reduce_min_sync(kFull, mine);
}

// step 3 on the RIGHT-ALIGNED forward planes of read 2 (shifted by lane): offset o sits at bit k = o + D, the read ends at bit
// 32*NW, so which rounds see the full adapter window is known at compile time (all but the last) and the last round's window
// length min(a_size, 32-lane) does not depend on the read at all. No warp-uniform branches, one vote.
template <int NW, bool HASN>
__device__ __forceinline__ int adapter_scan_right(const KArgs& A, const SmemTables& T, const uint32_t (&sh)[NW], const uint32_t (&sl)[NW], const uint32_t (&sn)[NW],
                                                  int D, uint32_t ah, uint32_t al, uint32_t an, uint32_t amask, uint32_t pass_by_mm, int lane)
{
	uint32_t pm = 0;
#pragma unroll
	for (int q = NW - 1; q >= 0; --q)
	{
		uint32_t bit;
		const uint32_t x = (sh[q] ^ ah) | (sl[q] ^ al);
		if (q < NW - 1 && !HASN) bit = pass_by_mm >> __popc(x & amask);
		else
		{
			uint32_t valid = (q < NW - 1) ? amask : (low_bits(min(A.a_size, 32 - lane)) & ~an);
			if (HASN) valid &= ~sn[q];
			bit = T.passM[__popc(valid)] >> __popc(x & valid);
		}
		pm = pm * 2u + (bit & 1u);
	}
	// bits k < D are padding in front of the read: drop the rounds that start there (32*q + lane < D)
	pm &= ~low_bits((D - lane + 31) >> 5);
	if (ballot_bits(pm, kFull) == 0) return -1;
	const uint32_t mine = pm ? (uint32_t)(32 * (__ffs(pm) - 1) + lane - D) : 0xFFFFFFFFu;
	return (int)__reduce_min_sync(kFull, mine);
}

// f1: forward planes of read 1 (left aligned); f2r: forward planes of read 2, right aligned by D2 = 32*NW - len2.
template <int NW, bool HASN>
__device__ __forceinline__ Step123 steps_planes(const KArgs& A, const SmemTables& T, const Pair& P, const Planes<NW>& f1, const Planes<NW>& f2r, int D2,
                                                const uint32_t (&n1)[NW], const uint32_t (&n2r)[NW], int lane)
{
	Step123 r;
	r.fwd = r.rev = -1;
	{
		Planes<NW> s2s;
		uint32_t n2s[NW];
		planes_revcomp_shifted<NW>(f2r, lane, s2s);
		if (HASN) n_revcomp_shifted<NW>(n2r, lane, n2s);
		r.best_offset = step1_planes<NW, HASN>(A, T, P, f1, s2s, n1, n2s, lane);
	}
	if (r.best_offset < 0)
	{
		uint32_t sh[NW], sl[NW], sn[NW];
		shift_words<NW>(f1.h, lane, sh);
		shift_words<NW>(f1.l, lane, sl);
		if (HASN) shift_words<NW>(n1, lane, sn);
		r.fwd = adapter_scan_planes<NW, HASN>(A, T, sh, sl, sn, P.len1, A.a1h, A.a1l, A.a1n, A.a1mask, A.a1pass, lane);
		shift_words<NW>(f2r.h, lane, sh);
		shift_words<NW>(f2r.l, lane, sl);
		if (HASN) shift_words<NW>(n2r, lane, sn);
		r.rev = adapter_scan_right<NW, HASN>(A, T, sh, sl, sn, D2, A.a2h, A.a2l, A.a2n, A.a2mask, A.a2pass, lane);
	}
	return r;
}

// ---- full-length fast path ------------------------------------------------------------------------------------------------------------
// Reads of one sequencing run share one length before trimming. trim_kernel<.., FULL> is compiled for that length: for pairs with
// len1 == len2 == FULL (and only A/C/G/T) every position test, every window mask of the sweep and every "does the adapter window
// fit" question is a compile-time constant or a per-lane constant, which the CTA tabulates once in shared memory (FullTab).
// Any other pair of the batch takes the general path above; the results are the same by construction and are cross-checked
// against the byte-wise path in the tests.
__host__ __device__ constexpr int full_qf(int FULL) { return FULL >= 51 ? ((FULL - 51) >> 5) + 1 : 0; } // rounds of the read-1 scan whose 20-base window fits for every lane

template <int NW, int FULL>
struct FullTab
{
	static constexpr int QF = full_qf(FULL) < NW ? full_qf(FULL) : NW;
	static constexpr int NT = NW - QF > 0 ? NW - QF : 1;
	int16_t thr[32 * NW]; // [o]: most mismatches with which insert offset o survives the pre-filter (mmin), -1: never
	int16_t thr_env[32 * NW]; // [o]: the largest thr of any overlap of at most FULL-o compared bases (reads with N: fewer positions are compared)
	// adapter scans: a window of cnt compared bases is "x << (32-cnt)" (a multiplication by .x = 2^(32-cnt), which runs on the FMA
	// pipe) and passes with at most .y mismatches (-1: never, e.g. no base left or a round that starts in front of the read)
	int2 r1tail[NT][32];  // read-1 scan, round QF+t, lane
	int2 r2tail[32];      // read-2 scan, last round, lane
	int16_t r2lim[NW][32]; // read-2 scan, round, lane: mismatch limit of a full window, -1 where the round starts in the padding
};

__device__ __forceinline__ int max_mismatches(uint32_t pass_by_mm) { return 31 - __clz(pass_by_mm); } // pass sets are intervals 0..k (full_ok); 0 -> -1

template <int NW, int FULL>
__device__ __forceinline__ void full_tab_init(const KArgs& A, const SmemTables& T, FullTab<NW, FULL>& F, int tid, int nthreads)
{
	constexpr int QF = FullTab<NW, FULL>::QF;
	constexpr int D = 32 * NW - FULL;
	for (int o = tid; o < 32 * NW; o += nthreads)
	{
		const int tot = FULL - o;
		int t = -1;
		if (o >= 1 && tot > 0) t = max(tot - (int)T.mmin[tot], -1);
		F.thr[o] = (int16_t)t;
	}
	if (tid == 0) // running maximum over growing overlaps (o descending); a few hundred steps once per CTA
	{
		int best = -1;
		for (int o = 32 * NW - 1; o >= 0; --o)
		{
			const int tot = FULL - o;
			if (o >= 1 && tot > 0) best = max(best, max(tot - (int)T.mmin[tot], -1));
			F.thr_env[o] = (int16_t)(o >= 1 && tot > 0 ? best : -1);
		}
	}
	for (int i = tid; i < 32 * (NW - QF); i += nthreads)
	{
		const int q = QF + (i >> 5), l = i & 31;
		const int cnt = min(A.a_size, FULL - 32 * q - l);
		F.r1tail[i >> 5][l] = cnt > 0 ? make_int2((int)(1u << (32 - cnt)), max_mismatches(T.passM[cnt])) : make_int2(0, -1);
	}
	for (int i = tid; i < 32 * NW; i += nthreads)
	{
		const int q = i >> 5, l = i & 31;
		const bool inside = 32 * q + l >= D; // the round starts inside the read (bits below D are padding)
		if (q == NW - 1)
		{
			const int cnt = min(A.a_size, 32 - l);
			F.r2tail[l] = inside ? make_int2((int)(1u << (32 - cnt)), max_mismatches(T.passM[cnt])) : make_int2(0, -1);
		}
		F.r2lim[q][l] = (int16_t)(inside ? A.a2maxmm : -1);
	}
}

// planes of a read of length FULL whose position p sits at bit p+D (D = 0: left aligned, D = 32*NW-FULL: right aligned).
// Loads are unconditional with the word offset folded into the instruction; lanes outside the read may read a neighbouring row
// of the staged tile (always inside the stage) and are replaced by 'A' where a word is cut by the read's ends.
template <int NW, int FULL, int D>
__device__ __forceinline__ uint32_t pack_full(const SmemTables& T, uint32_t row, int lane, Planes<NW>& pl)
{
	uint32_t bad = 0;
	const uint32_t lut = smem_u32(T.not_acgt);
	const uint32_t base = row + lane - D;
	static_for<NW>([&](auto wc) {
		constexpr int w = decltype(wc)::value;
		constexpr int lo = 32 * w - D, hi = 32 * w + 31 - D; // positions held by lanes 0 and 31
		if constexpr (hi < 0 || lo >= FULL) pl.h[w] = pl.l[w] = 0;
		else
		{
			uint32_t c = lds_u8_at<32 * w>(base);
			if constexpr (lo < 0 || hi >= FULL) c = ((unsigned)(32 * w + lane - D) < (unsigned)FULL) ? c : (uint32_t)'A';
			pl.h[w] = ballot_bits(c, 4u);
			pl.l[w] = ballot_bits(c, 2u);
			bad |= lds_u8(lut + c);
		}
	});
	return bad;
}

// steps 1-3 for a pair of two FULL-length reads without N (same results as steps_planes<NW,false>)
// bad: this lane saw a byte other than A/C/G/T while packing; not_plain is set (and nothing else decided) if any lane did -- the test
// rides on the vote of the sweep, so that packing and sweep form one stretch of straight-line code.
template <int NW, int FULL>
__device__ __forceinline__ Step123 steps_full(const KArgs& A, const FullTab<NW, FULL>& F, const Pair& P, const Planes<NW>& f1, const Planes<NW>& f2r, uint32_t bad,
                                              int lane, bool& not_plain)
{
	not_plain = false;
	constexpr int QF = FullTab<NW, FULL>::QF;
	Step123 r;
	r.fwd = r.rev = -1;
	r.best_offset = -1;
	// ---- step 1 ----
	{
		// Pre-filter on the lo plane alone: two bases that differ in their lo bit are a mismatch, so the number of lo-plane
		// differences never exceeds the number of mismatches and an offset whose lo-plane count is already above the limit
		// (F.thr) is out. For unrelated reads half of the lo bits differ against a limit of <= 20 %: about one pair in twenty
		// keeps a false survivor, which the exact count below then removes.
		uint32_t l[NW], s2l[NW], mk[NW];
#pragma unroll
		for (int w = 0; w < NW; ++w) l[w] = __brev(f2r.l[NW - 1 - w]); // lo plane of revcomp(read 2): the reversed bit string
#pragma unroll
		for (int w = 0; w < NW; ++w)
		{
			s2l[w] = __funnelshift_r(l[w], (w + 1 < NW) ? l[w + 1] : 0u, lane); // shifted right by lane: positions 32*w+lane ..
			mk[w] = kFull;
			if (!(32 * w + 62 < FULL)) // the word holds positions >= FULL - lane for some lane: mask of compared positions (i < FULL - o)
			{
				const int n0 = FULL - 32 * w, n1 = FULL - 32 * (w + 1);
				const uint32_t c0 = n0 >= 32 ? kFull : (n0 <= 0 ? 0u : ((1u << (n0 & 31)) - 1u));
				const uint32_t c1 = n1 >= 32 ? kFull : (n1 <= 0 ? 0u : ((1u << (n1 & 31)) - 1u));
				mk[w] = __funnelshift_r(c0, c1, lane);
			}
		}
		const uint32_t thr_addr = smem_u32(F.thr) + 2u * (uint32_t)lane;
		int mmlq[NW];
		bool any = false;
		static_for<NW>([&](auto qc) {
			constexpr int q = decltype(qc)::value;
			int mml = 0;
#pragma unroll
			for (int k = 0; k < NW - q; ++k)
			{
				const int w = q + k;
				uint32_t x;
				if (32 * w + 62 < FULL) x = s2l[w] ^ f1.l[k];
				else x = xor_and(s2l[w], f1.l[k], mk[w]);
				mml += __popc(x);
			}
			mmlq[q] = mml;
			any |= mml <= lds_s16_at<64 * q>(thr_addr); // the limit of offset 32*q+lane
		});
		if (__any_sync(kFull, any || (bad & 1u))) // pairs with an insert match, false survivors of the pre-filter, pairs with N etc.
		{
			if (__any_sync(kFull, bad & 1u))
			{
				not_plain = true;
				return r;
			}
			// exact count for the rounds that hold a survivor. The complement of the hi plane of revcomp(read 2) is folded into
			// the comparison (xnor), so h is the plain reversed plane; zeros shifted in at the top read as mismatches, but only at
			// positions >= FULL, which mk removes.
			uint32_t h[NW], s2h[NW];
#pragma unroll
			for (int w = 0; w < NW; ++w) h[w] = __brev(f2r.h[NW - 1 - w]);
#pragma unroll
			for (int w = 0; w < NW; ++w) s2h[w] = __funnelshift_r(h[w], (w + 1 < NW) ? h[w + 1] : 0u, lane);
			uint32_t key = kNoKey;
#pragma unroll
			for (int q = 0; q < NW; ++q)
			{
				const int t = F.thr[32 * q + lane];
				if (__ballot_sync(kFull, mmlq[q] <= t) == 0) continue;
				int mm = 0;
#pragma unroll
				for (int k = 0; k < NW - q; ++k)
				{
					const int w = q + k;
					mm += __popc((~(s2h[w] ^ f1.h[k]) | (s2l[w] ^ f1.l[k])) & mk[w]);
				}
				uint32_t b = __ballot_sync(kFull, mm <= t); // implies the pre-filter: mm >= the lo-plane count
				while (b)
				{
					const int src = __ffs(b) - 1;
					b &= b - 1;
					const int mms = __shfl_sync(kFull, mm, src);
					const int tot = FULL - 32 * q - src;
					key = min(key, candidate_key_warp(A, P, 32 * q + src, tot - mms, mms, lane));
				}
			}
			if (key != kNoKey) r.best_offset = (int)(key & 0xFFFFu);
		}
	}
	if (r.best_offset >= 0) return r;
	// ---- step 2: read 1 against adapter 1 ----
	// lane l owns the offsets 32*q+l; the a_size-base window of a round is isolated by a multiplication (a left shift that drops
	// the bits above it, FMA pipe) and passes with at most a?maxmm mismatches; one vote per read, positions only for hits
	const uint32_t a1mul = 1u << (32 - A.a_size);
	constexpr int D = 32 * NW - FULL;
	int mm1[NW], mm2[NW];
	bool any1 = false, any2 = false;
	const int2 t2 = F.r2tail[lane];
	{
		uint32_t sh[NW], sl[NW];
		shift_words<NW>(f1.h, lane, sh);
		shift_words<NW>(f1.l, lane, sl);
		const uint32_t tail_addr = smem_u32(F.r1tail) + 8u * (uint32_t)lane;
		static_for<NW>([&](auto qc) {
			constexpr int q = decltype(qc)::value;
			const uint32_t x = (sh[q] ^ A.a1h) | (sl[q] ^ A.a1l);
			if constexpr (q < QF)
			{
				mm1[q] = __popc(x * a1mul);
				any1 |= mm1[q] <= A.a1maxmm;
			}
			else
			{
				const uint2 t = lds_v2_at<256 * (q - QF)>(tail_addr); // F.r1tail[q - QF][lane]
				mm1[q] = __popc(x * t.x);
				any1 |= mm1[q] <= (int)t.y;
			}
		});
	}
	// ---- step 3: read 2 (original orientation, right-aligned planes) against adapter 2 ----
	{
		uint32_t sh[NW], sl[NW];
		shift_words<NW>(f2r.h, lane, sh);
		shift_words<NW>(f2r.l, lane, sl);
		const uint32_t lim_addr = smem_u32(F.r2lim) + 2u * (uint32_t)lane;
		static_for<NW>([&](auto qc) {
			constexpr int q = decltype(qc)::value;
			if constexpr (32 * q + 31 < D) mm2[q] = 0x7fff; // the whole round lies in the padding in front of the read
			else
			{
				const uint32_t x = (sh[q] ^ A.a2h) | (sl[q] ^ A.a2l);
				if constexpr (q < NW - 1)
				{
					mm2[q] = __popc(x * a1mul);
					int lim = A.a2maxmm;
					if constexpr (32 * q < D) lim = lds_s16_at<64 * q>(lim_addr); // some lanes of this round start in the padding: F.r2lim[q][lane]
					any2 |= mm2[q] <= lim;
				}
				else
				{
					mm2[q] = __popc(x * (uint32_t)t2.x);
					any2 |= mm2[q] <= t2.y;
				}
			}
		});
	}
	if (__any_sync(kFull, any1 || any2)) // one vote for both reads; the positions are worked out only for hits
	{
		if (__any_sync(kFull, any1))
		{
			uint32_t pm = 0;
#pragma unroll
			for (int q = NW - 1; q >= 0; --q) pm = pm * 2u + (mm1[q] <= (q < QF ? A.a1maxmm : F.r1tail[q < QF ? 0 : q - QF][lane].y) ? 1u : 0u);
			const uint32_t mine = pm ? (uint32_t)(32 * (__ffs(pm) - 1) + lane) : 0xFFFFFFFFu;
			r.fwd = (int)__reduce_min_sync(kFull, mine);
		}
		if (__any_sync(kFull, any2))
		{
			uint32_t pm = 0;
#pragma unroll
			for (int q = NW - 1; q >= 0; --q) pm = pm * 2u + (mm2[q] <= (q < NW - 1 ? (int)F.r2lim[q][lane] : t2.y) ? 1u : 0u);
			const uint32_t mine = pm ? (uint32_t)(32 * (__ffs(pm) - 1) + lane - D) : 0xFFFFFFFFu;
			r.rev = (int)__reduce_min_sync(kFull, mine);
		}
	}
	return r;
}

// pairs in which a byte other than A/C/G/T was seen: N planes, or the byte-wise path for anything else.
// status: SPG_PAIR_OK or SPG_PAIR_BAD_BASE_R2.
template <int NW>
__device__ __noinline__ RareSteps steps_special(const KArgs& A, const SmemTables& T, Pair P, int lane)
{
	RareSteps r;
	r.st.best_offset = r.st.fwd = r.st.rev = -1;
	r.status = SPG_PAIR_OK;
	const int D2 = 32 * NW - P.len2;
	uint32_t n1[NW], n2r[NW];
	bool hasN1, hasN2, other1, other2;
	pack_special<NW>(P.r1, P.len1, 0, lane, n1, hasN1, other1);
	pack_special<NW>(P.r2, P.len2, D2, lane, n2r, hasN2, other2);
	{
		// hasN only gates trimN: a read with fewer than ncut N cannot hold a run of ncut of them
		int c1 = 0, c2 = 0;
#pragma unroll
		for (int w = 0; w < NW; ++w)
		{
			c1 += __popc(n1[w]);
			c2 += __popc(n2r[w]);
		}
		r.hasN = (hasN1 && c1 >= A.ncut ? 1 : 0) | (hasN2 && c2 >= A.ncut ? 2 : 0);
	}
	if (other2) r.status = SPG_PAIR_BAD_BASE_R2; // Sequence::complement throws (Sequence.cpp:46-71)
	else if (other1) r.st = steps_bytewise(A, T, P, lane); // read 1 bytes are compared as plain bytes by the reference
	else
	{
		Planes<NW> f1, f2r;
		pack_forward<NW>(T, P.r1, P.len1, 0, lane, f1);
		pack_forward<NW>(T, P.r2, P.len2, D2, lane, f2r);
		r.st = steps_planes<NW, true>(A, T, P, f1, f2r, D2, n1, n2r, lane);
	}
	return r;
}

// pairs that do not fit the plane path (long reads, forced byte-wise mode)
__device__ __noinline__ RareSteps steps_long(const KArgs& A, const SmemTables& T, Pair P, int lane)
{
	RareSteps r;
	r.st.best_offset = r.st.fwd = r.st.rev = -1;
	r.status = SPG_PAIR_OK;
	bool bad2 = false, n1 = false, n2 = false;
	for (int i = lane; i < P.len2 && i < A.stride; i += 32)
	{
		const uint32_t c = lds_u8(P.r2 + i);
		bad2 |= !is_acgtn(c);
		n2 |= (c == 'N');
	}
	for (int i = lane; i < P.len1 && i < A.stride; i += 32) n1 |= (lds_u8(P.r1 + i) == 'N');
	bad2 = __any_sync(kFull, bad2);
	r.hasN = (__any_sync(kFull, n1) ? 1 : 0) | (__any_sync(kFull, n2) ? 2 : 0);
	const int maxlen = max(P.len1, P.len2);
	if (bad2) r.status = SPG_PAIR_BAD_BASE_R2;
	else if (maxlen >= SPG_MAXLEN || maxlen > A.stride) r.status = SPG_PAIR_TOO_LONG; // AnalysisWorker.cpp:131-134
	else r.st = steps_bytewise(A, T, P, lane);
	return r;
}

// ---- one read pair, one warp ------------------------------------------------------------------------------------------------------------
template <int NW, int FULL>
__device__ __forceinline__ void process_pair(const KArgs& A, const SmemTables& T, const FullTab<(NW > 0 ? NW : 1), (FULL > 0 ? FULL : 1)>& F, const Pair& P, int lane,
                                             spg_result* out, bool& edited)
{
	constexpr int NWP = NW > 0 ? NW : 1;
	constexpr int FULLP = FULL > 0 ? FULL : 1;
	int status = SPG_PAIR_OK;
	bool hasN1 = false, hasN2 = false;
	Step123 st;
	bool rare = true;
	if (NW > 0 && FULL > 0 && P.len1 == FULL && P.len2 == FULL && !A.force_bytewise) // the host picks FULL variants only with A.full_ok
	{
		Planes<NWP> f1, f2r;
		const uint32_t bad = pack_full<NWP, FULLP, 0>(T, P.r1, lane, f1) | pack_full<NWP, FULLP, 32 * NWP - FULLP>(T, P.r2, lane, f2r);
		bool not_plain;
		st = steps_full<NWP, FULLP>(A, F, P, f1, f2r, bad, lane, not_plain);
		rare = not_plain; // a byte other than A/C/G/T somewhere
	}
	else if (NW > 0 && max(P.len1, P.len2) <= 32 * NW && !A.force_bytewise)
	{
		Planes<NWP> f1, f2r;
		const int D2 = 32 * NWP - P.len2; // read 2 is packed right aligned
		const uint32_t bad = pack_forward<NWP>(T, P.r1, P.len1, 0, lane, f1) | pack_forward<NWP>(T, P.r2, P.len2, D2, lane, f2r);
		if (ballot_bits(bad, 1u) == 0) // the common case: only A/C/G/T in both reads
		{
			uint32_t none[NWP];
			st = steps_planes<NWP, false>(A, T, P, f1, f2r, D2, none, none, lane);
			rare = false;
		}
	}
	if (rare)
	{
		const RareSteps r = (NW > 0 && max(P.len1, P.len2) <= 32 * NW && !A.force_bytewise) ? steps_special<NWP>(A, T, P, lane) : steps_long(A, T, P, lane);
		st = r.st;
		status = r.status;
		hasN1 = r.hasN & 1;
		hasN2 = r.hasN & 2;
	}

	int n1 = P.len1, n2 = P.len2;
	uint32_t flags = 0;
	if (status == SPG_PAIR_OK)
	{
		if (st.best_offset >= 0) // insert hit (AnalysisWorker.cpp:269-302)
		{
			const int new_length = P.len2 - st.best_offset;
			n1 = min(n1, new_length);
			n2 = min(n2, new_length);
			flags |= SPG_F_INSERT;
			if (A.ec)
			{
				bool nn1 = false, nn2 = false;
				if (!correct_errors_warp(A, P, n1, n2, lane, nn1, nn2)) status = SPG_PAIR_BAD_BASE_EC;
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
	}
	if (status == SPG_PAIR_OK)
	{
		if (A.qcut > 0) // :430-434
		{
			int t1, t2;
			trim_quality_pair(A, P, n1, n2, lane, t1, t2);
			if (t1 < n1) flags |= SPG_F_Q1;
			if (t2 < n2) flags |= SPG_F_Q2;
			n1 = t1;
			n2 = t2;
		}
		if (A.ncut > 0 && (hasN1 || hasN2)) // :437-441 (a read without any N cannot be cut)
		{
			if (hasN1)
			{
				const int t1 = trim_n_warp(P.r1, n1, A.ncut, lane);
				if (t1 < n1) flags |= SPG_F_N1;
				n1 = t1;
			}
			if (hasN2)
			{
				const int t2 = trim_n_warp(P.r2, n2, A.ncut, lane);
				if (t2 < n2) flags |= SPG_F_N2;
				n2 = t2;
			}
		}
	}
	{
		// one 8-byte store by lane 0: len1 | len2<<16 , best_offset | flags<<16 | status<<24 (the values are warp-uniform)
		uint2 rec;
		if (status == SPG_PAIR_OK)
		{
			rec.x = (uint32_t)n1 | ((uint32_t)n2 << 16);
			rec.y = ((uint32_t)st.best_offset & 0xFFFFu) | (flags << 16);
		}
		else
		{
			rec.x = 0;
			rec.y = 0xFFFFu | ((uint32_t)status << 24);
		}
		stg_v2_lane0(out, rec.x, rec.y, lane);
	}
}

// ---- the kernel: persistent CTAs, one producer warp + CW consumer warps, NS-stage TMA ring ------------------------------------------------------
// dynamic shared memory: [stages][ b1 | q1 | b2 | q2 : tile_pairs*stride each ][ len1 | len2 : tile_pairs u16 each ]
// The pairs of a tile are dealt to the consumer warps round robin (pair index = warp, warp + CW, ...).
// FULL > 0: additionally compiled for pairs of two reads of exactly FULL bases (the fast path above); 0: general code only.
template <int NW, int CW, int MINB, int FULL = 0>
__global__ void __launch_bounds__((CW + 1) * 32, MINB) trim_kernel(const __grid_constant__ KArgs A)
{
	static_assert(FULL == 0 || (NW > 0 && FULL <= 32 * NW && FULL >= 52), "FULL must fit the plane words");
	constexpr int kThreads = (CW + 1) * 32;
	extern __shared__ __align__(128) uint8_t smem[];
	__shared__ __align__(8) uint64_t full_bar[kMaxStages];
	__shared__ __align__(8) uint64_t empty_bar[kMaxStages];
	__shared__ SmemTables T;
	__shared__ FullTab<(NW > 0 ? NW : 1), (FULL > 0 ? FULL : 1)> F;

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int TP = A.tile_pairs;
	const uint32_t plane_bytes = (uint32_t)TP * (uint32_t)A.stride;
	const uint32_t stage_bytes = 4u * plane_bytes + 4u * (uint32_t)TP;
	// pair and tile indices fit 32 bits: a launch holds fewer than 2^31 pairs (checked by the host)
	const uint32_t n_pairs = A.n_dev ? (uint32_t)*A.n_dev : (uint32_t)A.n_pairs;
	const uint32_t n_tiles = (n_pairs + (uint32_t)TP - 1u) / (uint32_t)TP;
	const uint32_t smem_base = smem_u32(smem);

	for (int i = threadIdx.x; i < SPG_MAXLEN; i += kThreads) T.mmin[i] = A.mmin[i];
	for (int i = threadIdx.x; i < 256; i += kThreads) T.not_acgt[i] = (i == 'A' || i == 'C' || i == 'G' || i == 'T') ? 0 : 1;
	if (threadIdx.x < 21)
	{
		T.passA[threadIdx.x] = A.passA[threadIdx.x];
		uint32_t bm = 0; // re-index by mismatches: bit j <- bit (T - j)
		for (int j = 0; j <= (int)threadIdx.x; ++j) bm |= ((A.passA[threadIdx.x] >> (threadIdx.x - j)) & 1u) << j;
		T.passM[threadIdx.x] = bm;
	}
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
	if (NW > 0 && FULL > 0)
	{
		full_tab_init(A, T, F, (int)threadIdx.x, kThreads);
		__syncthreads();
	}

	if (warp == CW)
	{
		// ===== producer: one lane issues the bulk copies of each tile =====
		if (lane == 0)
		{
			int it = 0;
			for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it)
			{
				const int s = it % A.stages;
				const uint32_t round = (uint32_t)(it / A.stages);
				if (round > 0)
				{
					mbar_wait(&empty_bar[s], (round - 1) & 1u);
				}
				const uint32_t first = t * (uint32_t)TP;
				const int cnt = (int)min((uint32_t)TP, n_pairs - first);
				// bulk copies move multiples of 16 bytes: a ragged last tile (cnt not a multiple of 8) reads up to 14 bytes past its
				// last row, which stay inside the row planes (they are allocated in multiples of 8 rows)
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
		int it = 0;
		for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it)
		{
			const int s = it % A.stages;
			const uint32_t round = (uint32_t)(it / A.stages);
			mbar_wait(&full_bar[s], round & 1u);
			const uint32_t first = t * (uint32_t)TP;
			const int cnt = (int)min((uint32_t)TP, n_pairs - first);
			const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
			const uint32_t lens = st + 4 * plane_bytes;
			// pairs of a tile are dealt to the consumer warps round robin: no claim, the row addresses advance by additions. The ring
			// buffers one tile of imbalance between the warps (a warp that is done moves on to the next stage on its own).
			// (only the pair index is carried through the loop; everything else is derived from it and from per-tile values)
			for (int pr = warp; pr < cnt; pr += CW)
			{
				Pair P;
				P.r1 = st + (uint32_t)pr * (uint32_t)A.stride;
				P.q1 = P.r1 + plane_bytes;
				P.r2 = P.r1 + 2 * plane_bytes;
				P.q2 = P.r1 + 3 * plane_bytes;
				P.len1 = (int)lds_u16(lens + 2u * (uint32_t)pr);
				P.len2 = (int)lds_u16(lens + 2u * (uint32_t)(TP + pr));
				spg_result* const outp = A.out + (first + (uint32_t)pr);
				bool edited = false;
				process_pair<NW, FULL>(A, T, F, P, lane, outp, edited);
				if (edited) // -ec: write the edited rows back
				{
					__syncwarp();
					const size_t goff = (size_t)(first + pr) * A.stride;
					const int halves = A.stride / 2; // rows are 2-byte aligned (stride is even)
					for (int v = lane; v < halves; v += 32)
					{
						reinterpret_cast<uint16_t*>(A.b1 + goff)[v] = (uint16_t)lds_u16(P.r1 + 2u * v);
						reinterpret_cast<uint16_t*>(A.q1 + goff)[v] = (uint16_t)lds_u16(P.q1 + 2u * v);
						reinterpret_cast<uint16_t*>(A.b2 + goff)[v] = (uint16_t)lds_u16(P.r2 + 2u * v);
						reinterpret_cast<uint16_t*>(A.q2 + goff)[v] = (uint16_t)lds_u16(P.q2 + 2u * v);
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
