// spg_engine.cu -- host side of libseqpurge_b200.so: the C ABI of include/seqpurge_b200.h.
//
// Builds the integer decision tables on the host (with the host's libm, evaluating the same expressions as
// BasicStatistics::matchProbability / AnalysisWorker::run of the reference, src/cppCORE/BasicStatistics.cpp:281-307,
// src/SeqPurge/AnalysisWorker.cpp:170,178-179,246-259,346-348), owns the pinned slots and the per-device buffers and
// streams, and launches spg::trim_kernel. No CPU implementation of the trimming itself exists in this library.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <string>
#include <vector>
#include <algorithm>

#include "spg_kernel.cuh"
#include "spg_lanes.cuh"
#include "spg_qc.cuh"
#include "spg_qc_lanes.cuh"
#include "spg_fastq.cuh"

namespace
{

std::string g_create_error;

// ---- decision tables ---------------------------------------------------------------------------------------------------------------------
struct HostTables
{
	std::vector<uint16_t> mmin;    // [1000]
	std::vector<uint16_t> ranktab; // [171*171]
	std::vector<double> psmall;    // [(ao+1)^2]
	uint32_t passA[21];
	int qthr;
};

struct Factorials
{
	std::vector<double> f; // k! while it fits a double: 0..170
	Factorials()
	{
		double v = 1.0;
		int i = 0;
		while (std::isfinite(v))
		{
			f.push_back(v);
			++i;
			v *= i;
		}
	}
};
const Factorials& factorials()
{
	static Factorials F;
	return F;
}

// P(X >= n), X ~ Binomial(count, p), evaluated term by term in increasing i exactly like the reference does
double match_probability(double p, int n, int count)
{
	const std::vector<double>& f = factorials().f;
	int mismatches = count - n;
	while (count >= (int)f.size())
	{
		n /= 2;
		mismatches /= 2;
		count = n + mismatches;
	}
	double output = 0.0;
	for (int i = n; i <= count; ++i)
	{
		double q = std::pow(1.0 - p, (double)(count - i)) * std::pow(p, (double)i) * f[count] / f[i] / f[count - i];
		output += q;
	}
	return output;
}

bool build_tables(const spg_params& p, int a_size, HostTables& t, std::string& err)
{
	if ((int)factorials().f.size() != spg::kRankDim)
	{
		err = "unexpected factorial cache size";
		return false;
	}
	// minimum matches per number of compared bases T: smallest m with !(100.0*m/T < match_perc)
	t.mmin.assign(SPG_MAXLEN, 0xFFFF);
	for (int T = 1; T < SPG_MAXLEN; ++T)
	{
		for (int m = 0; m <= T; ++m)
		{
			if (!(100.0 * m / T < p.match_perc))
			{
				t.mmin[T] = (uint16_t)m;
				break;
			}
		}
	}
	// dense ranks of the probabilities that pass -mep
	std::vector<double> ps;
	std::vector<double> cell((size_t)spg::kRankDim * spg::kRankDim, std::numeric_limits<double>::quiet_NaN());
	for (int count = 0; count < spg::kRankDim; ++count)
	{
		for (int n = 0; n <= count; ++n)
		{
			double v = match_probability(0.25, n, count);
			if (!std::isfinite(v))
			{
				err = "match probability is not finite";
				return false;
			}
			cell[(size_t)count * spg::kRankDim + n] = v;
			// an offset can only become the best one with p <= mep AND p < best_p, which starts at 1.0 (AnalysisWorker.cpp:138,261):
			// a cell with p == 1.0 never wins, whatever -mep says
			if (!(v > p.mep) && v < 1.0) ps.push_back(v);
		}
	}
	std::sort(ps.begin(), ps.end());
	ps.erase(std::unique(ps.begin(), ps.end()), ps.end());
	if (ps.size() >= 0xFFFF)
	{
		err = "too many distinct probabilities";
		return false;
	}
	t.ranktab.assign((size_t)spg::kRankDim * spg::kRankDim, 0xFFFF);
	for (int count = 0; count < spg::kRankDim; ++count)
	{
		for (int n = 0; n <= count; ++n)
		{
			double v = cell[(size_t)count * spg::kRankDim + n];
			if (v > p.mep || !(v < 1.0)) continue;
			t.ranktab[(size_t)count * spg::kRankDim + n] = (uint16_t)(std::lower_bound(ps.begin(), ps.end(), v) - ps.begin());
		}
	}
	// Fold the -mep test into the pre-filter where no halving is involved (T <= 170): an offset with T compared bases needs
	// at least mmin[T] matches to pass -match_perc AND to have p <= mep. Only done when the pass set is an upper interval in
	// m (it always is for a binomial tail; verified here rather than assumed), so the filter stays exact.
	for (int T = 1; T < spg::kRankDim; ++T)
	{
		int first = T + 1;
		for (int m = 0; m <= T; ++m)
			if (t.ranktab[(size_t)T * spg::kRankDim + m] != 0xFFFF)
			{
				first = m;
				break;
			}
		bool interval = true;
		for (int m = first; m <= T; ++m)
			if (t.ranktab[(size_t)T * spg::kRankDim + m] == 0xFFFF) interval = false;
		if (!interval) continue;
		if (first > T) t.mmin[T] = 0xFFFF;
		else if (t.mmin[T] != 0xFFFF && first > t.mmin[T]) t.mmin[T] = (uint16_t)first;
	}
	// probabilities of the short adapter fragments of the presence check
	const int ao = p.adapter_overlap;
	t.psmall.assign((size_t)(ao + 1) * (ao + 1), 1.0);
	for (int count = 0; count <= ao; ++count)
		for (int n = 0; n <= count; ++n) t.psmall[(size_t)count * (ao + 1) + n] = match_probability(0.25, n, count);
	// adapter-only scans: pass bit per (compared bases T, matches m); T=0 gives 0/0 = NaN, which is not "< match_perc"
	for (int T = 0; T <= 20; ++T)
	{
		t.passA[T] = 0;
		for (int m = 0; m <= T && T <= a_size; ++m)
		{
			int mm = T - m;
			if (100.0 * m / (m + mm) < p.match_perc) continue;
			if (match_probability(0.25, m, m + mm) > p.mep) continue;
			t.passA[T] |= 1u << m;
		}
	}
	// quality window: smallest integer sum s with (double)s/window >= cutoff (FastqFileStream.cpp:69)
	{
		long long s = (long long)p.qcut * p.qwin - 8;
		while (!((double)s / p.qwin >= p.qcut)) ++s;
		t.qthr = (int)s;
	}
	return true;
}

void adapter_planes(const char* a, int a_size, uint32_t& h, uint32_t& l, uint32_t& n)
{
	h = l = n = 0;
	for (int i = 0; i < a_size; ++i)
	{
		unsigned c = (unsigned char)a[i];
		if (c & 4u) h |= 1u << i;
		if (c & 2u) l |= 1u << i;
		if (c == 'N') n |= 1u << i;
	}
}

bool all_acgtn(const char* a, int n)
{
	for (int i = 0; i < n; ++i)
	{
		char c = a[i];
		if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'N') return false;
	}
	return true;
}

// ---- context ---------------------------------------------------------------------------------------------------------------------------------
struct Device
{
	int id = -1;
	int sm_count = 0;
	cudaStream_t stream = nullptr;
	uint16_t* d_mmin = nullptr;
	uint16_t* d_rank = nullptr;
	double* d_psmall = nullptr;
	unsigned long long* d_ec = nullptr; // 3 * SPG_MAXLEN
	unsigned long long* d_qc = nullptr; // spg::kQcWords accumulators of the -qc statistics
	// per kernel instantiation: the dynamic shared memory the function attribute was raised to, and the resident CTAs per SM the
	// occupancy calculator gave for the last (threads, smem) it was asked about. Guarded by spg_ctx::mu.
	struct Occ
	{
		size_t smem_attr = 0; // cudaFuncAttributeMaxDynamicSharedMemorySize set so far
		size_t smem = 0;      // launch geometry the cached value belongs to (0: none yet)
		int ctas = 0;
	};
	Occ occ[6][3];      // NW = 0,5,8,10,16,32 x kernel variant (min blocks 2,3,4; the two long-read widths have one variant each)
	Occ full_occ[4][9]; // the variants compiled for one read length (full_index)
	Occ lane_occ[4][9]; // the lane-per-pair kernels of the same read lengths
	Occ qc_occ[3];      // qc_kernel, NW = 5,8,10
	Occ qc_lane_occ[3]; // qc_lanes_kernel, NW = 5,8,10
};

enum SlotState
{
	SLOT_IDLE = 0,
	SLOT_SUBMITTED = 1
};

struct Slot
{
	int dev = 0;
	uint8_t* h_block = nullptr; // pinned: b1|q1|b2|q2|len1|len2|qtail1|qtail2
	uint8_t* h_block_dev = nullptr; // the same memory as the slot's device sees it (mapped pinned memory)
	bool zero_copy = false;     // last submit left the quality planes in the slot (see spg_submit)
	uint8_t* d_block = nullptr;
	spg_result* h_res = nullptr;
	spg_result* d_res = nullptr;
	cudaEvent_t done = nullptr;
	cudaStream_t stream = nullptr; // own stream: H2D / kernel / D2H of different slots of one device overlap
	int state = SLOT_IDLE;
	int n = 0;
};

} // namespace

namespace
{
void fq_free(spg_fq* fq); // spg_fastq_engine.inc
}

struct spg_ctx
{
	spg_params params;
	std::string a1, a2;
	int a_size = 0;
	bool adapters_plain = true; // adapters consist of ACGTN only
	HostTables tables;
	std::vector<Device> devs;
	std::vector<Slot> slots;
	int max_pairs = 0, max_len = 0, stride = 0;
	size_t plane_bytes = 0, len_bytes = 0, tail_bytes = 0, block_bytes = 0; // slot block: b1|q1|b2|q2|len1|len2|qtail1|qtail2
	std::string err;
	std::mutex mu;
	int force_bytewise = 0;
	int ctas_per_sm = 0; // 0 = occupancy
	int min_blocks = 3; // __launch_bounds__ min CTAs/SM of the kernel variant (register budget)
	int full_len = -1;  // SPG_OPT_FULL_LEN: -1 = automatic (max_len of the context / longest line of the previous chunk), 0 = general kernel
	int tile_pairs = 0; // 0 = automatic
	int stages = 0;     // 0 = automatic
	long long launches = 0;
	int n_lanes = 1;                   // SPG_OPT_N_LANES: pairs with N go through the lane kernel's N-aware path (0: warp-cooperative general path)
	int qual_tails = 0;                // SPG_OPT_QUAL_TAILS: the caller fills the slots' qtail1 / qtail2 (spg_slot_qtails)
	int zero_copy_quals = 1;           // SPG_OPT_ZERO_COPY_QUALS: slots leave the quality planes in pinned host memory when the lane kernel runs
	int seed_scan = 1;                 // SPG_OPT_SEED_SCAN: 0 = the lane kernel evaluates every offset of the adapter scans (no pigeonhole filter)
	int kernel_layout = 0;             // SPG_OPT_KERNEL: 0 automatic, 1 warp per pair only, 2 lane per pair where it applies
	std::atomic<int> last_kernel{0};   // spg_last_kernel: layout * 100000 + NW * 1000 + FULL of the last trimming launch
	std::vector<spg_fq*> fqs; // FASTQ streams attached to this context (closed by spg_destroy if the caller did not)
};

namespace
{

int fail(spg_ctx* ctx, int code, const std::string& msg)
{
	if (ctx)
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		ctx->err = msg;
	}
	else g_create_error = msg;
	return code;
}
#define SPG_CUDA(ctx, call)                                                                                           \
	do                                                                                                                \
	{                                                                                                                 \
		cudaError_t e_ = (call);                                                                                      \
		if (e_ != cudaSuccess) return fail(ctx, SPG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
	} while (0)

// plane words per read of the trimming kernels: rows of up to 160 / 256 / 320 bytes keep a read in 5 / 8 / 10 registers per plane;
// longer rows (up to MAXLEN-1 = 999 bases) use 16 or 32 words (more registers, fewer resident warps, still bit planes instead of bytes)
int nw_for_stride(int stride) { return stride <= 160 ? 5 : stride <= 256 ? 8 : stride <= 320 ? 10 : stride <= 512 ? 16 : stride <= 1024 ? 32 : 0; }
int nw_index(int nw) { return nw == 5 ? 1 : nw == 8 ? 2 : nw == 10 ? 3 : nw == 16 ? 4 : nw == 32 ? 5 : 0; }
int nw_for_qc(int stride) { return stride <= 320 ? nw_for_stride(stride) : 0; } // the statistics kernel keeps its generic form beyond 320

void tile_geometry(int stride, int& tile_pairs, int& stages, size_t& smem)
{
	// sweeps (profiles/README.md): 2 stages of about 29 KB keep 3 CTAs per SM resident; with the pairs of a tile dealt round robin to the
	// 8 consumer warps 48 pairs (6 per warp) beat 40 and 32 by 2 %, smaller tiles lose to the barrier traffic
	int tp = (int)(29184 / (4 * (size_t)stride + 4)) / 8 * 8;
	if (tp < 8) tp = 8;
	if (tp > 48) tp = 48;
	tile_pairs = tp;
	stages = 2;

	smem = (size_t)stages * (4 * (size_t)tp * stride + 4 * (size_t)tp);
}

#ifndef SPG_CW
#define SPG_CW 8
#endif
#ifndef SPG_FULL_MINB
#define SPG_FULL_MINB 3 // resident CTAs per SM the read-length variants are compiled for
#endif
constexpr int kCW = SPG_CW; // consumer warps per CTA (+1 producer warp); geometry sweeps showed 4/6/8 within 3%

// Resident CTAs per SM of one kernel instantiation for a launch with `smem` bytes of dynamic shared memory. The function attribute
// is raised whenever a launch needs more than it was set to (one context launches the same instantiation with different row strides,
// hence different ring sizes: a later, larger request must not meet the first launch's limit), and the occupancy is asked again
// whenever the geometry differs from the cached one. `mu` serialises the cache (slots may be submitted from different threads).
template <typename K>
cudaError_t resident_ctas(K kernel, Device::Occ& c, int threads, size_t smem, std::mutex& mu, int& out)
{
	std::lock_guard<std::mutex> g(mu);
	if (smem > c.smem_attr)
	{
		cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess) return e;
		c.smem_attr = smem;
	}
	if (c.ctas == 0 || c.smem != smem)
	{
		int n = 0;
		cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem);
		if (e != cudaSuccess) return e;
		c.ctas = n < 1 ? 1 : n;
		c.smem = smem;
	}
	out = c.ctas;
	return cudaSuccess;
}

template <int NW, int MINB, int FULL = 0>
cudaError_t launch_cfg(const spg::KArgs& a, int sm_count, int ctas_per_sm, long long n_tiles, size_t smem, cudaStream_t stream, Device::Occ* occ_cache, std::mutex& mu)
{
	int occ = 1;
	cudaError_t e = resident_ctas(spg::trim_kernel<NW, kCW, MINB, FULL>, *occ_cache, (kCW + 1) * 32, smem, mu, occ);
	if (e != cudaSuccess) return e;
	const int per_sm = ctas_per_sm > 0 ? std::min(ctas_per_sm, occ) : occ;
	const int grid = (int)std::min<long long>(n_tiles, (long long)sm_count * per_sm);
	spg::trim_kernel<NW, kCW, MINB, FULL><<<grid, (kCW + 1) * 32, smem, stream>>>(a);
	return cudaGetLastError();
}

template <int NW>
cudaError_t launch_nw(const spg::KArgs& a, int minb, int sm_count, int ctas_per_sm, long long n_tiles, size_t smem, cudaStream_t stream, Device::Occ* occ_cache, std::mutex& mu)
{
	switch (minb)
	{
		case 2: return launch_cfg<NW, 2>(a, sm_count, ctas_per_sm, n_tiles, smem, stream, occ_cache, mu);
		case 4: return launch_cfg<NW, 4>(a, sm_count, ctas_per_sm, n_tiles, smem, stream, occ_cache, mu);
		default: return launch_cfg<NW, 3>(a, sm_count, ctas_per_sm, n_tiles, smem, stream, occ_cache, mu);
	}
}

// The lane-per-pair kernel (spg_lanes.cuh) of one read length: 8 consumer warps + 1 producer warp, compiled for 2 resident CTAs per
// SM. The ring is as deep as two resident CTAs allow (2..4 stages of 32 pairs' base rows).
#ifndef SPG_LANE_CW
#define SPG_LANE_CW 24
#endif
#ifndef SPG_LANE_MINB
#define SPG_LANE_MINB 1
#endif
#ifndef SPG_LANE_CW_LONG
#define SPG_LANE_CW_LONG 24 // reads of more than 160 bases (8 or 10 plane words per read: more registers per lane)
#endif
#ifndef SPG_LANE_MINB_LONG
#define SPG_LANE_MINB_LONG 1
#endif
template <int NW>
struct LaneCfg
{
	static constexpr int CW = NW <= 5 ? SPG_LANE_CW : SPG_LANE_CW_LONG;
	static constexpr int MINB = NW <= 5 ? SPG_LANE_MINB : SPG_LANE_MINB_LONG;
};
template <int NW, int FULL>
cudaError_t launch_lanes(spg::KArgs a, int sm_count, int ctas_per_sm, int stages_opt, cudaStream_t stream, Device::Occ* occ_cache, std::mutex& mu)
{
	constexpr int CW = LaneCfg<NW>::CW, MINB = LaneCfg<NW>::MINB;
	const size_t stage = spg::lane_stage_bytes(a.stride);
	const size_t warps = (size_t)CW * spg::LaneSmem<NW>::kWarpBytes;
	int stages = spg::kLaneStagesMax;
	while (stages > 2 && MINB * (stages * stage + warps + 6 * 1024) > 227 * 1024) --stages;
	if (stages_opt >= 2 && stages_opt <= spg::kLaneStagesMax) stages = stages_opt;
	a.stages = stages;
	a.tile_pairs = 32;
	const size_t smem = stages * stage + warps;
	int occ = 1;
	cudaError_t e = resident_ctas(spg::trim_lanes_kernel<NW, FULL, CW, MINB>, *occ_cache, (CW + 1) * 32, smem, mu, occ);
	if (e != cudaSuccess) return e;
	const int per_sm = ctas_per_sm > 0 ? std::min(ctas_per_sm, occ) : occ;
	const long long n_tiles = (a.n_pairs + 31) / 32;
	const int grid = (int)std::min<long long>(n_tiles, (long long)sm_count * per_sm);
	spg::trim_lanes_kernel<NW, FULL, CW, MINB><<<grid, (CW + 1) * 32, smem, stream>>>(a);
	return cudaGetLastError();
}

// Kernel variants compiled for one read length (fast path for pairs of two full-length reads, spg_kernel.cuh): the usual
// lengths of Illumina runs. Any other length runs the general kernel; results do not depend on the choice.
int full_index(int nw, int full_len)
{
	static const int lens5[] = {150, 151, 100, 101, 125, 126, 75, 76};
	static const int lens8[] = {250, 251, 200, 201};
	static const int lens10[] = {300, 301};
	const int* tab = nw == 5 ? lens5 : nw == 8 ? lens8 : nw == 10 ? lens10 : nullptr;
	const int n = nw == 5 ? 8 : nw == 8 ? 4 : nw == 10 ? 2 : 0;
	for (int i = 0; i < n; ++i)
		if (tab[i] == full_len) return i + 1;
	return 0;
}

// n_dev: optional device pointer to the actual pair count (<= n); n then sizes the grid only.
// full_hint: read length most pairs of the batch are expected to have (0 = unknown); selects the kernel variant only.
int launch_trim(spg_ctx* ctx, Device& d, uint8_t* b1, uint8_t* q1, uint8_t* b2, uint8_t* q2, const uint16_t* len1, const uint16_t* len2, int stride, long long n,
                spg_result* out, cudaStream_t stream, const int* n_dev = nullptr, int full_hint = -1, bool* lanes_query = nullptr, bool quals_on_host = false,
                const uint8_t* qt1 = nullptr, const uint8_t* qt2 = nullptr)
{
	// lanes_query: only answer whether this launch would run the lane-per-pair kernel (nothing is launched)
	if (n <= 0 && !lanes_query) return SPG_OK;
	if (n > 0x7fffffffLL) return fail(ctx, SPG_ERR_PARAM, "at most 2^31-1 pairs per launch");
	if (full_hint < 0) full_hint = ctx->max_len;
	if (ctx->full_len >= 0) full_hint = ctx->full_len; // SPG_OPT_FULL_LEN
	spg::KArgs a;
	memset(&a, 0, sizeof(a));
	a.n_dev = n_dev;
	a.quals_on_host = quals_on_host ? 1 : 0;
	a.qt1 = qt1;
	a.qt2 = qt2;
	a.n_lanes = ctx->n_lanes;
	a.b1 = b1;
	a.q1 = q1;
	a.b2 = b2;
	a.q2 = q2;
	a.len1 = len1;
	a.len2 = len2;
	a.out = out;
	a.n_pairs = n;
	a.stride = stride;
	size_t smem;
	tile_geometry(stride, a.tile_pairs, a.stages, smem);
	if (ctx->tile_pairs > 0) a.tile_pairs = ctx->tile_pairs;
	if (ctx->stages > 0) a.stages = std::min(ctx->stages, spg::kMaxStages);
	smem = (size_t)a.stages * (4 * (size_t)a.tile_pairs * stride + 4 * (size_t)a.tile_pairs);
	if (smem > 200 * 1024) return fail(ctx, SPG_ERR_PARAM, "tile geometry exceeds shared memory");
	a.mmin = d.d_mmin;
	a.ranktab = d.d_rank;
	a.psmall = d.d_psmall;
	a.ec_m1 = d.d_ec;
	a.ec_m2 = d.d_ec + SPG_MAXLEN;
	a.ec_epr = d.d_ec + 2 * SPG_MAXLEN;
	const spg_params& p = ctx->params;
	a.mep = p.mep;
	a.a_size = ctx->a_size;
	a.ao = p.adapter_overlap;
	a.qcut = p.qcut;
	a.qwin = p.qwin;
	a.qoff = p.qoff;
	a.qthr = ctx->tables.qthr;
	a.ncut = p.ncut;
	a.ec = p.ec;
	a.force_bytewise = (ctx->force_bytewise || !ctx->adapters_plain) ? 1 : 0;
	adapter_planes(ctx->a1.data(), ctx->a_size, a.a1h, a.a1l, a.a1n);
	adapter_planes(ctx->a2.data(), ctx->a_size, a.a2h, a.a2l, a.a2n);
	memcpy(a.passA, ctx->tables.passA, sizeof(a.passA));
	{
		const uint32_t full = (ctx->a_size >= 32) ? 0xffffffffu : ((1u << ctx->a_size) - 1u);
		a.a1mask = full & ~a.a1n;
		a.a2mask = full & ~a.a2n;
		auto by_mm = [&](uint32_t mask) {
			const int tot = __builtin_popcount(mask);
			uint32_t bits = 0;
			for (int mm = 0; mm <= tot; ++mm)
				if ((ctx->tables.passA[tot] >> (tot - mm)) & 1u) bits |= 1u << mm;
			return bits;
		};
		a.a1pass = by_mm(a.a1mask);
		a.a2pass = by_mm(a.a2mask);
		// the read-length variants compare mismatch counts with a limit instead of looking a pass bit up: needs adapters without N in
		// their first a_size bases and pass sets that are intervals 0..k of the mismatch count (they are, for any sane parameters)
		auto interval = [](uint32_t v) { return (v & (v + 1u)) == 0u; };
		bool ok = ((a.a1n | a.a2n) & full) == 0u;
		for (int tot = 0; tot <= ctx->a_size && ok; ++tot)
		{
			uint32_t bits = 0;
			for (int mm = 0; mm <= tot; ++mm)
				if ((ctx->tables.passA[tot] >> (tot - mm)) & 1u) bits |= 1u << mm;
			ok = interval(bits);
		}
		a.full_ok = ok ? 1 : 0;
		a.a1maxmm = a.a1pass ? 31 - __builtin_clz(a.a1pass) : -1;
		a.a2maxmm = a.a2pass ? 31 - __builtin_clz(a.a2pass) : -1;
		// seed filter of the lane kernel's adapter scans (pigeonhole): a window of tot compared bases that passes must have fewer
		// mismatches than it holds complete 4-base blocks of the adapter's first 4 * (a_size / 4) bases
		bool seed = ok && ctx->adapters_plain;
		for (int tot = 1; tot <= ctx->a_size && seed; ++tot)
		{
			int kmax = -1;
			for (int mm = 0; mm <= tot; ++mm)
				if ((ctx->tables.passA[tot] >> (tot - mm)) & 1u) kmax = mm;
			if (kmax >= std::min(tot / 4, ctx->a_size / 4)) seed = false;
		}
		a.seed_ok = (seed && ctx->seed_scan) ? 1 : 0;
		{
			// a full window (a_size bases) with one N compares a_size-1 bases; the N spoils at most one block
			int kmax = -1;
			const int tot = ctx->a_size - 1;
			for (int mm = 0; mm <= tot; ++mm)
				if ((ctx->tables.passA[tot] >> (tot - mm)) & 1u) kmax = mm;
			a.seed_n1_ok = (a.seed_ok && kmax < ctx->a_size / 4 - 1) ? 1 : 0;
		}
		const int nwi = nw_for_stride(stride);
		auto code = [](char c) { return c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 0; };
		for (int i = 0; i < 20; ++i)
		{
			a.a1off[i] = (uint16_t)(i < ctx->a_size ? code(ctx->a1[(size_t)i]) * (nwi + 1) * 128 : 0);
			a.a2off[i] = (uint16_t)(i < ctx->a_size ? code(ctx->a2[(size_t)i]) * (nwi + 1) * 128 : 0);
		}
	}
	memset(a.a1, 'N', sizeof(a.a1)); // never read beyond the adapter length: a_size, adapter_overlap <= min(|a1|,|a2|,32)
	memset(a.a2, 'N', sizeof(a.a2));
	memcpy(a.a1, ctx->a1.data(), std::min<size_t>(32, ctx->a1.size()));
	memcpy(a.a2, ctx->a2.data(), std::min<size_t>(32, ctx->a2.size()));

	const int nw = nw_for_stride(stride);
	const int cw = ctx->min_blocks;
	Device::Occ* occ = &d.occ[nw_index(nw)][cw == 2 ? 0 : cw == 4 ? 2 : 1];
	const long long n_tiles = (n + a.tile_pairs - 1) / a.tile_pairs;
	cudaError_t e;
	const int fi = (cw == SPG_FULL_MINB && full_hint <= stride && !a.force_bytewise && a.full_ok) ? full_index(nw, full_hint) : 0;
	// the lane-per-pair kernel serves the read-length variants for everything but -ec (rows are edited in place), quality windows
	// above 8 and presence fragments longer than the packed adapter planes; SPG_OPT_KERNEL 1 keeps the warp-per-pair kernel
	const bool lanes = fi > 0 && ctx->kernel_layout != 1 && !p.ec && (p.qcut == 0 || p.qwin <= 8) && p.adapter_overlap <= ctx->a_size;
	if (lanes_query)
	{
		*lanes_query = lanes;
		return SPG_OK;
	}
	ctx->last_kernel.store(fi > 0 ? (lanes ? 2 : 1) * 100000 + nw * 1000 + full_hint : nw * 1000, std::memory_order_relaxed);
	if (lanes)
	{
		Device::Occ* locc = &d.lane_occ[nw_index(nw)][fi];
#define SPG_LANES(NW_, FL_) e = launch_lanes<NW_, FL_>(a, d.sm_count, ctx->ctas_per_sm, ctx->stages, stream, locc, ctx->mu)
		switch (nw * 16 + fi)
		{
			case 5 * 16 + 1: SPG_LANES(5, 150); break;
			case 5 * 16 + 2: SPG_LANES(5, 151); break;
			case 5 * 16 + 3: SPG_LANES(5, 100); break;
			case 5 * 16 + 4: SPG_LANES(5, 101); break;
			case 5 * 16 + 5: SPG_LANES(5, 125); break;
			case 5 * 16 + 6: SPG_LANES(5, 126); break;
			case 5 * 16 + 7: SPG_LANES(5, 75); break;
			case 5 * 16 + 8: SPG_LANES(5, 76); break;
			case 8 * 16 + 1: SPG_LANES(8, 250); break;
			case 8 * 16 + 2: SPG_LANES(8, 251); break;
			case 8 * 16 + 3: SPG_LANES(8, 200); break;
			case 8 * 16 + 4: SPG_LANES(8, 201); break;
			case 10 * 16 + 1: SPG_LANES(10, 300); break;
			default: SPG_LANES(10, 301); break;
		}
#undef SPG_LANES
	}
	else if (fi > 0)
	{
		Device::Occ* focc = &d.full_occ[nw_index(nw)][fi];
#define SPG_FULL(NW_, FL_) e = launch_cfg<NW_, SPG_FULL_MINB, FL_>(a, d.sm_count, ctx->ctas_per_sm, n_tiles, smem, stream, focc, ctx->mu)
		switch (nw * 16 + fi)
		{
			case 5 * 16 + 1: SPG_FULL(5, 150); break;
			case 5 * 16 + 2: SPG_FULL(5, 151); break;
			case 5 * 16 + 3: SPG_FULL(5, 100); break;
			case 5 * 16 + 4: SPG_FULL(5, 101); break;
			case 5 * 16 + 5: SPG_FULL(5, 125); break;
			case 5 * 16 + 6: SPG_FULL(5, 126); break;
			case 5 * 16 + 7: SPG_FULL(5, 75); break;
			case 5 * 16 + 8: SPG_FULL(5, 76); break;
			case 8 * 16 + 1: SPG_FULL(8, 250); break;
			case 8 * 16 + 2: SPG_FULL(8, 251); break;
			case 8 * 16 + 3: SPG_FULL(8, 200); break;
			case 8 * 16 + 4: SPG_FULL(8, 201); break;
			case 10 * 16 + 1: SPG_FULL(10, 300); break;
			default: SPG_FULL(10, 301); break;
		}
#undef SPG_FULL
	}
	else switch (nw)
	{
		case 5: e = launch_nw<5>(a, cw, d.sm_count, ctx->ctas_per_sm, n_tiles, smem, stream, occ, ctx->mu); break;
		case 8: e = launch_nw<8>(a, cw, d.sm_count, ctx->ctas_per_sm, n_tiles, smem, stream, occ, ctx->mu); break;
		case 10: e = launch_nw<10>(a, cw, d.sm_count, ctx->ctas_per_sm, n_tiles, smem, stream, occ, ctx->mu); break;
		case 16: e = launch_cfg<16, 2>(a, d.sm_count, ctx->ctas_per_sm, n_tiles, smem, stream, &d.occ[4][0], ctx->mu); break;
		case 32: e = launch_cfg<32, 1>(a, d.sm_count, ctx->ctas_per_sm, n_tiles, smem, stream, &d.occ[5][0], ctx->mu); break;
		default: e = launch_nw<0>(a, cw, d.sm_count, ctx->ctas_per_sm, n_tiles, smem, stream, occ, ctx->mu); break;
	}
	if (e != cudaSuccess) return fail(ctx, SPG_ERR_CUDA, std::string("trim_kernel launch: ") + cudaGetErrorString(e));
	{
		std::lock_guard<std::mutex> g(ctx->mu);
		++ctx->launches;
	}
	return SPG_OK;
}

// the statistics kernel keeps little state in registers, so its occupancy is set by the ring: 3 stages of the trimming kernel's tile
template <int NW>
cudaError_t launch_qc_cfg(spg::QcArgs& a, int sm_count, cudaStream_t stream, Device::Occ* occ_cache, std::mutex& mu)
{
	size_t smem;
	tile_geometry(a.stride, a.tile_pairs, a.stages, smem);
	a.stages = 3;
	smem = (size_t)a.stages * (4 * (size_t)a.tile_pairs * a.stride + 4 * (size_t)a.tile_pairs);
	int occ = 1;
	cudaError_t e0 = resident_ctas(spg::qc_kernel<NW, kCW>, *occ_cache, (kCW + 1) * 32, smem, mu, occ);
	if (e0 != cudaSuccess) return e0;
	const long long n_tiles = (a.n_pairs + a.tile_pairs - 1) / a.tile_pairs;
	const int grid = (int)std::min<long long>(n_tiles, (long long)sm_count * occ);
	spg::qc_kernel<NW, kCW><<<grid, (kCW + 1) * 32, smem, stream>>>(a);
	return cudaGetLastError();
}

// the lane-per-pair form of the statistics kernel (spg_qc_lanes.cuh): one CTA of 16-24 consumer warps + producer per SM, ring as deep as
// fits (a stage is one plane of one read of 32 pairs; every warp holds its stage during a whole pass, the rest of the ring is prefetch)
template <int NW, int CW>
cudaError_t launch_qc_lanes_cw(spg::QcArgs& a, int sm_count, cudaStream_t stream, Device::Occ* occ_cache, std::mutex& mu)
{
	const size_t stage = spg::qc_lane_stage_bytes(a.stride);
	int stages = spg::kQcLaneStagesMax;
	while (stages > 2 && stages * stage > 208 * 1024) --stages;
	a.stages = stages;
	a.tile_pairs = 32;
	const size_t smem = stages * stage;
	int occ = 1;
	cudaError_t e0 = resident_ctas(spg::qc_lanes_kernel<NW, CW>, *occ_cache, (CW + 1) * 32, smem, mu, occ);
	if (e0 != cudaSuccess) return e0;
	const long long n_tiles = (a.n_pairs + 31) / 32;
	const int grid = (int)std::min<long long>(n_tiles, (long long)sm_count * occ);
	spg::qc_lanes_kernel<NW, CW><<<grid, (CW + 1) * 32, smem, stream>>>(a);
	return cudaGetLastError();
}
template <int NW>
cudaError_t launch_qc_lanes_cfg(spg::QcArgs& a, int sm_count, cudaStream_t stream, Device::Occ* occ_cache, std::mutex& mu)
{
	// consumer warps (measured, profiles/qc_sweep_r2.txt): 24 for rows of up to 256 bytes; wider rows leave room for 20 stages only, and every
	// warp holds one during its pass
	if constexpr (NW <= 8) return launch_qc_lanes_cw<NW, 24>(a, sm_count, stream, occ_cache, mu);
	else return launch_qc_lanes_cw<NW, 16>(a, sm_count, stream, occ_cache, mu);
}

int launch_qc(spg_ctx* ctx, Device& d, const uint8_t* b1, const uint8_t* q1, const uint8_t* b2, const uint8_t* q2, const uint16_t* len1, const uint16_t* len2, int stride,
              long long n, cudaStream_t stream, const int* n_dev = nullptr, bool forward_only = false, int* bad_flag = nullptr)
{
	if (n <= 0) return SPG_OK;
	spg::QcArgs a;
	a.bad_flag = bad_flag;
	a.forward_only = forward_only ? 1 : 0;
	a.strict = (ctx->params.qc & 3) == 2 ? 1 : 0;
	a.plots = (ctx->params.qc & 4) ? 1 : 0;
	a.n_dev = n_dev;
	a.b1 = b1;
	a.q1 = q1;
	a.b2 = b2;
	a.q2 = q2;
	a.len1 = len1;
	a.len2 = len2;
	a.n_pairs = n;
	a.stride = stride;
	a.acc = d.d_qc;
	// Histogram(0, 60, 1)::binIndex of an integral mean quality k, evaluated in double exactly as the reference does
	// (floor((k - 0) / (60 - 0) * 60), src/cppCORE/Histogram.cpp:124): k / 60 * 60 may fall just below k
	for (int k = 0; k < 100; ++k)
	{
		const double hmin = 0.0, hmax = 60.0;
		volatile double x = ((double)k - hmin) / (hmax - hmin);
		volatile double y = x * 60;
		long b = (long)std::floor(y);
		a.bin_of_int[k] = (uint8_t)std::min(59L, std::max(0L, b));
	}
	cudaError_t e = cudaSuccess;
	const bool lanes = ctx->kernel_layout != 1; // SPG_OPT_KERNEL 1 keeps the warp-per-pair form
	switch (nw_for_qc(stride))
	{
		case 5: e = lanes ? launch_qc_lanes_cfg<5>(a, d.sm_count, stream, &d.qc_lane_occ[0], ctx->mu) : launch_qc_cfg<5>(a, d.sm_count, stream, &d.qc_occ[0], ctx->mu); break;
		case 8: e = lanes ? launch_qc_lanes_cfg<8>(a, d.sm_count, stream, &d.qc_lane_occ[1], ctx->mu) : launch_qc_cfg<8>(a, d.sm_count, stream, &d.qc_occ[1], ctx->mu); break;
		case 10: e = lanes ? launch_qc_lanes_cfg<10>(a, d.sm_count, stream, &d.qc_lane_occ[2], ctx->mu) : launch_qc_cfg<10>(a, d.sm_count, stream, &d.qc_occ[2], ctx->mu); break;
		default:
		{
			const long long warps_wanted = std::min<long long>((n + 7) / 8, (long long)d.sm_count * 32); // >= 8 pairs per warp, at most 4 CTAs of 8 warps per SM
			spg::qc_kernel_generic<<<(int)((warps_wanted + 7) / 8), 256, 0, stream>>>(a);
			e = cudaGetLastError();
		}
	}
	if (e != cudaSuccess) return fail(ctx, SPG_ERR_CUDA, std::string("qc_kernel launch: ") + cudaGetErrorString(e));
	std::lock_guard<std::mutex> g(ctx->mu);
	++ctx->launches;
	return SPG_OK;
}

} // namespace

extern "C"
{

int spg_create(spg_ctx** out, const spg_params* params, const int* device_ids, int n_devices, int n_slots, int max_pairs, int max_len)
{
	if (!out || !params) return fail(nullptr, SPG_ERR_PARAM, "null argument");
	*out = nullptr;
	if (!params->a1 || !params->a2 || params->a1_len < 15 || params->a2_len < 15)
		return fail(nullptr, SPG_ERR_PARAM, "adapters must have at least 15 bases"); // main.cpp:68,70
	if (params->adapter_overlap < 1 || params->adapter_overlap > 32 || params->adapter_overlap > std::min(params->a1_len, params->a2_len))
		return fail(nullptr, SPG_ERR_PARAM, "adapter_overlap out of range");
	if (params->qcut > 0 && (params->qwin < 1 || params->qwin >= SPG_MAXLEN)) return fail(nullptr, SPG_ERR_PARAM, "qwin out of range");
	if (params->ncut < 0 || params->qcut < 0) return fail(nullptr, SPG_ERR_PARAM, "negative qcut/ncut");
	if (!(params->match_perc == params->match_perc) || !(params->mep == params->mep)) return fail(nullptr, SPG_ERR_PARAM, "match_perc/mep is NaN");
	if (n_devices < 1 || !device_ids) return fail(nullptr, SPG_ERR_PARAM, "at least one device is required");
	if (n_slots < 0 || (n_slots > 0 && (max_pairs < 1 || max_len < 1))) return fail(nullptr, SPG_ERR_PARAM, "invalid slot geometry");
	if (max_len >= SPG_MAXLEN) return fail(nullptr, SPG_ERR_PARAM, "max_len must be below 1000 (MAXLEN of the reference)");

	int dev_count = 0;
	cudaError_t ce = cudaGetDeviceCount(&dev_count);
	if (ce != cudaSuccess || dev_count == 0)
		return fail(nullptr, SPG_ERR_CUDA, std::string("no CUDA device available (this library has no CPU path): ") + cudaGetErrorString(ce));
	for (int i = 0; i < n_devices; ++i)
		if (device_ids[i] < 0 || device_ids[i] >= dev_count) return fail(nullptr, SPG_ERR_PARAM, "device id out of range");

	spg_ctx* ctx = new (std::nothrow) spg_ctx();
	if (!ctx) return fail(nullptr, SPG_ERR_NOMEM, "out of memory");
	ctx->params = *params;
	ctx->a1.assign(params->a1, (size_t)params->a1_len);
	ctx->a2.assign(params->a2, (size_t)params->a2_len);
	ctx->params.a1 = ctx->a1.data();
	ctx->params.a2 = ctx->a2.data();
	ctx->a_size = std::min(20, std::min(params->a1_len, params->a2_len)); // main.cpp:71
	ctx->adapters_plain = all_acgtn(ctx->a1.data(), std::min(32, params->a1_len)) && all_acgtn(ctx->a2.data(), std::min(32, params->a2_len));
	std::string err;
	if (!build_tables(ctx->params, ctx->a_size, ctx->tables, err))
	{
		delete ctx;
		return fail(nullptr, SPG_ERR_PARAM, err);
	}
	if (const char* e = getenv("SPG_ZERO_COPY_QUALS")) ctx->zero_copy_quals = atoi(e) ? 1 : 0;                                       // tuning runs only
	if (const char* e = getenv("SPG_N_LANES")) ctx->n_lanes = atoi(e) ? 1 : 0;                                                        // tuning runs only
	if (const char* e = getenv("SPG_STAGES")) ctx->stages = std::max(0, std::min(atoi(e), (int)spg::kLaneStagesMax)); // tuning runs only
	ctx->max_pairs = max_pairs;
	ctx->max_len = max_len;
	ctx->stride = n_slots > 0 ? std::max(16, (max_len + 1) / 2 * 2) : 0; // even: a tile of 8 rows is then a multiple of 16 bytes (TMA)
	const int cap = (max_pairs + 7) / 8 * 8;
	ctx->plane_bytes = (size_t)cap * ctx->stride;
	ctx->len_bytes = (size_t)cap * sizeof(uint16_t);
	ctx->tail_bytes = (size_t)cap * SPG_QTAIL;
	ctx->block_bytes = 4 * ctx->plane_bytes + 2 * ctx->len_bytes + 2 * ctx->tail_bytes;

#define CREATE_CUDA(call)                                                                     \
	do                                                                                        \
	{                                                                                         \
		cudaError_t e_ = (call);                                                              \
		if (e_ != cudaSuccess)                                                                \
		{                                                                                     \
			std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_);              \
			spg_destroy(ctx);                                                                 \
			return fail(nullptr, SPG_ERR_CUDA, m_);                                           \
		}                                                                                     \
	} while (0)

	ctx->devs.resize((size_t)n_devices);
	for (int i = 0; i < n_devices; ++i)
	{
		Device& d = ctx->devs[(size_t)i];
		d.id = device_ids[i];
		CREATE_CUDA(cudaSetDevice(d.id));
		cudaDeviceProp prop;
		CREATE_CUDA(cudaGetDeviceProperties(&prop, d.id));
		if (prop.major < 10)
		{
			spg_destroy(ctx);
			return fail(nullptr, SPG_ERR_CUDA, "device is not sm_100 (this library is built for B200 only)");
		}
		d.sm_count = prop.multiProcessorCount;
		CREATE_CUDA(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
		CREATE_CUDA(cudaMalloc(&d.d_mmin, ctx->tables.mmin.size() * sizeof(uint16_t)));
		CREATE_CUDA(cudaMalloc(&d.d_rank, ctx->tables.ranktab.size() * sizeof(uint16_t)));
		CREATE_CUDA(cudaMalloc(&d.d_psmall, ctx->tables.psmall.size() * sizeof(double)));
		CREATE_CUDA(cudaMalloc(&d.d_ec, 3 * SPG_MAXLEN * sizeof(unsigned long long)));
		CREATE_CUDA(cudaMemcpy(d.d_mmin, ctx->tables.mmin.data(), ctx->tables.mmin.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
		CREATE_CUDA(cudaMemcpy(d.d_rank, ctx->tables.ranktab.data(), ctx->tables.ranktab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
		CREATE_CUDA(cudaMemcpy(d.d_psmall, ctx->tables.psmall.data(), ctx->tables.psmall.size() * sizeof(double), cudaMemcpyHostToDevice));
		CREATE_CUDA(cudaMemset(d.d_ec, 0, 3 * SPG_MAXLEN * sizeof(unsigned long long)));
		CREATE_CUDA(cudaMalloc(&d.d_qc, spg::kQcWords * sizeof(unsigned long long)));
		CREATE_CUDA(cudaMemset(d.d_qc, 0, spg::kQcWords * sizeof(unsigned long long)));
	}
	ctx->slots.resize((size_t)n_slots);
	for (int s = 0; s < n_slots; ++s)
	{
		Slot& sl = ctx->slots[(size_t)s];
		sl.dev = s % n_devices;
		CREATE_CUDA(cudaSetDevice(ctx->devs[(size_t)sl.dev].id));
		CREATE_CUDA(cudaHostAlloc(&sl.h_block, ctx->block_bytes, cudaHostAllocPortable | cudaHostAllocMapped));
		CREATE_CUDA(cudaHostGetDevicePointer((void**)&sl.h_block_dev, sl.h_block, 0));
		CREATE_CUDA(cudaHostAlloc(&sl.h_res, (size_t)cap * sizeof(spg_result), cudaHostAllocPortable));
		CREATE_CUDA(cudaMalloc(&sl.d_block, ctx->block_bytes));
		CREATE_CUDA(cudaMalloc(&sl.d_res, (size_t)cap * sizeof(spg_result)));
		CREATE_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
		CREATE_CUDA(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
		memset(sl.h_block, 0, ctx->block_bytes);
	}
#undef CREATE_CUDA
	*out = ctx;
	return SPG_OK;
}

int spg_slot_buffers(spg_ctx* ctx, int slot, spg_slot_view* v)
{
	if (!ctx || !v) return SPG_ERR_PARAM;
	if (slot < 0 || slot >= (int)ctx->slots.size()) return fail(ctx, SPG_ERR_PARAM, "slot out of range");
	Slot& sl = ctx->slots[(size_t)slot];
	v->bases1 = sl.h_block;
	v->quals1 = sl.h_block + ctx->plane_bytes;
	v->bases2 = sl.h_block + 2 * ctx->plane_bytes;
	v->quals2 = sl.h_block + 3 * ctx->plane_bytes;
	v->len1 = reinterpret_cast<uint16_t*>(sl.h_block + 4 * ctx->plane_bytes);
	v->len2 = reinterpret_cast<uint16_t*>(sl.h_block + 4 * ctx->plane_bytes + ctx->len_bytes);
	v->stride = ctx->stride;
	v->max_pairs = ctx->max_pairs;
	return SPG_OK;
}

int spg_slot_qtails(spg_ctx* ctx, int slot, uint8_t** qtail1, uint8_t** qtail2)
{
	if (!ctx || !qtail1 || !qtail2) return SPG_ERR_PARAM;
	if (slot < 0 || slot >= (int)ctx->slots.size()) return fail(ctx, SPG_ERR_PARAM, "slot out of range");
	Slot& sl = ctx->slots[(size_t)slot];
	*qtail1 = sl.h_block + 4 * ctx->plane_bytes + 2 * ctx->len_bytes;
	*qtail2 = *qtail1 + ctx->tail_bytes;
	return SPG_OK;
}

int spg_submit(spg_ctx* ctx, int slot, int n_pairs)
{
	if (!ctx) return SPG_ERR_PARAM;
	if (slot < 0 || slot >= (int)ctx->slots.size()) return fail(ctx, SPG_ERR_PARAM, "slot out of range");
	if (n_pairs < 0 || n_pairs > ctx->max_pairs) return fail(ctx, SPG_ERR_PARAM, "n_pairs out of range");
	Slot& sl = ctx->slots[(size_t)slot];
	if (sl.state != SLOT_IDLE) return fail(ctx, SPG_ERR_STATE, "slot was submitted and not waited for");
	Device& d = ctx->devs[(size_t)sl.dev];
	SPG_CUDA(ctx, cudaSetDevice(d.id));
	sl.n = n_pairs;
	if (n_pairs > 0)
	{
		const size_t rows = (size_t)n_pairs * ctx->stride;
		const size_t lens = (size_t)((n_pairs + 7) / 8 * 8) * sizeof(uint16_t);
		const size_t pb = ctx->plane_bytes;
		// The lane-per-pair kernel reads only the last qualities of a read (quality trimming): the quality planes then stay in the
		// pinned slot and the kernel fetches the few sectors it needs over PCIe itself (zero copy: pinned memory is mapped into the
		// device's address space), which cuts the bytes per pair on the link from 4L+4 to 2L+4 plus those sectors. Not with -qc
		// (the statistics kernel reads every quality) and not for the other kernels (they stage whole quality rows).
		bool zero_copy = false;
		if (ctx->zero_copy_quals && !ctx->params.qc)
		{
			int qrc = launch_trim(ctx, d, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ctx->stride, n_pairs, nullptr, sl.stream, nullptr, -1, &zero_copy);
			if (qrc != SPG_OK) return qrc;
		}
		sl.zero_copy = zero_copy;
		const bool tails = zero_copy && ctx->qual_tails;
		if (zero_copy)
		{
			SPG_CUDA(ctx, cudaMemcpyAsync(sl.d_block, sl.h_block, rows, cudaMemcpyHostToDevice, sl.stream));
			SPG_CUDA(ctx, cudaMemcpyAsync(sl.d_block + 2 * pb, sl.h_block + 2 * pb, rows, cudaMemcpyHostToDevice, sl.stream));
			const size_t lb = ctx->len_bytes, tb = ctx->tail_bytes;
			if (2 * (size_t)n_pairs >= (size_t)ctx->max_pairs) // lengths [and tails] of a well-filled slot in one copy (they lie side by side)
			{
				SPG_CUDA(ctx, cudaMemcpyAsync(sl.d_block + 4 * pb, sl.h_block + 4 * pb, 2 * lb + (tails ? 2 * tb : 0), cudaMemcpyHostToDevice, sl.stream));
			}
			else
			{
				SPG_CUDA(ctx, cudaMemcpyAsync(sl.d_block + 4 * pb, sl.h_block + 4 * pb, lens, cudaMemcpyHostToDevice, sl.stream));
				SPG_CUDA(ctx, cudaMemcpyAsync(sl.d_block + 4 * pb + lb, sl.h_block + 4 * pb + lb, lens, cudaMemcpyHostToDevice, sl.stream));
				if (tails)
				{
					const size_t tn = (size_t)n_pairs * SPG_QTAIL;
					SPG_CUDA(ctx, cudaMemcpyAsync(sl.d_block + 4 * pb + 2 * lb, sl.h_block + 4 * pb + 2 * lb, tn, cudaMemcpyHostToDevice, sl.stream));
					SPG_CUDA(ctx, cudaMemcpyAsync(sl.d_block + 4 * pb + 2 * lb + tb, sl.h_block + 4 * pb + 2 * lb + tb, tn, cudaMemcpyHostToDevice, sl.stream));
				}
			}
		}
		else if (n_pairs == ctx->max_pairs)
		{
			SPG_CUDA(ctx, cudaMemcpyAsync(sl.d_block, sl.h_block, ctx->block_bytes, cudaMemcpyHostToDevice, sl.stream));
		}
		else
		{
			for (int k = 0; k < 4; ++k) SPG_CUDA(ctx, cudaMemcpyAsync(sl.d_block + k * pb, sl.h_block + k * pb, rows, cudaMemcpyHostToDevice, sl.stream));
			SPG_CUDA(ctx, cudaMemcpyAsync(sl.d_block + 4 * pb, sl.h_block + 4 * pb, lens, cudaMemcpyHostToDevice, sl.stream));
			SPG_CUDA(ctx, cudaMemcpyAsync(sl.d_block + 4 * pb + ctx->len_bytes, sl.h_block + 4 * pb + ctx->len_bytes, lens, cudaMemcpyHostToDevice, sl.stream));
		}
		if (ctx->params.qc) // raw-read statistics of the untrimmed batch, in front of the trimming kernel on the same stream
		{
			int qrc = launch_qc(ctx, d, sl.d_block, sl.d_block + pb, sl.d_block + 2 * pb, sl.d_block + 3 * pb, reinterpret_cast<const uint16_t*>(sl.d_block + 4 * pb),
			                    reinterpret_cast<const uint16_t*>(sl.d_block + 4 * pb + ctx->len_bytes), ctx->stride, n_pairs, sl.stream);
			if (qrc != SPG_OK) return qrc;
		}
		uint8_t* const q1 = zero_copy ? sl.h_block_dev + pb : sl.d_block + pb;
		uint8_t* const q2 = zero_copy ? sl.h_block_dev + 3 * pb : sl.d_block + 3 * pb;
		int rc = launch_trim(ctx, d, sl.d_block, q1, sl.d_block + 2 * pb, q2, reinterpret_cast<const uint16_t*>(sl.d_block + 4 * pb),
		                     reinterpret_cast<const uint16_t*>(sl.d_block + 4 * pb + ctx->len_bytes), ctx->stride, n_pairs, sl.d_res, sl.stream, nullptr, -1, nullptr, zero_copy,
		                     tails ? sl.d_block + 4 * pb + 2 * ctx->len_bytes : nullptr, tails ? sl.d_block + 4 * pb + 2 * ctx->len_bytes + ctx->tail_bytes : nullptr);
		if (rc != SPG_OK) return rc;
		SPG_CUDA(ctx, cudaMemcpyAsync(sl.h_res, sl.d_res, (size_t)n_pairs * sizeof(spg_result), cudaMemcpyDeviceToHost, sl.stream));
		if (ctx->params.ec) // edited rows come back in the slot
		{
			for (int k = 0; k < 4; ++k) SPG_CUDA(ctx, cudaMemcpyAsync(sl.h_block + k * pb, sl.d_block + k * pb, rows, cudaMemcpyDeviceToHost, sl.stream));
		}
	}
	SPG_CUDA(ctx, cudaEventRecord(sl.done, sl.stream));
	sl.state = SLOT_SUBMITTED;
	return SPG_OK;
}

int spg_wait(spg_ctx* ctx, int slot, const spg_result** results)
{
	if (!ctx) return SPG_ERR_PARAM;
	if (slot < 0 || slot >= (int)ctx->slots.size()) return fail(ctx, SPG_ERR_PARAM, "slot out of range");
	Slot& sl = ctx->slots[(size_t)slot];
	if (sl.state != SLOT_SUBMITTED) return fail(ctx, SPG_ERR_STATE, "slot was not submitted");
	SPG_CUDA(ctx, cudaEventSynchronize(sl.done));
	sl.state = SLOT_IDLE;
	if (results) *results = sl.h_res;
	return SPG_OK;
}

int spg_trim_device(spg_ctx* ctx, int device_index, void* bases1, void* quals1, void* bases2, void* quals2, const uint16_t* len1, const uint16_t* len2, int stride,
                    int64_t n_pairs, spg_result* results, void* cuda_stream)
{
	if (!ctx) return SPG_ERR_PARAM;
	if (device_index < 0 || device_index >= (int)ctx->devs.size()) return fail(ctx, SPG_ERR_PARAM, "device index out of range");
	if (stride < 16 || stride % 2 != 0 || stride > 1008) return fail(ctx, SPG_ERR_PARAM, "stride must be even and in [16,1008]");
	const uintptr_t bits = (uintptr_t)bases1 | (uintptr_t)quals1 | (uintptr_t)bases2 | (uintptr_t)quals2 | (uintptr_t)len1 | (uintptr_t)len2;
	if (bits & 15u) return fail(ctx, SPG_ERR_PARAM, "device pointers must be 16-byte aligned");
	if ((uintptr_t)results & 7u) return fail(ctx, SPG_ERR_PARAM, "results must be 8-byte aligned");
	if (n_pairs < 0) return fail(ctx, SPG_ERR_PARAM, "negative n_pairs");
	Device& d = ctx->devs[(size_t)device_index];
	SPG_CUDA(ctx, cudaSetDevice(d.id));
	return launch_trim(ctx, d, (uint8_t*)bases1, (uint8_t*)quals1, (uint8_t*)bases2, (uint8_t*)quals2, len1, len2, stride, (long long)n_pairs, results,
	                   (cudaStream_t)cuda_stream);
}

int spg_ec_stats_get(spg_ctx* ctx, spg_ec_stats* out)
{
	if (!ctx || !out) return SPG_ERR_PARAM;
	memset(out, 0, sizeof(*out));
	std::vector<unsigned long long> tmp(3 * SPG_MAXLEN);
	for (Device& d : ctx->devs)
	{
		SPG_CUDA(ctx, cudaSetDevice(d.id));
		SPG_CUDA(ctx, cudaMemcpy(tmp.data(), d.d_ec, tmp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
		for (int i = 0; i < SPG_MAXLEN; ++i)
		{
			out->mismatch_r1[i] += (int64_t)tmp[(size_t)i];
			out->mismatch_r2[i] += (int64_t)tmp[(size_t)SPG_MAXLEN + i];
			out->errors_per_read[i] += (int64_t)tmp[(size_t)2 * SPG_MAXLEN + i];
		}
	}
	return SPG_OK;
}

int spg_qc_device(spg_ctx* ctx, int device_index, const void* bases1, const void* quals1, const void* bases2, const void* quals2, const uint16_t* len1,
                  const uint16_t* len2, int stride, int64_t n_pairs, void* cuda_stream)
{
	if (!ctx) return SPG_ERR_PARAM;
	if (device_index < 0 || device_index >= (int)ctx->devs.size()) return fail(ctx, SPG_ERR_PARAM, "device index out of range");
	if (stride < 16 || stride % 2 != 0 || stride > 1008) return fail(ctx, SPG_ERR_PARAM, "stride must be even and in [16,1008]");
	if (n_pairs < 0) return fail(ctx, SPG_ERR_PARAM, "negative n_pairs");
	const uintptr_t bits = (uintptr_t)bases1 | (uintptr_t)quals1 | (uintptr_t)bases2 | (uintptr_t)quals2 | (uintptr_t)len1 | (uintptr_t)len2;
	if (bits & 15u) return fail(ctx, SPG_ERR_PARAM, "device pointers must be 16-byte aligned");
	Device& d = ctx->devs[(size_t)device_index];
	SPG_CUDA(ctx, cudaSetDevice(d.id));
	return launch_qc(ctx, d, (const uint8_t*)bases1, (const uint8_t*)quals1, (const uint8_t*)bases2, (const uint8_t*)quals2, len1, len2, stride, (long long)n_pairs,
	                 (cudaStream_t)cuda_stream);
}

int spg_qc_stats_get(spg_ctx* ctx, spg_qc_stats* out)
{
	if (!ctx || !out) return SPG_ERR_PARAM;
	memset(out, 0, sizeof(*out));
	std::vector<unsigned long long> tmp((size_t)spg::kQcWords);
	for (Device& d : ctx->devs)
	{
		SPG_CUDA(ctx, cudaSetDevice(d.id));
		SPG_CUDA(ctx, cudaDeviceSynchronize());
		SPG_CUDA(ctx, cudaMemcpy(tmp.data(), d.d_qc, tmp.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
		out->reads_forward += (int64_t)tmp[spg::kQcReadsF];
		out->reads_reverse += (int64_t)tmp[spg::kQcReadsR];
		out->bases_sequenced += (int64_t)tmp[spg::kQcBases];
		out->read_q20 += (int64_t)tmp[spg::kQcReadQ20];
		out->base_q20 += (int64_t)tmp[spg::kQcBaseQ20];
		out->base_q30 += (int64_t)tmp[spg::kQcBaseQ30];
		out->errors += (int64_t)tmp[spg::kQcErrors];
		for (int i = 0; i < SPG_MAXLEN; ++i)
		{
			out->read_lengths[i] += (int64_t)tmp[(size_t)spg::kQcLen + i];
			for (int k = 0; k < 5; ++k) out->pileup[i][k] += (int64_t)tmp[(size_t)spg::kQcPile + 5 * i + k];
			out->qsum_forward[i] += (int64_t)tmp[(size_t)spg::kQcQf + i];
			out->qsum_reverse[i] += (int64_t)tmp[(size_t)spg::kQcQr + i];
		}
		for (int i = 0; i < 100; ++i)
		{
			out->base_qualities[i] += (int64_t)tmp[(size_t)spg::kQcBaseQual + i];
			out->read_qualities[i] += (int64_t)tmp[(size_t)spg::kQcReadQual + i];
		}
		for (int i = 0; i < 60; ++i)
		{
			out->qscore_dist_forward[i] += (int64_t)tmp[(size_t)spg::kQcDistF + i];
			out->qscore_dist_reverse[i] += (int64_t)tmp[(size_t)spg::kQcDistR + i];
		}
	}
	return SPG_OK;
}

const char* spg_last_error(spg_ctx* ctx)
{
	if (!ctx) return g_create_error.c_str();
	std::lock_guard<std::mutex> g(ctx->mu);
	return ctx->err.c_str();
}

void spg_destroy(spg_ctx* ctx)
{
	if (!ctx) return;
	while (!ctx->fqs.empty()) fq_free(ctx->fqs.back());
	for (Slot& sl : ctx->slots)
	{
		if (sl.dev < (int)ctx->devs.size() && ctx->devs[(size_t)sl.dev].id >= 0) cudaSetDevice(ctx->devs[(size_t)sl.dev].id);
		if (sl.done)
		{
			cudaEventSynchronize(sl.done);
			cudaEventDestroy(sl.done);
		}
		if (sl.stream) cudaStreamDestroy(sl.stream);
		if (sl.h_block) cudaFreeHost(sl.h_block);
		if (sl.h_res) cudaFreeHost(sl.h_res);
		if (sl.d_block) cudaFree(sl.d_block);
		if (sl.d_res) cudaFree(sl.d_res);
	}
	for (Device& d : ctx->devs)
	{
		if (d.id < 0) continue;
		cudaSetDevice(d.id);
		if (d.stream)
		{
			cudaStreamSynchronize(d.stream);
			cudaStreamDestroy(d.stream);
		}
		cudaFree(d.d_mmin);
		cudaFree(d.d_rank);
		cudaFree(d.d_psmall);
		cudaFree(d.d_ec);
		cudaFree(d.d_qc);
	}
	delete ctx;
}

int spg_set_option(spg_ctx* ctx, int option, int value)
{
	if (!ctx) return SPG_ERR_PARAM;
	switch (option)
	{
		case SPG_OPT_FORCE_BYTEWISE: ctx->force_bytewise = value ? 1 : 0; return SPG_OK;
		case SPG_OPT_GRID_CTAS_PER_SM: ctx->ctas_per_sm = value; return SPG_OK;
		case SPG_OPT_FULL_LEN: ctx->full_len = value < 0 ? -1 : value; return SPG_OK;
		case SPG_OPT_SEED_SCAN: ctx->seed_scan = value ? 1 : 0; return SPG_OK;
		case SPG_OPT_ZERO_COPY_QUALS: ctx->zero_copy_quals = value ? 1 : 0; return SPG_OK;
		case SPG_OPT_N_LANES: ctx->n_lanes = value ? 1 : 0; return SPG_OK;
		case SPG_OPT_QUAL_TAILS: ctx->qual_tails = value ? 1 : 0; return SPG_OK;
		case SPG_OPT_KERNEL:
			if (value < 0 || value > 2) return fail(ctx, SPG_ERR_PARAM, "kernel layout must be 0 (automatic), 1 (warp per pair) or 2 (lane per pair)");
			ctx->kernel_layout = value;
			return SPG_OK;
		case SPG_OPT_MIN_BLOCKS:
			if (value != 2 && value != 3 && value != 4) return fail(ctx, SPG_ERR_PARAM, "min blocks must be 2, 3 or 4");
			ctx->min_blocks = value;
			return SPG_OK;
		case SPG_OPT_TILE_PAIRS:
			if (value < 0 || value % 8 != 0 || value > 256) return fail(ctx, SPG_ERR_PARAM, "tile pairs must be a multiple of 8, at most 256");
			ctx->tile_pairs = value;
			for (Device& d : ctx->devs)
			{
				for (auto& row : d.occ) for (auto& c : row) c.ctas = 0;
				for (auto& row : d.full_occ) for (auto& c : row) c.ctas = 0;
			}
			return SPG_OK;
		case SPG_OPT_STAGES:
			if (value < 0 || value > spg::kLaneStagesMax || value == 1) return fail(ctx, SPG_ERR_PARAM, "stages must be 2..8 (2..4 for the warp-per-pair kernels)");
			ctx->stages = value;
			for (Device& d : ctx->devs)
			{
				for (auto& row : d.occ) for (auto& c : row) c.ctas = 0;
				for (auto& row : d.full_occ) for (auto& c : row) c.ctas = 0;
			}
			return SPG_OK;
		default: return fail(ctx, SPG_ERR_PARAM, "unknown option");
	}
}

int spg_get_option(spg_ctx* ctx, int option, int* value)
{
	if (!ctx || !value) return SPG_ERR_PARAM;
	switch (option)
	{
		case SPG_OPT_FORCE_BYTEWISE: *value = ctx->force_bytewise; return SPG_OK;
		case SPG_OPT_GRID_CTAS_PER_SM: *value = ctx->ctas_per_sm; return SPG_OK;
		case SPG_OPT_MIN_BLOCKS: *value = ctx->min_blocks; return SPG_OK;
		case SPG_OPT_TILE_PAIRS: *value = ctx->tile_pairs; return SPG_OK;
		case SPG_OPT_STAGES: *value = ctx->stages; return SPG_OK;
		case SPG_OPT_FULL_LEN: *value = ctx->full_len; return SPG_OK;
		case SPG_OPT_KERNEL: *value = ctx->kernel_layout; return SPG_OK;
		case SPG_OPT_SEED_SCAN: *value = ctx->seed_scan; return SPG_OK;
		case SPG_OPT_ZERO_COPY_QUALS: *value = ctx->zero_copy_quals; return SPG_OK;
		case SPG_OPT_N_LANES: *value = ctx->n_lanes; return SPG_OK;
		case SPG_OPT_QUAL_TAILS: *value = ctx->qual_tails; return SPG_OK;
		default: return fail(ctx, SPG_ERR_PARAM, "unknown option");
	}
}

int spg_last_kernel(spg_ctx* ctx, char* name, int cap)
{
	if (!ctx) return SPG_ERR_PARAM;
	const int v = ctx->last_kernel.load(std::memory_order_relaxed);
	const int layout = v / 100000, nw = (v / 1000) % 100, full = v % 1000;
	if (name && cap > 0)
	{
		if (v == 0) snprintf(name, (size_t)cap, "none");
		else if (layout == 2) snprintf(name, (size_t)cap, "spg::trim_lanes_kernel<NW=%d,FULL=%d,CW=%d,MINB=%d>", nw, full, nw <= 5 ? LaneCfg<5>::CW : LaneCfg<8>::CW, nw <= 5 ? LaneCfg<5>::MINB : LaneCfg<8>::MINB);
		else snprintf(name, (size_t)cap, "spg::trim_kernel<NW=%d,CW=%d,MINB=%d,FULL=%d>", nw, kCW, nw == 16 ? 2 : nw == 32 ? 1 : ctx->min_blocks, layout == 1 ? full : 0);
	}
	return v;
}

int64_t spg_launch_count(spg_ctx* ctx)
{
	if (!ctx) return 0;
	std::lock_guard<std::mutex> g(ctx->mu);
	return ctx->launches;
}

} // extern "C"

#include "spg_fastq_engine.inc"
