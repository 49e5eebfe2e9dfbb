// spg_synth.cu -- synthetic paired-end reads generated on the device (bench / test input).
//
// Counter based: every byte is a pure function of (seed, pair index, read, position), so any slice [first, first+n) of a
// stream can be regenerated independently (e.g. to hand the same pairs to the CPU oracle after a D2H copy).
// Model: fragment of `insert` uniform ACGT bases; read 1 = fragment[0..L) then adapter a1 then random filler when the
// insert is shorter than the read; read 2 = revcomp(fragment)[0..L) then a2 then filler; i.i.d. substitutions and Ns;
// Q40 ('I') qualities with an exponential-length low-quality ('#') 3' tail, optional NovaSeq-like binned qualities and
// injected N runs.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/seqpurge_b200.h"

namespace
{

struct SynthArgs
{
	spg_synth_config cfg;
	long long first;
	long long n;
	uint8_t *b1, *q1, *b2, *q2;
	uint16_t *len1, *len2;
	int stride;
	uint8_t a1[33], a2[33];
	int a1_len, a2_len;
};

__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
	x ^= x >> 30;
	x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27;
	x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}
__device__ __forceinline__ uint64_t rnd(uint64_t seed, uint64_t pair, uint32_t stream, uint32_t idx)
{
	return mix64(mix64(seed ^ (pair * 0x9E3779B97F4A7C15ull + stream)) + (uint64_t)idx * 0xD1B54A32D192ED03ull);
}
__device__ __forceinline__ float u01(uint64_t r) { return ((float)(uint32_t)(r >> 40) + 0.5f) * (1.0f / 16777216.0f); }

__device__ __forceinline__ uint32_t base_char(uint32_t code) { return (0x54474341u >> (8 * (code & 3u))) & 0xFFu; } // "ACGT"
__device__ __forceinline__ uint32_t base_code(uint32_t c) { return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u; }

__global__ void synth_kernel(const __grid_constant__ SynthArgs A)
{
	const int lane = threadIdx.x & 31;
	const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
	const spg_synth_config& C = A.cfg;
	const int L = C.read_len;
	for (long long r = warp; r < A.n; r += n_warps)
	{
		const uint64_t pair = (uint64_t)(A.first + r);
		// insert size: Irwin-Hall(12) approximation of a normal deviate
		int insert;
		{
			float s = 0.f;
			for (int k = 0; k < 12; ++k) s += u01(rnd(C.seed, pair, 7, (uint32_t)k));
			float z = s - 6.0f;
			insert = (int)lrintf(C.insert_mean + C.insert_sd * z);
			insert = max(C.insert_min, min(C.insert_max, insert));
		}
		for (int read = 0; read < 2; ++read)
		{
			uint8_t* brow = (read == 0 ? A.b1 : A.b2) + (size_t)r * A.stride;
			uint8_t* qrow = (read == 0 ? A.q1 : A.q2) + (size_t)r * A.stride;
			const uint8_t* adapter = read == 0 ? A.a1 : A.a2;
			const int alen = read == 0 ? A.a1_len : A.a2_len;
			int tail = 0;
			if (C.lowq_tail_mean > 0.f) tail = (int)floorf(-logf(u01(rnd(C.seed, pair, 10 + read, 0))) * C.lowq_tail_mean);
			int run_start = -1, run_len = 0;
			if (C.n_run_rate > 0.f && u01(rnd(C.seed, pair, 12 + read, 0)) < C.n_run_rate)
			{
				uint64_t rr = rnd(C.seed, pair, 12 + read, 1);
				run_len = 7 + (int)(rr % 6u);
				run_start = (int)((rr >> 16) % (uint64_t)max(1, L - run_len));
			}
			for (int j = lane; j < A.stride; j += 32)
			{
				uint32_t b = 0, q = 0;
				if (j < L)
				{
					uint32_t code;
					if (j < insert)
					{
						if (read == 0) code = (uint32_t)rnd(C.seed, pair, 0, (uint32_t)j) & 3u;
						else code = 3u - ((uint32_t)rnd(C.seed, pair, 0, (uint32_t)(insert - 1 - j)) & 3u); // complement in ACGT order
						b = base_char(code);
					}
					else if (j - insert < alen)
					{
						b = adapter[j - insert];
						code = base_code(b);
					}
					else
					{
						code = (uint32_t)rnd(C.seed, pair, 1 + read, (uint32_t)j) & 3u;
						b = base_char(code);
					}
					const uint64_t e = rnd(C.seed, pair, 3 + read, (uint32_t)j);
					if (u01(e) < C.error_rate) b = base_char(code + 1u + (uint32_t)((e >> 8) % 3u));
					if (u01(e << 24) < C.n_rate) b = 'N';
					q = 'I';
					if (C.binned_quals)
					{
						float u = u01(rnd(C.seed, pair, 5 + read, (uint32_t)j));
						q = u < 0.80f ? 'F' : u < 0.93f ? ':' : u < 0.98f ? ',' : '#';
					}
					if (j >= L - tail) q = '#';
					if (run_start >= 0 && j >= run_start && j < run_start + run_len)
					{
						b = 'N';
						q = '#';
					}
				}
				brow[j] = (uint8_t)b;
				qrow[j] = (uint8_t)q;
			}
		}
		if (lane == 0)
		{
			A.len1[r] = (uint16_t)L;
			A.len2[r] = (uint16_t)L;
		}
	}
}

} // namespace

extern "C" int spg_synth_device(int device_id, const spg_synth_config* cfg, int64_t first_pair, int64_t n_pairs, void* bases1, void* quals1, void* bases2, void* quals2,
                                uint16_t* len1, uint16_t* len2, int stride, void* cuda_stream)
{
	if (!cfg || cfg->read_len < 1 || cfg->read_len >= SPG_MAXLEN || stride < cfg->read_len || n_pairs < 0) return SPG_ERR_PARAM;
	if (n_pairs == 0) return SPG_OK;
	if (cudaSetDevice(device_id) != cudaSuccess) return SPG_ERR_CUDA;
	SynthArgs a;
	a.cfg = *cfg;
	a.first = first_pair;
	a.n = n_pairs;
	a.b1 = (uint8_t*)bases1;
	a.q1 = (uint8_t*)quals1;
	a.b2 = (uint8_t*)bases2;
	a.q2 = (uint8_t*)quals2;
	a.len1 = len1;
	a.len2 = len2;
	a.stride = stride;
	std::string s1 = cfg->a1 ? cfg->a1 : "AGATCGGAAGAGCACACGTCTGAACTCCAGTCA"; // src/SeqPurge/main.cpp:25-26
	std::string s2 = cfg->a2 ? cfg->a2 : "AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGT";
	a.a1_len = (int)(s1.size() < 33 ? s1.size() : 33);
	a.a2_len = (int)(s2.size() < 33 ? s2.size() : 33);
	for (int i = 0; i < 33; ++i)
	{
		a.a1[i] = i < a.a1_len ? (uint8_t)s1[(size_t)i] : 'A';
		a.a2[i] = i < a.a2_len ? (uint8_t)s2[(size_t)i] : 'A';
	}
	a.cfg.a1 = a.cfg.a2 = nullptr;
	long long warps = n_pairs < 148LL * 64 ? n_pairs : 148LL * 64;
	int blocks = (int)((warps + 7) / 8);
	synth_kernel<<<blocks, 256, 0, (cudaStream_t)cuda_stream>>>(a);
	return cudaGetLastError() == cudaSuccess ? SPG_OK : SPG_ERR_CUDA;
}
