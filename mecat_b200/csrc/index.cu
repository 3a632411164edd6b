// mecat_b200/csrc/index.cu -- k = 13 direct-address index of an index volume (A1).
//
// Replaces create_ref_index + fill_ref_index_offsets_func (src/common/lookup_table.cpp:64-160,
// 26-61).  Same content: for every read, every 13-mer start; k-mers that occur more than 128
// times in the volume are dropped (:97); the start positions of a k-mer are stored in
// ascending order (every reference fill thread scans the reads in order, :36-58).
//
// Layout in HBM: CSR.  begin[2^26 + 1] (uint32) and pos[num_kmers] (int32), i.e. the
// reference's kmer_counts / kmer_starts / kmer_offsets triple without the pointer table.
//
// Kernels (all HBM bound): count (atomic histogram over the 2^26 codes), cutoff + exclusive
// scan, fill (atomic cursor per code), and a per-list register bitonic sort that restores the
// ascending order the atomics lost.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

#include <vector>

namespace mb {

namespace {

constexpr uint32_t CODE_MASK = (1u << (2 * KMER)) - 1;
constexpr int RUN = 16;      // consecutive k-mer starts handled by one thread

// Calls f(code, p) for every 13-mer start p of the read at `o` (offset, size).  A thread takes RUN
// consecutive starts: one 3-word window, the first code by a 2-bit-group reversal (first base most
// significant, lookup_table.cpp:79-90) and the next fifteen by rolling in one base each.
template <class F>
__device__ __forceinline__ void for_each_kmer(const uint32_t* __restrict__ fwd, const int2 o, F f)
{
	const int nk = o.y - (KMER - 1);
	for (int i0 = threadIdx.x * RUN; i0 < nk; i0 += blockDim.x * RUN) {
		const uint32_t p = (uint32_t)(o.x + i0);
		const uint32_t w = p >> 4, sh = (p & 15u) << 1;
		const uint32_t a0 = __ldg(fwd + w), a1 = __ldg(fwd + w + 1), a2 = __ldg(fwd + w + 2);
		const uint32_t w0 = __funnelshift_r(a0, a1, sh), w1 = __funnelshift_r(a1, a2, sh);
		uint32_t code = rev_groups2(w0) >> 6;
		f(code, p);
		if (nk - i0 >= RUN) {
#pragma unroll
			for (int j = 1; j < RUN; ++j) {
				const int b = KMER - 1 + j;
				const uint32_t base = (b < 16 ? w0 >> (2 * b) : w1 >> (2 * (b - 16))) & 3u;
				code = ((code << 2) & CODE_MASK) | base;
				f(code, p + j);
			}
		} else {
			const int n = nk - i0;
#pragma unroll
			for (int j = 1; j < RUN; ++j) {
				if (j >= n) break;
				const int b = KMER - 1 + j;
				const uint32_t base = (b < 16 ? w0 >> (2 * b) : w1 >> (2 * (b - 16))) & 3u;
				code = ((code << 2) & CODE_MASK) | base;
				f(code, p + j);
			}
		}
	}
}

// [code_lo, code_hi): the slice of the code space this launch is responsible for
__global__ void __launch_bounds__(256) k_kmer_count(const uint32_t* __restrict__ fwd, const int2* __restrict__ offsz, int nreads,
                                                    uint32_t* __restrict__ counts, uint32_t code_lo, uint32_t code_hi)
{
	const uint32_t span = code_hi - code_lo;
	for (int r = blockIdx.x; r < nreads; r += gridDim.x)
		for_each_kmer(fwd, offsz[r], [&](uint32_t code, uint32_t) {
			if (code - code_lo < span) atomicAdd(&counts[code], 1u);
		});
}

// cursor[code] starts at begin[code] for kept k-mers and at DROPPED for the others, so one atomic
// both tests the >128 cutoff and yields the slot.
constexpr uint32_t DROPPED = 0x80000000u;

__global__ void __launch_bounds__(256) k_kmer_fill(const uint32_t* __restrict__ fwd, const int2* __restrict__ offsz, int nreads,
                                                   uint32_t* __restrict__ cursor, int32_t* __restrict__ pos, uint32_t code_lo, uint32_t code_hi)
{
	const uint32_t span = code_hi - code_lo;
	for (int r = blockIdx.x; r < nreads; r += gridDim.x)
		for_each_kmer(fwd, offsz[r], [&](uint32_t code, uint32_t p) {
			if (code - code_lo < span) {
				const uint32_t slot = atomicAdd(&cursor[code], 1u);
				if (!(slot & DROPPED)) pos[slot] = (int32_t)p;
			}
		});
}

// ---- scatter by multi-split.  The fill above recomputes every k-mer of the volume once per code-range pass (32 passes)
// because a pass must keep its cursors and its slice of pos[] L2 resident.  The multi-split computes the k-mers ONCE:
// k_kmer_split bins (code, position) pairs into partitions of 2^19 codes -- each CTA ranks a tile of 4 096 k-mers in
// shared memory, reserves room in every partition with one global atomic per partition and tile, and writes its pairs
// there -- and k_pairs_scatter then walks the pairs partition by partition (coalesced), so that the cursors (2 MB) and
// the destination slice of pos[] (~50 MB) of the partitions in flight stay in L2.
constexpr int PART_BITS = 19;
constexpr uint32_t PART_CODES = 1u << PART_BITS;
constexpr int MAX_PARTS = (int)(NCODES >> PART_BITS);     // 128

__global__ void __launch_bounds__(256) k_part_totals(const uint32_t* __restrict__ counts, uint32_t code_lo, uint32_t code_hi,
                                                     uint32_t* __restrict__ totals)
{
	__shared__ uint32_t red[8];
	const uint32_t lo = code_lo + blockIdx.x * PART_CODES;
	const uint32_t hi = min(code_hi, lo + PART_CODES);
	uint32_t s = 0;
	for (uint32_t c = lo + threadIdx.x; c < hi; c += 256) s += counts[c];
	s = __reduce_add_sync(0xFFFFFFFFu, s);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
	__syncthreads();
	if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < 8; ++w) t += red[w]; totals[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(256) k_kmer_split(const uint32_t* __restrict__ fwd, const int2* __restrict__ offsz, int nreads,
                                                    uint32_t code_lo, uint32_t code_hi, int nparts, uint32_t* __restrict__ pcur,
                                                    uint2* __restrict__ pairs)
{
	__shared__ uint32_t hist[MAX_PARTS];       // pairs of this tile per partition
	__shared__ uint32_t offs[MAX_PARTS + 1];   // their exclusive prefix: the tile's pairs are staged partition by partition
	__shared__ uint32_t base[MAX_PARTS];       // where the partition's run starts in the global pair array
	__shared__ uint2 buf[256 * RUN];           // staged pairs: the write-out is then runs of consecutive pairs per partition
	const uint32_t span = code_hi - code_lo;
	for (int r = blockIdx.x; r < nreads; r += gridDim.x) {
		const int2 o = offsz[r];
		const int nk = o.y - (KMER - 1);
		for (int t0 = 0; t0 < nk; t0 += 256 * RUN) {              // one tile of 4 096 k-mer starts per iteration, all threads in step
			for (int q = threadIdx.x; q < MAX_PARTS; q += 256) hist[q] = 0;
			__syncthreads();
			const int i0 = t0 + threadIdx.x * RUN;
			uint32_t code[RUN], where[RUN];                           // where = partition << 16 | rank inside the CTA's tile
			if (i0 < nk) {
				const int n = min(RUN, nk - i0);
				const uint32_t p = (uint32_t)(o.x + i0);
				const uint32_t w = p >> 4, sh = (p & 15u) << 1;
				const uint32_t a0 = __ldg(fwd + w), a1 = __ldg(fwd + w + 1), a2 = __ldg(fwd + w + 2);
				const uint32_t w0 = __funnelshift_r(a0, a1, sh), w1 = __funnelshift_r(a1, a2, sh);
				uint32_t c = rev_groups2(w0) >> 6;
#pragma unroll
				for (int j = 0; j < RUN; ++j) {
					if (j > 0) {
						const int b = KMER - 1 + j;
						const uint32_t nb = (b < 16 ? w0 >> (2 * b) : w1 >> (2 * (b - 16))) & 3u;
						c = ((c << 2) & CODE_MASK) | nb;
					}
					code[j] = c;
					where[j] = 0xFFFFFFFFu;
					if (j < n && c - code_lo < span) {
						const uint32_t q = (c - code_lo) >> PART_BITS;
						where[j] = (q << 16) | atomicAdd(&hist[q], 1u);
					}
				}
			}
			__syncthreads();
			if (threadIdx.x < 32) {                                   // exclusive prefix of the 128 counters, one room reservation each
				const int l = threadIdx.x;
				const uint32_t h0 = hist[4 * l], h1 = hist[4 * l + 1], h2 = hist[4 * l + 2], h3 = hist[4 * l + 3];
				const uint32_t sum = h0 + h1 + h2 + h3;
				uint32_t inc = sum;
				for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (l >= d) inc += v; }
				const uint32_t ex = inc - sum;
				offs[4 * l] = ex; offs[4 * l + 1] = ex + h0; offs[4 * l + 2] = ex + h0 + h1; offs[4 * l + 3] = ex + h0 + h1 + h2;
				if (l == 31) offs[MAX_PARTS] = inc;
				if (h0) base[4 * l] = atomicAdd(&pcur[4 * l], h0);
				if (h1) base[4 * l + 1] = atomicAdd(&pcur[4 * l + 1], h1);
				if (h2) base[4 * l + 2] = atomicAdd(&pcur[4 * l + 2], h2);
				if (h3) base[4 * l + 3] = atomicAdd(&pcur[4 * l + 3], h3);
			}
			__syncthreads();
			if (i0 < nk) {
				const uint32_t p = (uint32_t)(o.x + i0);
#pragma unroll
				for (int j = 0; j < RUN; ++j)
					if (where[j] != 0xFFFFFFFFu) buf[offs[where[j] >> 16] + (where[j] & 0xFFFFu)] = make_uint2(code[j], p + (uint32_t)j);
			}
			__syncthreads();
			const uint32_t total = offs[MAX_PARTS];
			for (uint32_t e = threadIdx.x; e < total; e += 256) {
				const uint2 kp = buf[e];
				const uint32_t q = (kp.x - code_lo) >> PART_BITS;
				pairs[base[q] + (e - offs[q])] = kp;
			}
			__syncthreads();
		}
	}
}

__global__ void __launch_bounds__(256) k_pairs_scatter(const uint2* __restrict__ pairs, size_t n, uint32_t* __restrict__ cursor,
                                                       int32_t* __restrict__ pos)
{
	const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	const uint2 kp = pairs[i];
	const uint32_t slot = atomicAdd(&cursor[kp.x], 1u);
	if (!(slot & DROPPED)) pos[slot] = (int32_t)kp.y;
}

// ---- exclusive scan of min(count, cutoff -> 0) over 2^26 codes: reduce / top / down-sweep
constexpr int SCAN_T = 256, SCAN_E = 16, SCAN_TILE = SCAN_T * SCAN_E;   // 4096 codes per block

__device__ __forceinline__ uint32_t kept(uint32_t c) { return c > (uint32_t)MAX_OCC ? 0u : c; }

__global__ void __launch_bounds__(SCAN_T) k_scan_reduce(const uint32_t* __restrict__ counts, uint32_t* __restrict__ tile_sum)
{
	__shared__ uint32_t red[SCAN_T / 32];
	const size_t base = (size_t)blockIdx.x * SCAN_TILE;
	uint32_t s = 0;
	for (int e = 0; e < SCAN_E; ++e) s += kept(counts[base + (size_t)e * SCAN_T + threadIdx.x]);
	s = __reduce_add_sync(0xFFFFFFFFu, s);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t t = 0;
		for (int i = 0; i < SCAN_T / 32; ++i) t += red[i];
		tile_sum[blockIdx.x] = t;
	}
}

// single block: exclusive scan of ntiles (= 16384) tile sums in place; writes the grand total
__global__ void __launch_bounds__(1024) k_scan_top(uint32_t* __restrict__ tile_sum, int ntiles, uint32_t* __restrict__ total)
{
	__shared__ uint32_t part[1024];
	const int per = (ntiles + 1023) / 1024;
	const int lo = threadIdx.x * per, hi = min(ntiles, lo + per);
	uint32_t s = 0;
	for (int i = lo; i < hi; ++i) s += tile_sum[i];
	part[threadIdx.x] = s;
	__syncthreads();
	for (int off = 1; off < 1024; off <<= 1) {   // Hillis-Steele inclusive
		uint32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0u;
		__syncthreads();
		part[threadIdx.x] += v;
		__syncthreads();
	}
	uint32_t run = part[threadIdx.x] - s;
	for (int i = lo; i < hi; ++i) { uint32_t v = tile_sum[i]; tile_sum[i] = run; run += v; }
	if (threadIdx.x == 1023) *total = part[1023];
}

// writes begin[] and turns counts[] into the fill cursors
__global__ void __launch_bounds__(SCAN_T) k_scan_down(uint32_t* __restrict__ counts, const uint32_t* __restrict__ tile_sum,
                                                       uint32_t* __restrict__ begin)
{
	// thread t owns SCAN_E consecutive codes so that the running sum stays in registers
	__shared__ uint32_t wsum[SCAN_T / 32];
	const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_E;
	uint32_t v[SCAN_E];
	uint4* src = reinterpret_cast<uint4*>(counts + base);
#pragma unroll
	for (int e = 0; e < SCAN_E / 4; ++e) {
		uint4 q = src[e];
		v[4 * e] = kept(q.x); v[4 * e + 1] = kept(q.y); v[4 * e + 2] = kept(q.z); v[4 * e + 3] = kept(q.w);
	}
	uint32_t s = 0;
#pragma unroll
	for (int e = 0; e < SCAN_E; ++e) s += v[e];
	// exclusive scan of s across the block
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = s;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, off);
		if (lane >= off) inc += n;
	}
	if (lane == 31) wsum[warp] = inc;
	__syncthreads();
	uint32_t woff = 0;
	for (int w = 0; w < warp; ++w) woff += wsum[w];
	uint32_t run = tile_sum[blockIdx.x] + woff + inc - s;
	uint32_t o[SCAN_E];
#pragma unroll
	for (int e = 0; e < SCAN_E; ++e) { o[e] = run; run += v[e]; }
	uint4* dst = reinterpret_cast<uint4*>(begin + base);
#pragma unroll
	for (int e = 0; e < SCAN_E / 4; ++e) dst[e] = make_uint4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
#pragma unroll
	for (int e = 0; e < SCAN_E; ++e) if (v[e] == 0) o[e] = DROPPED;    // empty or over the cutoff: nothing is stored
#pragma unroll
	for (int e = 0; e < SCAN_E / 4; ++e) src[e] = make_uint4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
}

// ---- per-list ascending sort: one warp per k-mer list, R registers per lane (list <= 32*R)
template <int R>
__device__ __forceinline__ void warp_sort_list(int32_t* __restrict__ p, int n, int lane)
{
	int32_t v[R];
#pragma unroll
	for (int j = 0; j < R; ++j) { int i = j * 32 + lane; v[j] = i < n ? p[i] : 0x7fffffff; }
#pragma unroll
	for (int size = 2; size <= 32 * R; size <<= 1) {
#pragma unroll
		for (int stride = size >> 1; stride > 0; stride >>= 1) {
			if (stride >= 32) {
				const int rs = stride >> 5;
#pragma unroll
				for (int j = 0; j < R; ++j) {
					const int pj = j ^ rs;
					if (pj > j) {
						const int i = j * 32 + lane;
						const bool up = ((i & size) == 0) || size == 32 * R;
						int32_t a = v[j], b = v[pj];
						if ((a > b) == up) { v[j] = b; v[pj] = a; }
					}
				}
			} else {
#pragma unroll
				for (int j = 0; j < R; ++j) {
					const int i = j * 32 + lane;
					const bool up = ((i & size) == 0) || size == 32 * R;
					const int32_t o = __shfl_xor_sync(0xFFFFFFFFu, v[j], stride);
					const bool lower = (lane & stride) == 0;
					v[j] = (lower == up) ? min(v[j], o) : max(v[j], o);
				}
			}
		}
	}
#pragma unroll
	for (int j = 0; j < R; ++j) { int i = j * 32 + lane; if (i < n) p[i] = v[j]; }
}

constexpr int SORT_WARPS = 8, SORT_CODES_PER_WARP = 32;

__global__ void __launch_bounds__(SORT_WARPS * 32) k_sort_lists(const uint32_t* __restrict__ begin, int32_t* __restrict__ pos,
                                                                 uint32_t code_lo)
{
	const int lane = threadIdx.x & 31;
	const uint32_t c0 = code_lo + ((uint32_t)blockIdx.x * SORT_WARPS + (threadIdx.x >> 5)) * SORT_CODES_PER_WARP;
	const uint32_t b_lane = begin[c0 + lane];
	const uint32_t b_next = begin[c0 + lane + 1];
	for (int i = 0; i < SORT_CODES_PER_WARP; ++i) {
		const uint32_t b = __shfl_sync(0xFFFFFFFFu, b_lane, i);
		const int n = (int)(__shfl_sync(0xFFFFFFFFu, b_next, i) - b);
		if (n < 2) continue;
		if (n <= 32) warp_sort_list<1>(pos + b, n, lane);
		else if (n <= 64) warp_sort_list<2>(pos + b, n, lane);
		else warp_sort_list<4>(pos + b, n, lane);
	}
}

}  // namespace

// Number of code sub-ranges a histogram / scatter over [code_lo, code_hi) is split into; `full` is the
// measured optimum for the whole 2^26 code space (profiles/README.md), scaled for a slice of it.
static int index_passes(const char* env, int full, uint32_t code_lo, uint32_t code_hi)
{
	const char* e = getenv(env);      // tuning knob; the result does not depend on it
	if (e) full = atoi(e);
	const int n = (int)(((uint64_t)full * (code_hi - code_lo) + (1u << 25)) >> 26);
	return n < 1 ? 1 : (n > 256 ? 256 : n);
}

static int grid_for_reads(const DVolume* v) { return v->num_reads < 1 ? 1 : (v->num_reads > 65535 * 8 ? 65535 * 8 : v->num_reads); }

// Stage 1: histogram of the k-mers whose code lies in [code_lo, code_hi) (all codes for one GPU).
int index_count_part(Ctx* c, const DVolume* v, uint32_t code_lo, uint32_t code_hi, DIndex** out)
{
	if (code_lo > code_hi || code_hi > NCODES || (code_lo % (SORT_WARPS * SORT_CODES_PER_WARP)) || (code_hi % (SORT_WARPS * SORT_CODES_PER_WARP)))
		MB_FAIL(c, "index: code range must be aligned to %d", SORT_WARPS * SORT_CODES_PER_WARP);
	DIndex* I = new DIndex;
	auto body = [&]() -> int {
		MB_CUDA(c, c->alloc(&I->counts, (size_t)NCODES));
		MB_CUDA(c, c->alloc(&I->begin, (size_t)NCODES + 4));
		MB_CUDA(c, cudaMemsetAsync(I->counts, 0, sizeof(uint32_t) * (size_t)NCODES, c->stream));
		if (v->num_reads > 0 && code_hi > code_lo) {
			// The histogram is built in passes over sub-ranges of the codes: recomputing the codes costs
			// ~1 ms per pass, but the counters of one sub-range (4 * 2^26 / passes bytes) stay in the 126 MB
			// L2, so the random atomics stop being DRAM read-modify-writes of 32-byte sectors.
			const int passes = index_passes("MECAT_B200_COUNT_PASSES", 4, code_lo, code_hi);
			KScope ks(c, MECAT_K_COUNT, passes);
			for (int r = 0; r < passes; ++r) {
				const uint32_t lo = code_lo + (uint32_t)((uint64_t)(code_hi - code_lo) * r / passes);
				const uint32_t hi = code_lo + (uint32_t)((uint64_t)(code_hi - code_lo) * (r + 1) / passes);
				if (hi > lo) k_kmer_count<<<grid_for_reads(v), 256, 0, c->stream>>>(v->fwd, v->offsz, v->num_reads, I->counts, lo, hi);
			}
		}
		MB_CUDA(c, cudaGetLastError());
		MB_CUDA(c, cudaStreamSynchronize(c->stream));
		c->resolve_timers();
		return 0;
	};
	int rc = body();
	if (rc) { index_release(c, I); return rc; }
	*out = I;
	return 0;
}

// Stage 2 (after the counts of all code slices are in place): cutoff + scan over all codes, then
// positions and per-list order for [code_lo, code_hi).
int index_finish_part(Ctx* c, const DVolume* v, DIndex* I, uint32_t code_lo, uint32_t code_hi)
{
	uint32_t* d_tiles = nullptr;
	uint32_t* d_total = nullptr;
	uint32_t* d_pcur = nullptr;
	uint2* d_pairs = nullptr;
	const int ntiles = (int)(NCODES / SCAN_TILE);
	auto body = [&]() -> int {
		if (!I->counts) MB_FAIL(c, "index: finish called twice");
		MB_CUDA(c, c->alloc(&d_tiles, (size_t)ntiles));
		MB_CUDA(c, c->alloc(&d_total, 1));
		// multi-split scatter (default) or the code-range passes it replaced (MECAT_B200_FILL=passes)
		const char* fill_env = getenv("MECAT_B200_FILL");
		const bool split = !(fill_env && !strcmp(fill_env, "passes")) && code_hi > code_lo && v->num_reads > 0;
		const int nparts = split ? (int)((code_hi - code_lo + PART_CODES - 1) >> PART_BITS) : 0;
		std::vector<uint32_t> h_pbase((size_t)nparts + 1, 0);
		if (split) {
			// k-mers per partition of 2^19 codes, from the histogram before the scan turns it into cursors
			MB_CUDA(c, c->alloc(&d_pcur, (size_t)nparts));
			{
				KScope ks(c, MECAT_K_FILL);
				k_part_totals<<<nparts, 256, 0, c->stream>>>(I->counts, code_lo, code_hi, d_pcur);
			}
			MB_CUDA(c, cudaGetLastError());
			std::vector<uint32_t> tot((size_t)nparts);
			MB_CUDA(c, cudaMemcpyAsync(tot.data(), d_pcur, sizeof(uint32_t) * nparts, cudaMemcpyDeviceToHost, c->stream));
			MB_CUDA(c, cudaStreamSynchronize(c->stream));
			uint64_t run = 0;
			for (int q = 0; q < nparts; ++q) { h_pbase[q] = (uint32_t)run; run += tot[q]; }
			if (run > 0xFFFFFFFFull) MB_FAIL(c, "index: more than 2^32 k-mers in one volume");
			h_pbase[nparts] = (uint32_t)run;
			MB_CUDA(c, cudaMemcpyAsync(d_pcur, h_pbase.data(), sizeof(uint32_t) * nparts, cudaMemcpyHostToDevice, c->stream));
		}
		{
			KScope ks(c, MECAT_K_SCAN, 3);
			k_scan_reduce<<<ntiles, SCAN_T, 0, c->stream>>>(I->counts, d_tiles);
			k_scan_top<<<1, 1024, 0, c->stream>>>(d_tiles, ntiles, d_total);
			k_scan_down<<<ntiles, SCAN_T, 0, c->stream>>>(I->counts, d_tiles, I->begin);
		}
		MB_CUDA(c, cudaGetLastError());
		uint32_t total = 0;
		MB_CUDA(c, cudaMemcpyAsync(&total, d_total, sizeof total, cudaMemcpyDeviceToHost, c->stream));
		MB_CUDA(c, cudaStreamSynchronize(c->stream));
		MB_CUDA(c, cudaMemcpyAsync(I->begin + NCODES, &total, sizeof total, cudaMemcpyHostToDevice, c->stream));
		I->num_kmers = total;
		MB_CUDA(c, c->alloc(&I->pos, (size_t)total + 1));
		if (total && code_hi > code_lo && split) {
			const size_t npairs = h_pbase[nparts];
			MB_CUDA(c, c->alloc(&d_pairs, npairs + 1));
			{
				KScope ks(c, MECAT_K_FILL, 2);
				k_kmer_split<<<grid_for_reads(v), 256, 0, c->stream>>>(v->fwd, v->offsz, v->num_reads, code_lo, code_hi, nparts, d_pcur, d_pairs);
				if (npairs) k_pairs_scatter<<<(unsigned)((npairs + 255) / 256), 256, 0, c->stream>>>(d_pairs, npairs, I->counts, I->pos);
			}
			{
				KScope ks(c, MECAT_K_SORT);
				k_sort_lists<<<(code_hi - code_lo) / (SORT_WARPS * SORT_CODES_PER_WARP), SORT_WARPS * 32, 0, c->stream>>>(I->begin, I->pos, code_lo);
			}
			MB_CUDA(c, cudaGetLastError());
		} else if (total && code_hi > code_lo) {
			{
				// same idea for the scatter: per pass the cursors are L2 resident.  The destination slice of
				// pos[] (197 MB at 32 passes) is not, ncu still shows one 32-byte DRAM sector written per
				// position (profiles/r1_ncu_fill_pass.csv); more passes cost more re-scans than they save.
				const int passes = index_passes("MECAT_B200_FILL_PASSES", 32, code_lo, code_hi);
				KScope ks(c, MECAT_K_FILL, passes);
				for (int r = 0; r < passes; ++r) {
					const uint32_t lo = code_lo + (uint32_t)((uint64_t)(code_hi - code_lo) * r / passes);
					const uint32_t hi = code_lo + (uint32_t)((uint64_t)(code_hi - code_lo) * (r + 1) / passes);
					if (hi > lo) k_kmer_fill<<<grid_for_reads(v), 256, 0, c->stream>>>(v->fwd, v->offsz, v->num_reads, I->counts, I->pos, lo, hi);
				}
			}
			{
				KScope ks(c, MECAT_K_SORT);
				k_sort_lists<<<(code_hi - code_lo) / (SORT_WARPS * SORT_CODES_PER_WARP), SORT_WARPS * 32, 0, c->stream>>>(I->begin, I->pos, code_lo);
			}
			MB_CUDA(c, cudaGetLastError());
		}
		MB_CUDA(c, cudaStreamSynchronize(c->stream));
		c->resolve_timers();
		c->stats.index_kmers += total;
		c->stats.index_bases += v->num_bases;
		return 0;
	};
	int rc = body();
	c->dfree(d_tiles); c->dfree(d_total); c->dfree(d_pcur); c->dfree(d_pairs);
	c->dfree(I->counts); I->counts = nullptr;
	return rc;
}

int index_build(Ctx* c, const DVolume* v, DIndex** out)
{
	DIndex* I = nullptr;
	int rc = index_count_part(c, v, 0, NCODES, &I);
	if (!rc) rc = index_finish_part(c, v, I, 0, NCODES);
	if (rc) { index_release(c, I); return rc; }
	*out = I;
	return 0;
}

void index_release(Ctx* c, DIndex* i)
{
	if (!i) return;
	c->dfree(i->begin);
	c->dfree(i->pos);
	c->dfree(i->counts);
	delete i;
}

}  // namespace mb
