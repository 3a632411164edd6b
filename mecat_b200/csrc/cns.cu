// mecat_b200/csrc/cns.cu -- consensus stage of mecat2cns on the GPU (rows C3-C7).
//
// The stage sequence lives in cns_pipeline.h, the per-unit bodies in cns_core.cuh; this file is the CUDA
// backend: every stage functor F becomes a launch of k_cns<F> (one thread per read / accepted alignment /
// region / segment), the arena sizes between stages come from a device scan, memory comes from the context's
// pool.  Inputs are the extension results exactly where align.cu left them in device memory; only the corrected
// bases and a few counters per segment cross to the host.
#include "common.cuh"
#include "dev_backend.cuh"
#include "cns_pipeline.h"

namespace mb {

namespace {

template <class F>
__global__ void __launch_bounds__(128) k_cns(const F f, const int64_t n)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) f(i);
}

struct WarpLanes          // the lanes interface of cns_core.cuh on a real warp
{
	static constexpr int count = 32;
	template <class F> __device__ void each(F&& f) const { f((int)(threadIdx.x & 31u)); __syncwarp(); }
	template <class F> __device__ int sum(F&& f) const { return __reduce_add_sync(0xffffffffu, f((int)(threadIdx.x & 31u))); }
	template <class F> __device__ uint32_t ballot(F&& f) const { return __ballot_sync(0xffffffffu, f((int)(threadIdx.x & 31u))); }
	template <class F> __device__ void ballot2(F&& f, uint32_t& m0, uint32_t& m1) const
	{
		const int v = f((int)(threadIdx.x & 31u));
		m0 = __ballot_sync(0xffffffffu, v & 1);
		m1 = __ballot_sync(0xffffffffu, v & 2);
	}
	__device__ bool leader() const { return (threadIdx.x & 31u) == 0; }
	__device__ void sync() const { __syncwarp(); }
	__device__ void* scratch() const { return buf; }         // mbcns::LANES_SCRATCH bytes of shared memory private to the warp
	void* buf;
};

template <class F>
__global__ void __launch_bounds__(128) k_cns_warp(const F f, const int64_t n)
{
	__shared__ uint64_t scratch[4][mbcns::LANES_SCRATCH / 8];      // 4 warps per CTA (launch_warp)
	const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	WarpLanes lanes;
	lanes.buf = scratch[threadIdx.x >> 5];
	if (i < n) f(i, lanes);
}

// The region graphs of one wave, a thread per region.  A graph is pointer chasing over a few hundred bytes of nodes
// and edges; in global memory every step of it is an L2 / DRAM round trip (ncu: 14 KB of DRAM traffic per region for
// a 2 KB arena, long-scoreboard bound).  So the CTA owns a pool of shared memory: each thread asks for the bytes its
// graph needs with the narrowest index type that holds it (8 bits for most), a CTA-wide scan hands out slices, and the threads whose slice fits build their
// graph there; the rest wait for the next round of the same pool.  Only graphs too large for the pool (or for 16-bit
// indices) use their exact-size arena in global memory.
constexpr int POA_BLOCK = 192;
constexpr int POA_POOL = 108 * 1024;       // two CTAs per SM; ~0.5 KB per graph with 8-bit indices

__global__ void __launch_bounds__(POA_BLOCK) k_cns_poa(const mbcns::PoaFn f, const int64_t n)
{
	extern __shared__ __align__(16) char pool[];
	__shared__ int wsum[POA_BLOCK / 32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int64_t k = (int64_t)blockIdx.x * POA_BLOCK + tid;
	bool pending = k < n;
	int need = 0, width = 4;
	if (pending) {
		int ncap, e0;
		f.shape(k, ncap, e0);
		width = mbcns::PoaFn::width(ncap, e0);
		const int64_t b = mbcns::PoaFn::bytes_for(width, ncap, e0);
		if (b > POA_POOL) {                  // too large for the pool: its own global arena
			f.solve_width(width, k, f.wide_arena(k));
			pending = false;
		} else need = (int)b;
	}
	while (__syncthreads_or(pending)) {
		// exclusive scan of the pending requests over the CTA
		const int v = pending ? need : 0;
		int inc = v;
		for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
		if (lane == 31) wsum[warp] = inc;
		__syncthreads();
		int off = inc - v;
		for (int w = 0; w < warp; ++w) off += wsum[w];
		if (pending && off + need <= POA_POOL) {
			f.solve_width(width, k, pool + off);
			pending = false;
		}
		__syncthreads();                     // the pool and wsum are reused by the next round
	}
}

// Exclusive prefix sum of n int32 values into n + 1 int64 values, three launches: totals of 4 096-element tiles
// (coalesced), a one-CTA scan of the tile totals, and the tile-local scan with its tile offset added.
constexpr int SCAN_TILE = 4096;      // 1 024 threads x 4 consecutive values

__device__ __forceinline__ int64_t block_exclusive(int64_t v, int64_t* total)     // 1 024 threads
{
	__shared__ int64_t wsum[32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int64_t inc = v;
	for (int d = 1; d < 32; d <<= 1) { const int64_t o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
	if (lane == 31) wsum[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		int64_t w = wsum[lane], winc = w;
		for (int d = 1; d < 32; d <<= 1) { const int64_t o = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += o; }
		wsum[lane] = winc - w;
		if (lane == 31) *total = winc;
	}
	__syncthreads();
	return wsum[warp] + inc - v;
}

__global__ void __launch_bounds__(1024) k_cns_scan_tiles(const int32_t* __restrict__ in, int64_t* __restrict__ tile_sum, const int64_t n)
{
	__shared__ int64_t total;
	const int64_t i0 = (int64_t)blockIdx.x * SCAN_TILE + 4 * threadIdx.x;
	int64_t s = 0;
	for (int k = 0; k < 4; ++k) if (i0 + k < n) s += in[i0 + k];
	block_exclusive(s, &total);
	__syncthreads();
	if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_cns_scan_top(int64_t* __restrict__ tile_sum, const int64_t ntiles, int64_t* __restrict__ grand)
{
	__shared__ int64_t total;
	int64_t carry = 0;
	for (int64_t base = 0; base < ntiles; base += 1024) {
		const int64_t i = base + threadIdx.x;
		const int64_t v = i < ntiles ? tile_sum[i] : 0;
		const int64_t ex = block_exclusive(v, &total);
		if (i < ntiles) tile_sum[i] = carry + ex;
		__syncthreads();
		carry += total;
		__syncthreads();
	}
	if (threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(1024) k_cns_scan_apply(const int32_t* __restrict__ in, const int64_t* __restrict__ tile_off,
                                                         int64_t* __restrict__ out, const int64_t n)
{
	__shared__ int64_t total;
	const int64_t i0 = (int64_t)blockIdx.x * SCAN_TILE + 4 * threadIdx.x;
	int32_t v[4];
	int64_t s = 0;
	for (int k = 0; k < 4; ++k) { v[k] = i0 + k < n ? in[i0 + k] : 0; s += v[k]; }
	int64_t run = tile_off[blockIdx.x] + block_exclusive(s, &total);
	for (int k = 0; k < 4; ++k) if (i0 + k < n) { out[i0 + k] = run; run += v[k]; }
}

struct DevBackend : PoolBackend
{
	explicit DevBackend(Ctx* ctx) : PoolBackend(ctx, "cns", false) {}

	const char* download_staged(const char* d, size_t n)      // through the context's pinned staging buffer
	{
		void* h = nullptr;
		if (!check(c->host_stage(3, n + 1, &h), "pinned staging")) return nullptr;
		if (n && !check(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, c->stream), "D2H")) return nullptr;
		c->stats.d2h_bytes += (int64_t)n;
		if (!check(cudaStreamSynchronize(c->stream), "kernel")) return nullptr;
		return (const char*)h;
	}
	template <class F> bool launch(int64_t n, const F& f, int stage)
	{
		if (n <= 0) return true;
		KScope ks(c, MECAT_K_CNS_ACCEPT + stage);
		k_cns<F><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(f, n);
		return check(cudaGetLastError(), "launch");
	}
	template <class F> bool launch_warp(int64_t n, const F& f, int stage)
	{
		if (n <= 0) return true;
		KScope ks(c, MECAT_K_CNS_ACCEPT + stage);
		k_cns_warp<F><<<(unsigned)((n + 3) / 4), 128, 0, c->stream>>>(f, n);
		return check(cudaGetLastError(), "launch");
	}
	bool launch_graphs(int64_t n, const mbcns::PoaFn& f, int stage)
	{
		if (n <= 0) return true;
		// per device, so per call (one host thread per GPU shares this code)
		if (!check(cudaFuncSetAttribute(k_cns_poa, cudaFuncAttributeMaxDynamicSharedMemorySize, POA_POOL), "shared memory opt-in")) return false;
		KScope ks(c, MECAT_K_CNS_ACCEPT + stage);
		k_cns_poa<<<(unsigned)((n + POA_BLOCK - 1) / POA_BLOCK), POA_BLOCK, POA_POOL, c->stream>>>(f, n);
		return check(cudaGetLastError(), "launch");
	}
	bool scan(const int32_t* in, int64_t* out, int64_t n, int64_t* total)
	{
		if (n <= 0) { *total = 0; const int64_t zero = 0; return upload(out, &zero, 1); }
		const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
		int64_t* tiles = alloc<int64_t>((size_t)ntiles);
		if (!tiles) return false;
		{
			KScope ks(c, MECAT_K_SCAN, 3);
			k_cns_scan_tiles<<<(unsigned)ntiles, 1024, 0, c->stream>>>(in, tiles, n);
			k_cns_scan_top<<<1, 1024, 0, c->stream>>>(tiles, ntiles, out + n);
			k_cns_scan_apply<<<(unsigned)ntiles, 1024, 0, c->stream>>>(in, tiles, out, n);
		}
		if (!check(cudaGetLastError(), "launch")) return false;
		return download(total, out + n, 1);
	}
	int64_t poa_budget_bytes() const
	{
		if (const char* e = getenv("MECAT_B200_POA_BUDGET_MB")) return (int64_t)atoll(e) << 20;      // test hook: force several waves
		return 24ll << 30;
	}
};

}  // namespace

// exclusive prefix sum of n int32 counts into n + 1 int64 offsets (all on the device), the total also on the host
int device_exclusive_scan(Ctx* c, const int32_t* d_in, int64_t* d_out, int64_t n, int64_t* h_total)
{
	DevBackend be(c);
	const bool ok = be.scan(d_in, d_out, n, h_total);
	be.end_batch();
	return ok ? 0 : 1;
}

int cns_consensus_device(Ctx* c, const mbcns::BatchIn& in, const mbcns::Params& P, CnsBlob& out)
{
	DevBackend be(c);
	return mbcns::consensus_batch(be, in, P, out);
}

}  // namespace mb
