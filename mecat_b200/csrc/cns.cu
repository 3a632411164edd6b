// mecat_b200/csrc/cns.cu -- consensus stage of mecat2cns on the GPU (rows C3-C7).
//
// The stage sequence lives in cns_pipeline.h, the per-unit bodies in cns_core.cuh; this file is the CUDA
// backend: every stage functor F becomes a launch of k_cns<F> (one thread per read / accepted alignment /
// region / segment), the arena sizes between stages come from a device scan, memory comes from the context's
// pool.  Inputs are the extension results exactly where align.cu left them in device memory; only the corrected
// bases and a few counters per segment cross to the host.
#include "common.cuh"
#include "cns_pipeline.h"

namespace mb {

namespace {

template <class F>
__global__ void __launch_bounds__(128) k_cns(const F f, const int64_t n)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) f(i);
}

struct WarpLanes
{
	static constexpr int count = 32;
	__device__ int lane() const { return (int)(threadIdx.x & 31u); }
	__device__ int sum(int v) const { return __reduce_add_sync(0xffffffffu, v); }
	__device__ void sync() const { __syncwarp(); }
};

template <class F>
__global__ void __launch_bounds__(128) k_cns_warp(const F f, const int64_t n)
{
	const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (i < n) f(i, WarpLanes());
}

// exclusive prefix sum of n int32 values into n + 1 int64 values; one CTA, each thread owns a contiguous chunk
__global__ void __launch_bounds__(1024) k_cns_scan(const int32_t* __restrict__ in, int64_t* __restrict__ out, const int64_t n)
{
	__shared__ int64_t part[1024];
	const int tid = threadIdx.x;
	const int64_t chunk = (n + 1023) / 1024;
	const int64_t b = min(n, (int64_t)tid * chunk), e = min(n, b + chunk);
	int64_t s = 0;
	for (int64_t i = b; i < e; ++i) s += in[i];
	part[tid] = s;
	__syncthreads();
	for (int d = 1; d < 1024; d <<= 1) {
		const int64_t v = tid >= d ? part[tid - d] : 0;
		__syncthreads();
		part[tid] += v;
		__syncthreads();
	}
	int64_t run = part[tid] - s;
	for (int64_t i = b; i < e; ++i) { out[i] = run; run += in[i]; }
	if (tid == 1023) out[n] = part[1023];
}

struct DevBackend
{
	Ctx* c;
	std::vector<void*> owned;

	template <class T> T* alloc(size_t n)
	{
		void* p = nullptr;
		const cudaError_t e = c->dmalloc(&p, (n ? n : 1) * sizeof(T));
		if (e != cudaSuccess) {
			char b[256];
			snprintf(b, sizeof b, "cns: device allocation of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e));
			c->err = b;
			return nullptr;
		}
		owned.push_back(p);
		return (T*)p;
	}
	bool check(cudaError_t e, const char* what)
	{
		if (e == cudaSuccess) return true;
		char b[256];
		snprintf(b, sizeof b, "cns: %s: %s", what, cudaGetErrorString(e));
		c->err = b;
		return false;
	}
	template <class T> bool upload(T* d, const T* h, size_t n)
	{
		if (!n) return true;
		c->stats.h2d_bytes += (int64_t)(n * sizeof(T));
		return check(cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, c->stream), "H2D");
	}
	template <class T> bool download(T* h, const T* d, size_t n)
	{
		if (n && !check(cudaMemcpyAsync(h, d, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream), "D2H")) return false;
		c->stats.d2h_bytes += (int64_t)(n * sizeof(T));
		return check(cudaStreamSynchronize(c->stream), "kernel");
	}
	bool fill(void* d, int byte, size_t bytes) { return !bytes || check(cudaMemsetAsync(d, byte, bytes, c->stream), "memset"); }
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
	bool scan(const int32_t* in, int64_t* out, int64_t n, int64_t* total)
	{
		{
			KScope ks(c, MECAT_K_SCAN);
			k_cns_scan<<<1, 1024, 0, c->stream>>>(in, out, n);
		}
		if (!check(cudaGetLastError(), "launch")) return false;
		return download(total, out + n, 1);
	}
	void fail(const char* m) { c->err = m; }
	void end_batch()
	{
		cudaStreamSynchronize(c->stream);
		for (void* p : owned) c->dfree(p);
		owned.clear();
		c->resolve_timers();
	}
};

}  // namespace

int cns_consensus_device(Ctx* c, const mbcns::BatchIn& in, const mbcns::Params& P, std::vector<mbcns::Piece>& out)
{
	DevBackend be{c, {}};
	return mbcns::consensus_batch(be, in, P, out);
}

}  // namespace mb
