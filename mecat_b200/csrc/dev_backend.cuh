// mecat_b200/csrc/dev_backend.cuh -- what the CUDA backends of the stage pipelines share (cns.cu for cns_pipeline.h,
// refmap.cu for ref_pipeline.h): device memory from the context's pool with an ownership list, copies on the context's
// stream with the traffic counters, error text into the context.
#pragma once
#include "common.cuh"

namespace mb {

struct PoolBackend
{
	Ctx* c;
	const char* tag;               // prefix of the error texts
	bool sync_uploads;             // true when the host sources of upload() may go out of scope right after the call
	std::vector<void*> owned;

	PoolBackend(Ctx* c_, const char* tag_, bool sync_uploads_) : c(c_), tag(tag_), sync_uploads(sync_uploads_) {}

	bool check(cudaError_t e, const char* what)
	{
		if (e == cudaSuccess) return true;
		char b[256];
		snprintf(b, sizeof b, "%s: %s: %s", tag, what, cudaGetErrorString(e));
		c->err = b;
		return false;
	}
	template <class T> T* alloc(size_t n)
	{
		void* p = nullptr;
		const cudaError_t e = c->dmalloc(&p, (n ? n : 1) * sizeof(T));
		if (e != cudaSuccess) {
			char b[256];
			snprintf(b, sizeof b, "%s: device allocation of %zu bytes failed: %s", tag, n * sizeof(T), cudaGetErrorString(e));
			c->err = b;
			return nullptr;
		}
		owned.push_back(p);
		return (T*)p;
	}
	template <class T> bool upload(T* d, const T* h, size_t n)
	{
		if (!n) return true;
		c->stats.h2d_bytes += (int64_t)(n * sizeof(T));
		if (!check(cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, c->stream), "H2D")) return false;
		return !sync_uploads || check(cudaStreamSynchronize(c->stream), "H2D");
	}
	template <class T> bool download(T* h, const T* d, size_t n)
	{
		if (n && !check(cudaMemcpyAsync(h, d, n * sizeof(T), cudaMemcpyDeviceToHost, c->stream), "D2H")) return false;
		c->stats.d2h_bytes += (int64_t)(n * sizeof(T));
		return check(cudaStreamSynchronize(c->stream), "kernel");
	}
	bool fill(void* d, int byte, size_t bytes) { return !bytes || check(cudaMemsetAsync(d, byte, bytes, c->stream), "memset"); }
	bool release(void* p)      // the pool only marks the block free; work queued on the stream before the next owner's is ordered
	{
		for (size_t i = 0; i < owned.size(); ++i)
			if (owned[i] == p) { c->dfree(p); owned[i] = owned.back(); owned.pop_back(); return true; }
		c->err = std::string(tag) + ": release of an unknown block";
		return false;
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

}  // namespace mb
