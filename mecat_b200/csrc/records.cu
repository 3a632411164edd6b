// mecat_b200/csrc/records.cu -- record assembly (row A12) and result text on the device.
//
// CUDA backend of m4_core.cuh.  After the extension kernel the results of a tile stay in device memory:
//   k_m4_fill    a thread per candidate: fill_m4record (pw_impl.cpp:467-506) + the sort key of its record
//   k_m4_order   a thread per query read: the library's std::sort on (key, candidate) items and the containment
//                filter of append_m4v (pw_impl.cpp:550-610); the kept candidates in order, and their number
//   scan         offsets of the reads' records
//   k_m4_gather  a warp per query read: the kept records, 104 bytes each, copied word by word to their final place
// and, for the command-line drivers, the lines of the result file are written there too:
//   k_text_len / scan / k_text_write   a thread per record: alignment.cpp:18-32,58-78
// so that a tile returns either packed records or finished text, and the host threads that used to sort, filter and
// print behind every extension chunk are no longer part of the path.
#include "common.cuh"
#include "m4_core.cuh"

namespace mb {

namespace {

__global__ void k_m4_fill(const ExtendTask* __restrict__ tasks, const mecat_extend_result* __restrict__ res,
                          const int32_t* __restrict__ scores, const int64_t* __restrict__ outpos, int nreads,
                          const int2* __restrict__ qoffsz, int qstart_id, const int2* __restrict__ soffsz, int sstart_id,
                          mecat_m4* __restrict__ tmp, mbm4::SortItem* __restrict__ items)
{
	const int r = blockIdx.x;
	if (r >= nreads) return;
	const int64_t k0 = outpos[r], k1 = outpos[r + 1];
	const int64_t qsize = qoffsz[r].y, qid = (int64_t)r + qstart_id;
	for (int64_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
		const mecat_extend_result R = res[k];
		mbm4::SortItem it;
		it.idx = (int32_t)(k - k0); it.pad = 0; it.key = ~0ull;           // not ok: never looked at
		if (R.ok) {
			const ExtendTask t = tasks[k];
			const mecat_m4 m = mbm4::make_m4((int64_t)t.sread + sstart_id, soffsz[t.sread].y, qid, qsize, t.qstrand, t.qstart, t.sstart, R, scores[k]);
			tmp[k] = m;
			it.key = mbm4::m4_key(m.qid, m.qend - m.qoff, m.send - m.soff);
		}
		items[k] = it;
	}
}

__global__ void k_m4_order(const mecat_extend_result* __restrict__ res, const mecat_m4* __restrict__ tmp,
                           const int64_t* __restrict__ outpos, int nreads, mbm4::SortItem* __restrict__ items,
                           int32_t* __restrict__ order, int32_t* __restrict__ nkept)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nreads) return;
	const int64_t k0 = outpos[r], k1 = outpos[r + 1];
	mbm4::SortItem* a = items + k0;
	// the accepted records in candidate order (the order pairwise_mapping appends them in, pw_impl.cpp:688-700)
	int n = 0;
	for (int64_t k = k0; k < k1; ++k) if (res[k].ok) a[n++] = items[k];
	mbm4::std_sort(a, n);
	// check_records_containment inside every run of equal qid; a dropped record is marked in its pad field
	const mecat_m4* m = tmp + k0;
	for (int i = 0; i < n;) {
		int j = i + 1;
		const int64_t qid = m[a[i].idx].qid;
		while (j < n && m[a[j].idx].qid == qid) ++j;
		for (int x = i; x < j; ++x) {
			if (a[x].pad) continue;
			const mecat_m4 A = m[a[x].idx];
			for (int y = x + 1; y < j; ++y)
				if (!a[y].pad && mbm4::m4_contained(A, m[a[y].idx])) a[y].pad = 1;
		}
		i = j;
	}
	int kept = 0;
	for (int i = 0; i < n; ++i) if (!a[i].pad) order[k0 + kept++] = a[i].idx;
	nkept[r] = kept;
}

__global__ void k_m4_gather(const mecat_m4* __restrict__ tmp, const int64_t* __restrict__ outpos, const int32_t* __restrict__ order,
                            const int32_t* __restrict__ nkept, const int64_t* __restrict__ dst, int nreads, mecat_m4* __restrict__ out)
{
	const int r = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (r >= nreads) return;
	const int64_t k0 = outpos[r];
	const int n = nkept[r];
	constexpr int W = (int)(sizeof(mecat_m4) / 4);      // 26 words
	for (int j = 0; j < n; ++j) {
		const uint32_t* s = (const uint32_t*)(tmp + k0 + order[k0 + j]);
		uint32_t* d = (uint32_t*)(out + dst[r] + j);
		if (lane < W) d[lane] = s[lane];
	}
}

template <int KIND>
__device__ __forceinline__ int line_of(char* b, const void* recs, size_t i, bool gapped)
{
	if (KIND == 0) return mbm4::line_candidate(b, ((const mecat_candidate*)recs)[i]);
	return mbm4::line_m4(b, ((const mecat_m4*)recs)[i], gapped);
}

template <int KIND>
__global__ void k_text_len(const void* __restrict__ recs, size_t n, int gapped, int32_t* __restrict__ len, int* __restrict__ bad)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	char b[mbm4::LINE_CAP];
	int l = line_of<KIND>(b, recs, i, gapped != 0);
	if (l < 0) { atomicExch(bad, 1); l = 0; }
	len[i] = l;
}

template <int KIND>
__global__ void k_text_write(const void* __restrict__ recs, size_t n, int gapped, const int64_t* __restrict__ off, char* __restrict__ text)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	char b[mbm4::LINE_CAP];
	const int l = line_of<KIND>(b, recs, i, gapped != 0);
	char* d = text + off[i];
	for (int k = 0; k < l; ++k) d[k] = b[k];
}

}  // namespace

int m4_assemble(Ctx* c, const DVolume* reads, const DVolume* ref, const ExtendTask* d_tasks, const mecat_extend_result* d_res,
                const int32_t* d_scores, const int64_t* d_outpos, int nreads, size_t total, mecat_m4** d_m4, size_t* nout)
{
	*d_m4 = nullptr; *nout = 0;
	if (!total || nreads <= 0) return 0;
	mecat_m4* d_tmp = nullptr;
	mbm4::SortItem* d_items = nullptr;
	int32_t *d_order = nullptr, *d_nkept = nullptr;
	int64_t* d_dst = nullptr;
	mecat_m4* d_out = nullptr;
	auto body = [&]() -> int {
		MB_CUDA(c, c->alloc(&d_tmp, total));
		MB_CUDA(c, c->alloc(&d_items, total));
		MB_CUDA(c, c->alloc(&d_order, total));
		MB_CUDA(c, c->alloc(&d_nkept, (size_t)nreads));
		MB_CUDA(c, c->alloc(&d_dst, (size_t)nreads + 1));
		{
			KScope ks(c, MECAT_K_FINAL, 2);
			k_m4_fill<<<nreads, 32, 0, c->stream>>>(d_tasks, d_res, d_scores, d_outpos, nreads, reads->offsz, reads->start_read_id, ref->offsz,
			                                       ref->start_read_id, d_tmp, d_items);
			k_m4_order<<<(nreads + 63) / 64, 64, 0, c->stream>>>(d_res, d_tmp, d_outpos, nreads, d_items, d_order, d_nkept);
		}
		MB_CUDA(c, cudaGetLastError());
		int64_t kept = 0;
		if (device_exclusive_scan(c, d_nkept, d_dst, nreads, &kept)) return 1;
		if (kept) {
			MB_CUDA(c, c->alloc(&d_out, (size_t)kept));
			KScope ks(c, MECAT_K_FINAL);
			k_m4_gather<<<(unsigned)(((size_t)nreads * 32 + 127) / 128), 128, 0, c->stream>>>(d_tmp, d_outpos, d_order, d_nkept, d_dst, nreads, d_out);
			MB_CUDA(c, cudaGetLastError());
		}
		*nout = (size_t)kept;
		return 0;
	};
	const int rc = body();
	c->dfree(d_tmp); c->dfree(d_items); c->dfree(d_order); c->dfree(d_nkept); c->dfree(d_dst);
	if (rc) { c->dfree(d_out); return rc; }
	*d_m4 = d_out;
	return 0;
}

int records_text_device(Ctx* c, int kind, int gapped, const void* d_records, size_t n, char** d_text, size_t* bytes)
{
	*d_text = nullptr; *bytes = 0;
	if (!n) return 0;
	int32_t* d_len = nullptr;
	int64_t* d_off = nullptr;
	char* d_out = nullptr;
	int* d_bad = (int*)(c->d_counters + 8);
	auto body = [&]() -> int {
		MB_CUDA(c, c->alloc(&d_len, n));
		MB_CUDA(c, c->alloc(&d_off, n + 1));
		MB_CUDA(c, cudaMemsetAsync(d_bad, 0, sizeof(int), c->stream));
		const unsigned grid = (unsigned)((n + 127) / 128);
		{
			KScope ks(c, MECAT_K_FINAL);
			if (kind == 0) k_text_len<0><<<grid, 128, 0, c->stream>>>(d_records, n, gapped, d_len, d_bad);
			else k_text_len<1><<<grid, 128, 0, c->stream>>>(d_records, n, gapped, d_len, d_bad);
		}
		MB_CUDA(c, cudaGetLastError());
		int64_t total = 0;
		if (device_exclusive_scan(c, d_len, d_off, (int64_t)n, &total)) return 1;
		int bad = 0;
		MB_CUDA(c, cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
		MB_CUDA(c, cudaStreamSynchronize(c->stream));
		if (bad) MB_FAIL(c, "records_text: an identity outside [0, 1e6) cannot be printed on the device");
		MB_CUDA(c, c->dmalloc((void**)&d_out, (size_t)total + 16));
		{
			KScope ks(c, MECAT_K_FINAL);
			if (kind == 0) k_text_write<0><<<grid, 128, 0, c->stream>>>(d_records, n, gapped, d_off, d_out);
			else k_text_write<1><<<grid, 128, 0, c->stream>>>(d_records, n, gapped, d_off, d_out);
		}
		MB_CUDA(c, cudaGetLastError());
		*bytes = (size_t)total;
		return 0;
	};
	const int rc = body();
	c->dfree(d_len); c->dfree(d_off);
	if (rc) { c->dfree(d_out); return rc; }
	*d_text = d_out;
	return 0;
}

}  // namespace mb
