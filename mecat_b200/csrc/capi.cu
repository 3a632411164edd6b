// mecat_b200/csrc/capi.cu -- the C ABI of include/mecat_b200.h.
//
// Host-side glue only: argument checks, device buffers, launches, and the per-read record
// assembly that the reference does on the CPU after its hot loops
// (candidate_detect pw_impl.cpp:767-793, fill_m4record :467-506, append_m4v :576-610).
#include "common.cuh"
#include "cns_pipeline.h"
#include "ref_pipeline.h"

#include <algorithm>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <chrono>
#include <thread>
#include <cstdlib>
#include <cstring>

using namespace mb;

struct mecat_b200_ctx : public mb::Ctx {};

namespace {

struct WallTimer   // host clock, for the copies and the host-side record assembly
{
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	float stop() const { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

__global__ void k_extend_finalize(const ExtendTask* __restrict__ tasks, const ExtendHalf* __restrict__ halves, size_t n,
                                  int min_aln, mecat_extend_result* __restrict__ out)
{
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if (i >= n) return;
	const ExtendTask t = tasks[i];
	const ExtendHalf L = halves[2 * i], R = halves[2 * i + 1];
	mecat_extend_result r;
	r.columns = L.cols + R.cols;
	r.matches = L.matches + R.matches;
	r.qstart = t.qstart - L.qadv; r.qend = t.qstart + R.qadv;
	r.sstart = t.sstart - L.tadv; r.send = t.sstart + R.tadv;
	r.ok = r.columns >= min_aln;
	r.pad_ = 0;
	// OutputStore::calc_ident, diff_gapalign.h:90-97: 100.0 * n / size in IEEE double
	r.ident = r.columns ? __ddiv_rn(__dmul_rn(100.0, (double)r.matches), (double)r.columns) : 0.0;
	out[i] = r;
}

// ExtensionCandidate assembly: candidate_detect, pw_impl.cpp:767-793
__global__ void k_make_ec(const RawCand* __restrict__ cands, const int32_t* __restrict__ counts,
                          const int64_t* __restrict__ outpos, int maxc, int nreads, const int2* __restrict__ qoffsz,
                          int qstart_id, const int2* __restrict__ soffsz, int sstart_id, mecat_candidate* __restrict__ ec,
                          ExtendTask* __restrict__ tasks, int32_t* __restrict__ scores)
{
	int r = blockIdx.x;
	if (r >= nreads) return;
	const int n = counts[r];
	const int64_t base = outpos[r];
	const int qsize = qoffsz[r].y;
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		const RawCand c = cands[(size_t)r * maxc + i];
		int qstart = c.loc2, sstart = c.loc1;
		if (qstart && sstart) { qstart += KMER / 2; sstart += KMER / 2; }
		const int sidx = c.readno - sstart_id;
		if (ec) {
			mecat_candidate e;
			e.qdir = c.chain; e.qid = r + qstart_id; e.qext = qstart; e.qsize = qsize; e.qoff = 0; e.qend = 0;
			e.sdir = 0; e.sid = c.readno; e.sext = sstart; e.ssize = soffsz[sidx].y; e.soff = 0; e.send = 0;
			e.score = c.score;
			if (e.qdir == 1) e.qext = e.qsize - 1 - e.qext;
			ec[base + i] = e;
		}
		if (tasks) {
			ExtendTask t;
			t.qread = r; t.qstrand = c.chain; t.qstart = qstart; t.sread = sidx; t.sstart = sstart;
			tasks[base + i] = t;
			scores[base + i] = c.score;
		}
	}
}

struct M4Less   // CmpM4RecordByQidAndOvlpSize, pw_impl.cpp:539-548
{
	bool operator()(const mecat_m4& a, const mecat_m4& b) const
	{
		if (a.qid != b.qid) return a.qid < b.qid;
		const int64_t oa = std::min(a.qend - a.qoff, a.send - a.soff), ob = std::min(b.qend - b.qoff, b.send - b.soff);
		return oa > ob;
	}
};

int check(mecat_b200_ctx* c) { return c ? 0 : 1; }

// registry of the pinned result blocks of all contexts: block -> owning context (nullptr: the context is gone, the block
// is still in the caller's hands and is unpinned when it comes back)
std::mutex g_out_mu;
std::unordered_map<void*, mb::Ctx*> g_out;

void host_out_destroy(mb::Ctx* c)
{
	std::lock_guard<std::mutex> g(g_out_mu);
	for (auto& h : c->host_out) {
		if (h.used) g_out[h.p] = nullptr;
		else { g_out.erase(h.p); cudaFreeHost(h.p); }
	}
	c->host_out.clear();
}

}  // namespace

namespace mb {

void* host_out_alloc(Ctx* c, size_t bytes)
{
	if (bytes == 0) bytes = 64;
	std::lock_guard<std::mutex> g(g_out_mu);
	int best = -1;
	for (size_t i = 0; i < c->host_out.size(); ++i)
		if (!c->host_out[i].used && c->host_out[i].bytes >= bytes && (best < 0 || c->host_out[i].bytes < c->host_out[(size_t)best].bytes)) best = (int)i;
	if (best >= 0) { c->host_out[(size_t)best].used = true; return c->host_out[(size_t)best].p; }
	for (size_t i = 0; i < c->host_out.size();)          // free blocks too small to serve this size again: let them go
		if (!c->host_out[i].used) { g_out.erase(c->host_out[i].p); cudaFreeHost(c->host_out[i].p); c->host_out[i] = c->host_out.back(); c->host_out.pop_back(); }
		else ++i;
	void* p = nullptr;
	const size_t want = bytes + bytes / 8 + 4096;
	if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
	c->host_out.push_back({p, want, true});
	g_out[p] = c;
	return p;
}

bool host_out_release(void* p)
{
	std::lock_guard<std::mutex> g(g_out_mu);
	auto it = g_out.find(p);
	if (it == g_out.end()) return false;
	if (!it->second) { cudaFreeHost(p); g_out.erase(it); return true; }
	for (auto& h : it->second->host_out) if (h.p == p) h.used = false;
	return true;
}

}  // namespace mb

extern "C" {

int mecat_b200_abi_version(void) { return MECAT_B200_ABI_VERSION; }

int mecat_b200_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

int mecat_b200_init(mecat_b200_ctx** out, int device, void* /*nccl_comm_or_null*/)
{
	if (!out) return 1;
	*out = nullptr;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return 2;   // no CPU fallback
	if (cudaSetDevice(device) != cudaSuccess) return 3;
	mecat_b200_ctx* c = new mecat_b200_ctx;
	c->device = device;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
	memset(&c->stats, 0, sizeof c->stats);
	{
		// column arenas of the string-producing extension: 1/16 of the device memory each (11 GB on a 180 GB B200), so a
		// batch holds a few hundred thousand extensions and the consensus stages behind it see enough units per launch
		size_t free_b = 0, total_b = 0;
		if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b) {
			c->align_arena = std::max<size_t>(64ull << 20, std::min<size_t>(total_b / 16, 12ull << 30));
			if (const char* e = getenv("MECAT_B200_ALIGN_ARENA_MB")) c->align_arena = std::max<size_t>(1, (size_t)atoll(e)) << 20;   // test hook: force several batches
		}
	}
	if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return 4; }
	if (cudaMalloc(&c->d_counters, 16 * sizeof(unsigned long long)) != cudaSuccess) { delete c; return 5; }
	cudaMemset(c->d_counters, 0, 16 * sizeof(unsigned long long));
	*out = c;
	return 0;
}

void mecat_b200_destroy(mecat_b200_ctx* c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	cudaStreamSynchronize(c->stream);
	c->resolve_timers();
	for (auto& b : c->blocks) b.used = false;
	c->trim();
	for (auto& e : c->pool) cudaEventDestroy(e);
	for (auto& h : c->hstage) if (h.p) cudaFreeHost(h.p);
	host_out_destroy(c);
	cudaFree(c->d_counters);
	cudaStreamDestroy(c->stream);
	delete c;
}

const char* mecat_b200_last_error(mecat_b200_ctx* c) { return c ? c->err.c_str() : "null context"; }
void mecat_b200_free(mecat_b200_ctx*, void* p)
{
	if (!p) return;
	if (host_out_release(p)) return;       // a pinned result block: back to its context's pool
	free(p);
}

int mecat_b200_get_stats(mecat_b200_ctx* c, mecat_b200_stats* out)
{
	if (check(c) || !out) return 1;
	*out = c->stats;
	return 0;
}

int mecat_b200_reset_stats(mecat_b200_ctx* c)
{
	if (check(c)) return 1;
	memset(&c->stats, 0, sizeof c->stats);
	return 0;
}

int mecat_b200_volume_upload(mecat_b200_ctx* c, const mecat_volume* v, void** dvol)
{
	if (check(c) || !dvol) return 1;
	cudaSetDevice(c->device);
	DVolume* d = nullptr;
	WallTimer t;
	int rc = volume_upload(c, v, &d);
	c->stats.h2d_ms += t.stop();
	if (rc) return rc;
	*dvol = d;
	return 0;
}

int mecat_b200_volume_release(mecat_b200_ctx* c, void* dvol)
{
	if (check(c)) return 1;
	cudaSetDevice(c->device);
	volume_release(c, (DVolume*)dvol);
	return 0;
}

int mecat_b200_index_build(mecat_b200_ctx* c, void* dvol_ref, void** index)
{
	if (check(c) || !dvol_ref || !index) return 1;
	cudaSetDevice(c->device);
	DIndex* idx = nullptr;
	WallTimer t;
	int rc = index_build(c, (DVolume*)dvol_ref, &idx);
	c->stats.wall_index_ms += t.stop();
	if (rc) return rc;
	*index = idx;
	return 0;
}

int mecat_b200_index_count_part(mecat_b200_ctx* c, void* dvol_ref, uint32_t code_lo, uint32_t code_hi, void** index)
{
	if (check(c) || !dvol_ref || !index) return 1;
	cudaSetDevice(c->device);
	DIndex* idx = nullptr;
	int rc = index_count_part(c, (DVolume*)dvol_ref, code_lo, code_hi, &idx);
	if (rc) return rc;
	*index = idx;
	return 0;
}

int mecat_b200_index_finish_part(mecat_b200_ctx* c, void* dvol_ref, void* index, uint32_t code_lo, uint32_t code_hi)
{
	if (check(c) || !dvol_ref || !index) return 1;
	cudaSetDevice(c->device);
	return index_finish_part(c, (DVolume*)dvol_ref, (DIndex*)index, code_lo, code_hi);
}

int mecat_b200_index_device_arrays(mecat_b200_ctx* c, void* index, void** d_counts, void** d_begin, void** d_positions,
                                   int64_t* num_kmers)
{
	if (check(c) || !index) return 1;
	DIndex* I = (DIndex*)index;
	if (d_counts) *d_counts = I->counts;
	if (d_begin) *d_begin = I->begin;
	if (d_positions) *d_positions = I->pos;
	if (num_kmers) *num_kmers = I->num_kmers;
	return 0;
}

int mecat_b200_index_release(mecat_b200_ctx* c, void* index)
{
	if (check(c)) return 1;
	cudaSetDevice(c->device);
	index_release(c, (DIndex*)index);
	return 0;
}

int mecat_b200_index_export(mecat_b200_ctx* c, void* index, int64_t* num_kmers, uint32_t* begin, int32_t* positions)
{
	if (check(c) || !index) return 1;
	cudaSetDevice(c->device);
	DIndex* I = (DIndex*)index;
	if (num_kmers) *num_kmers = I->num_kmers;
	if (begin) MB_CUDA(c, cudaMemcpy(begin, I->begin, sizeof(uint32_t) * ((size_t)NCODES + 1), cudaMemcpyDeviceToHost));
	if (positions && I->num_kmers) MB_CUDA(c, cudaMemcpy(positions, I->pos, sizeof(int32_t) * (size_t)I->num_kmers, cudaMemcpyDeviceToHost));
	return 0;
}

int mecat_b200_extend_batch(mecat_b200_ctx* c, int policy, void* dq, void* ds, const mecat_extend_task* tasks,
                            size_t ntasks, int min_align_size, mecat_extend_result** results)
{
	if (check(c) || !dq || !ds || !results) return 1;
	if (policy != 0 && policy != 2) MB_FAIL(c, "extend_batch: policy %d not available (0 = pw/ref flavour, 2 = nanopore pw/ref flavour)", policy);
	cudaSetDevice(c->device);
	*results = nullptr;
	if (!ntasks) return 0;
	const DVolume* Q = (const DVolume*)dq;
	const DVolume* S = (const DVolume*)ds;
	for (size_t i = 0; i < ntasks; ++i) {
		const mecat_extend_task& t = tasks[i];
		if (t.qread < 0 || t.qread >= Q->num_reads || t.sread < 0 || t.sread >= S->num_reads)
			MB_FAIL(c, "extend_batch: task %zu names a read outside its volume", i);
		if (t.qstart < 0 || t.qstart > Q->h_offsz[2 * t.qread + 1] || t.sstart < 0 || t.sstart > S->h_offsz[2 * t.sread + 1])
			MB_FAIL(c, "extend_batch: task %zu start point outside its read", i);
	}
	static_assert(sizeof(ExtendTask) == sizeof(mecat_extend_task), "task layout");
	ExtendTask* d_tasks = nullptr;
	ExtendHalf* d_halves = nullptr;
	mecat_extend_result* d_res = nullptr;
	int rc = 0;
	auto body = [&]() -> int {
		MB_CUDA(c, c->alloc(&d_tasks, (size_t)(ntasks)));
		MB_CUDA(c, c->alloc(&d_halves, (size_t)(2 * ntasks)));
		MB_CUDA(c, c->alloc(&d_res, (size_t)(ntasks)));
		MB_CUDA(c, cudaMemcpyAsync(d_tasks, tasks, sizeof(ExtendTask) * ntasks, cudaMemcpyHostToDevice, c->stream));
		if (policy == 2) {
			if (xdrop_extend(c, Q, S, d_tasks, ntasks, min_align_size, d_res)) return 1;
		} else {
			if (extend_launch(c, Q, S, d_tasks, ntasks, d_halves)) return 1;
			KScope ks(c, MECAT_K_FINAL);
			k_extend_finalize<<<(unsigned)((ntasks + 255) / 256), 256, 0, c->stream>>>(d_tasks, d_halves, ntasks, min_align_size, d_res);
		}
		MB_CUDA(c, cudaGetLastError());
		mecat_extend_result* h = (mecat_extend_result*)malloc(sizeof(mecat_extend_result) * ntasks);
		if (!h) MB_FAIL(c, "extend_batch: out of host memory");
		cudaError_t e = cudaMemcpyAsync(h, d_res, sizeof(mecat_extend_result) * ntasks, cudaMemcpyDeviceToHost, c->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
		if (e != cudaSuccess) { free(h); MB_FAIL(c, "extend_batch: D2H: %s", cudaGetErrorString(e)); }
		c->resolve_timers();
		c->stats.h2d_bytes += (int64_t)(sizeof(ExtendTask) * ntasks);
		c->stats.d2h_bytes += (int64_t)(sizeof(mecat_extend_result) * ntasks);
		*results = h;
		return 0;
	};
	rc = body();
	c->dfree(d_tasks); c->dfree(d_halves); c->dfree(d_res);
	return rc;
}

int mecat_b200_align_batch(mecat_b200_ctx* c, int policy, double err, void* dq, void* ds, const mecat_align_task* tasks,
                           size_t ntasks, int min_align_size, mecat_align_result** results, char** qstrings,
                           char** sstrings, size_t* string_bytes)
{
	if (check(c) || !dq || !ds || !results || !qstrings || !sstrings || !string_bytes) return 1;
	if (policy < 0 || policy > 2) MB_FAIL(c, "align_batch: policy must be 0 (pw/ref), 1 (cns) or 2 (nanopore pw/ref), not %d", policy);
	cudaSetDevice(c->device);
	*results = nullptr; *qstrings = nullptr; *sstrings = nullptr; *string_bytes = 0;
	if (!ntasks) return 0;
	const DVolume* Q = (const DVolume*)dq;
	const DVolume* S = (const DVolume*)ds;
	static_assert(sizeof(AlignTask) == sizeof(mecat_align_task), "task layout");
	for (size_t i = 0; i < ntasks; ++i) {
		const mecat_align_task& t = tasks[i];
		if (t.qread < 0 || t.qread >= Q->num_reads || t.sread < 0 || t.sread >= S->num_reads)
			MB_FAIL(c, "align_batch: task %zu names a read outside its volume", i);
		const int ql = Q->h_offsz[2 * t.qread + 1], sl_full = S->h_offsz[2 * t.sread + 1];
		if (t.swin_len < 0 || t.swin_off < 0 || (t.swin_len > 0 && (int64_t)t.swin_off + t.swin_len > sl_full))
			MB_FAIL(c, "align_batch: task %zu subject window outside its read", i);
		const int sl = t.swin_len > 0 ? t.swin_len : sl_full;
		if (t.qstart < 0 || t.qstart > ql || t.sstart < 0 || t.sstart > sl)
			MB_FAIL(c, "align_batch: task %zu start point outside its sequence", i);
	}
	mecat_align_result* res = (mecat_align_result*)malloc(sizeof(mecat_align_result) * ntasks);
	if (!res) MB_FAIL(c, "align_batch: out of host memory");
	std::vector<char> qs, ss;
	if (align_batch(c, policy, err, Q, S, (const AlignTask*)tasks, ntasks, min_align_size, res, qs, ss)) { free(res); return 1; }
	char* a = (char*)malloc(qs.size() + 1);
	char* b = (char*)malloc(ss.size() + 1);
	if (!a || !b) { free(res); free(a); free(b); MB_FAIL(c, "align_batch: out of host memory"); }
	if (!qs.empty()) { memcpy(a, qs.data(), qs.size()); memcpy(b, ss.data(), ss.size()); }
	a[qs.size()] = 0; b[ss.size()] = 0;
	*results = res; *qstrings = a; *sstrings = b; *string_bytes = qs.size();
	return 0;
}

// ------------------------------------------------------------------------------------------
// mecat2cns -i 0 for a set of reads: GPU extensions (policy 1, align.cu) feeding the GPU consensus stage (cns.cu).
// The host groups the candidates by read, orders them for the accept loop and cuts batches; alignments, votes,
// graphs and corrected bases stay in device memory until the finished pieces are copied back.
void mecat_b200_cns_sort_candidates(mecat_candidate* cnd, int n)   // CmpExtensionCandidateByScore, mecat_correction.cpp:362-370
{
	std::sort(cnd, cnd + n, [](const mecat_candidate& a, const mecat_candidate& b) {
		if (a.score != b.score) return a.score > b.score;
		if (a.qid != b.qid) return a.qid < b.qid;
		return a.qext < b.qext;
	});
}

void mecat_b200_host_free(void* p) { mecat_b200_free(nullptr, p); }

// ------------------------------------------------------------------------------------------
// mecat2asmpw / mecat2trimpw (asmpw.cu)
int mecat_b200_asm_index_build(mecat_b200_ctx* c, const mecat_asm_reads* subject, void** asmidx)
{
	if (check(c) || !subject || !asmidx) return 1;
	cudaSetDevice(c->device);
	WallTimer wt;
	AsmIndexDev* I = nullptr;
	const int rc = asm_index_build(c, subject, &I);
	c->stats.wall_index_ms += wt.stop();
	if (rc) return rc;
	*asmidx = I;
	return 0;
}

int mecat_b200_asm_index_release(mecat_b200_ctx* c, void* asmidx)
{
	if (check(c)) return 1;
	cudaSetDevice(c->device);
	asm_index_release(c, (AsmIndexDev*)asmidx);
	return 0;
}

int mecat_b200_asm_index_export(mecat_b200_ctx* c, void* asmidx, int64_t* num_positions, uint32_t* begin, int32_t* positions)
{
	if (check(c) || !asmidx || !num_positions) return 1;
	cudaSetDevice(c->device);
	return asm_index_export(c, (const AsmIndexDev*)asmidx, num_positions, begin, positions);
}

int mecat_b200_asm_overlaps(mecat_b200_ctx* c, void* asmidx, const mecat_asm_reads* query, const mecat_asm_params* p,
                            mecat_asm_overlap** overlaps, size_t* n)
{
	if (check(c) || !asmidx || !query || !p || !overlaps || !n) return 1;
	cudaSetDevice(c->device);
	WallTimer wt;
	const int rc = asm_overlaps(c, (const AsmIndexDev*)asmidx, query, p, overlaps, n);
	c->stats.total_ms += wt.stop();
	return rc;
}

// ------------------------------------------------------------------------------------------
// mecat2ref: genome upload + k-mer index, then batches of reads through refmap.cu
int mecat_b200_ref_index_build(mecat_b200_ctx* c, const mecat_ref_genome* g, void** refidx)
{
	if (check(c) || !g || !refidx) return 1;
	cudaSetDevice(c->device);
	WallTimer wt;
	RefIndex* R = nullptr;
	const int rc = ref_index_build(c, g, &R);
	c->stats.wall_index_ms += wt.stop();
	if (rc) return rc;
	*refidx = R;
	return 0;
}

int mecat_b200_ref_index_release(mecat_b200_ctx* c, void* refidx)
{
	if (check(c)) return 1;
	cudaSetDevice(c->device);
	ref_index_release(c, (RefIndex*)refidx);
	return 0;
}

int mecat_b200_ref_index_export(mecat_b200_ctx* c, void* refidx, int64_t* num_kmers, uint32_t* begin, int32_t* positions)
{
	if (check(c) || !refidx) return 1;
	return mecat_b200_index_export(c, ((RefIndex*)refidx)->index, num_kmers, begin, positions);
}

int mecat_b200_ref_raw_candidates(mecat_b200_ctx* c, void* refidx, const mecat_ref_reads* reads, const mecat_ref_params* p,
                                  int32_t** rows, int32_t** counts, size_t* n)
{
	if (check(c) || !refidx || !reads || !p || !rows || !counts || !n) return 1;
	cudaSetDevice(c->device);
	mbref::Sink sink;
	std::vector<int32_t> cnt, row;
	if (ref_map(c, (const RefIndex*)refidx, reads, p, sink, &cnt, &row)) return 1;
	int32_t* r = (int32_t*)malloc(sizeof(int32_t) * (row.size() ? row.size() : 1));
	int32_t* k = (int32_t*)malloc(sizeof(int32_t) * (cnt.size() ? cnt.size() : 1));
	if (!r || !k) { free(r); free(k); MB_FAIL(c, "ref_raw_candidates: out of host memory"); }
	if (!row.empty()) memcpy(r, row.data(), sizeof(int32_t) * row.size());
	if (!cnt.empty()) memcpy(k, cnt.data(), sizeof(int32_t) * cnt.size());
	*rows = r; *counts = k; *n = row.size() / 4;
	return 0;
}

int mecat_b200_ref_map(mecat_b200_ctx* c, void* refidx, const mecat_ref_reads* reads, const mecat_ref_params* p,
                       mecat_ref_result** results, size_t* n, char** qstrings, char** sstrings, size_t* string_bytes)
{
	if (check(c) || !refidx || !reads || !p || !results || !n || !qstrings || !sstrings || !string_bytes) return 1;
	cudaSetDevice(c->device);
	*results = nullptr; *n = 0; *qstrings = nullptr; *sstrings = nullptr; *string_bytes = 0;
	mbref::Sink sink;
	if (ref_map(c, (const RefIndex*)refidx, reads, p, sink)) return 1;
	mecat_ref_result* res = (mecat_ref_result*)malloc(sizeof(mecat_ref_result) * (sink.recs.size() ? sink.recs.size() : 1));
	if (!res) MB_FAIL(c, "ref_map: out of host memory");
	if (!sink.recs.empty()) memcpy(res, sink.recs.data(), sizeof(mecat_ref_result) * sink.recs.size());
	*results = res; *n = sink.recs.size(); *string_bytes = sink.q.size();
	*qstrings = sink.q.release(); *sstrings = sink.s.release();      // the blobs as they were filled: no copy
	return 0;
}

// The reads to correct of one call: candidates sorted by template read (sid), one group per template that has enough of
// them (reads_correction_func_can, reads_correction_can.cpp:27-33), candidates in trial order, at most MAX_TRIED.
struct CnsGroup { size_t b, e; };

static void cns_groups(std::vector<mecat_candidate>& ec, const mecat_cns_params* p, std::vector<CnsGroup>& groups)
{
	auto by_sid = [](const mecat_candidate& a, const mecat_candidate& b) { return a.sid < b.sid; };
	if (p->input_type == 1) {
		// M4 input: the order IS the reference's std::sort -- of the partition by sid (build_cns_thrd_data_can,
		// reads_correction_aux.cpp:102; equal keys stay where the algorithm leaves them), then of a read with more overlaps
		// than fit by overlap size (CompareOverlapByOverlapSize, mecat_correction.cpp:26-34,261-272).  This library is built
		// without the parallel mode, so std::sort here is the sequential introsort the reference runs with one OpenMP thread.
		std::sort(ec.begin(), ec.end(), by_sid);
		const size_t cap = p->tech == 1 ? (size_t)mbcns::MAX_ACCEPT : (size_t)mbcns::MAX_ACCEPT_PACBIO;
		const size_t nec = ec.size();
		for (size_t i = 0; i < nec;) {
			size_t j = i + 1;
			while (j < nec && ec[j].sid == ec[i].sid) ++j;
			if ((int64_t)(j - i) >= p->min_cov && !(ec[i].ssize < p->min_size * 0.95)) {
				if (j - i > cap)
					std::sort(ec.begin() + (ptrdiff_t)i, ec.begin() + (ptrdiff_t)j, [](const mecat_candidate& a, const mecat_candidate& b) {
						return std::max(a.qend - a.qoff, a.send - a.soff) > std::max(b.qend - b.qoff, b.send - b.soff);
					});
				groups.push_back(CnsGroup{i, std::min(j, i + cap)});
			}
			i = j;
		}
		return;
	}
	if (!std::is_sorted(ec.begin(), ec.end(), by_sid)) std::stable_sort(ec.begin(), ec.end(), by_sid);
	const size_t nec = ec.size();
	for (size_t i = 0; i < nec;) {
		size_t j = i + 1;
		while (j < nec && ec[j].sid == ec[i].sid) ++j;
		if ((int64_t)(j - i) >= p->min_cov && !(ec[i].ssize < p->min_size * 0.95)) {
			mecat_b200_cns_sort_candidates(ec.data() + i, (int)(j - i));
			groups.push_back(CnsGroup{i, std::min(j, i + (size_t)mbcns::MAX_TRIED)});
		}
		i = j;
	}
}

// Extensions + consensus of groups [g0, gend) whose reads all live in V (candidate ids are V's).  Pieces are appended to
// `all` under the candidates' template ids.
static int cns_core(mecat_b200_ctx* c, const DVolume* V, const std::vector<mecat_candidate>& ec, const std::vector<CnsGroup>& groups,
                    size_t g0, size_t gend, const mecat_cns_params* p, CnsBlob& all, bool debug)
{
	const int id0 = V->start_read_id;
	mbcns::Params P;
	P.min_mapping_ratio = p->min_mapping_ratio; P.min_align_size = p->min_align_size; P.min_cov = p->min_cov; P.min_size = p->min_size;
	P.tech = p->tech; P.input_type = p->input_type;
	const double err = p->tech == 1 ? 0.20 : 0.15;     // GetAlignment's error rate, mecat_correction.cpp:424,487
	{
		// the corrected reads are about as long as their templates
		size_t bases = 0;
		for (size_t g = g0; g < gend; ++g) bases += (size_t)ec[groups[g].b].ssize;
		all.reserve(bases + bases / 16);
	}
	const size_t TASKS_PER_BATCH = 400000;       // with the column arena (12 GB at ~30 kB per task) this bounds a batch; more units per launch suit the latency-bound stages
	std::vector<AlignTask> tasks;
	std::vector<int32_t> info, first, rsize, tqid, tqsize;
	std::vector<int64_t> rid;
	while (g0 < gend) {
		WallTimer t_batch;
		// a batch = whole reads, as many as fit the column arena of the extension kernels
		tasks.clear(); first.assign(1, 0); rsize.clear(); rid.clear(); tqid.clear(); tqsize.clear();
		size_t g1 = g0, cols = 0;
		while (g1 < gend) {
			size_t need = 0;
			const size_t mark = tasks.size();
			for (size_t k = groups[g1].b; k < groups[g1].e; ++k) {
				const mecat_candidate& e = ec[k];
				AlignTask t;
				t.qread = e.qid - id0; t.qstrand = e.qdir; t.qstart = e.qdir ? e.qsize - 1 - e.qext : e.qext;   // mecat_correction.cpp:421-423
				t.sread = e.sid - id0; t.sstart = e.sext; t.swin_off = 0; t.swin_len = 0;
				need += align_task_columns(V, V, t);
				tasks.push_back(t);
			}
			if (g1 > g0 && (cols + need > c->align_arena || tasks.size() > TASKS_PER_BATCH)) { tasks.resize(mark); break; }
			if (cols + need > c->align_arena) MB_FAIL(c, "cns_reads: the candidates of read %d alone exceed the column arena", ec[groups[g1].b].sid);
			cols += need;
			for (size_t k = groups[g1].b; k < groups[g1].e; ++k) { tqid.push_back(ec[k].qid); tqsize.push_back(ec[k].qsize); }
			first.push_back((int32_t)tasks.size());
			rsize.push_back(ec[groups[g1].b].ssize);
			rid.push_back(ec[groups[g1].b].sid);
			++g1;
		}
		const float ms_tasks = t_batch.stop();
		AlignDev dev;
		if (align_batch_device(c, 1, err, V, V, tasks.data(), tasks.size(), p->min_align_size, &dev, info)) return 1;
		const float ms_align = t_batch.stop();
		mbcns::BatchIn in;
		in.R = (int)(g1 - g0); in.T = (int64_t)tasks.size();
		in.h_first = first.data(); in.h_read_size = rsize.data(); in.h_read_id = rid.data(); in.h_tqid = tqid.data(); in.h_tqsize = tqsize.data();
		in.d_info = dev.d_info; in.d_q = dev.d_packq; in.d_s = dev.d_packt; in.d_outoff = dev.d_outoff;
		const int rc = cns_consensus_device(c, in, P, all);
		align_dev_release(c, &dev);
		if (rc) return 1;
		if (debug) fprintf(stderr, "[cns_reads] batch of %zu reads / %zu tasks: task list %.1f ms, extensions %.1f ms, consensus %.1f ms\n", g1 - g0,
		                   tasks.size(), ms_tasks, ms_align - ms_tasks, t_batch.stop() - ms_align);
		g0 = g1;
	}
	return 0;
}

static int cns_hand_out(mecat_b200_ctx* c, CnsBlob& all, mecat_cns_piece** pieces, size_t* npieces, char** seqs, size_t* seq_bytes)
{
	if (all.oom) MB_FAIL(c, "cns_reads: out of host memory");
	const size_t bytes = all.len, np = all.recs.size();
	mecat_cns_piece* out = (mecat_cns_piece*)malloc(sizeof(mecat_cns_piece) * (np ? np : 1));
	char* sq = all.buf ? all.buf : (char*)malloc(1);
	if (!out || !sq) { free(out); MB_FAIL(c, "cns_reads: out of host memory"); }
	all.buf = nullptr;                       // the blob now belongs to the caller (mecat_b200_free)
	if (np) memcpy(out, all.recs.data(), sizeof(mecat_cns_piece) * np);
	sq[bytes] = 0;
	c->stats.num_records += (int64_t)np;
	*pieces = out; *npieces = np; *seqs = sq; *seq_bytes = bytes;
	return 0;
}

int mecat_b200_cns_reads(mecat_b200_ctx* c, void* dvol_reads, const mecat_candidate* ec_in, size_t nec, const mecat_cns_params* p,
                         mecat_cns_piece** pieces, size_t* npieces, char** seqs, size_t* seq_bytes)
{
	if (check(c) || !dvol_reads || (!ec_in && nec) || !p || !pieces || !npieces || !seqs || !seq_bytes) return 1;
	cudaSetDevice(c->device);
	*pieces = nullptr; *npieces = 0; *seqs = nullptr; *seq_bytes = 0;
	const DVolume* V = (const DVolume*)dvol_reads;
	const int id0 = V->start_read_id;
	for (size_t i = 0; i < nec; ++i) {
		const mecat_candidate& e = ec_in[i];
		if (e.sid < id0 || e.sid >= id0 + V->num_reads || e.qid < id0 || e.qid >= id0 + V->num_reads)
			MB_FAIL(c, "cns_reads: candidate %zu names a read outside the volume", i);
		if (e.sdir != 0) MB_FAIL(c, "cns_reads: candidate %zu is not normalised (sdir must be 0)", i);
		if (e.qsize != V->h_offsz[2 * (e.qid - id0) + 1] || e.ssize != V->h_offsz[2 * (e.sid - id0) + 1])
			MB_FAIL(c, "cns_reads: candidate %zu disagrees with the volume about read sizes", i);
		if (e.qext < 0 || e.qext >= e.qsize || e.sext < 0 || e.sext >= e.ssize)
			MB_FAIL(c, "cns_reads: candidate %zu extension point outside its read", i);
	}
	const bool debug = getenv("MECAT_CNS_DEBUG") != nullptr;
	WallTimer t_prep;
	std::vector<mecat_candidate> ec(ec_in, ec_in + nec);
	std::vector<CnsGroup> groups;
	cns_groups(ec, p, groups);
	CnsBlob all;
	if (debug) fprintf(stderr, "[cns_reads] validate + sort + group: %.1f ms\n", t_prep.stop());
	if (cns_core(c, V, ec, groups, 0, groups.size(), p, all, debug)) return 1;
	return cns_hand_out(c, all, pieces, npieces, seqs, seq_bytes);
}

// The same for a read set that spans several resident volumes (PackedDB::load_fasta_db keeps the whole data set in one
// store with 64-bit offsets, src/common/packed_db.cpp:194; a device volume addresses 2^31 bases).  The templates are
// taken in runs whose reads -- templates plus the reads of their (at most MAX_TRIED) candidates -- fit one working
// volume; that volume is gathered ON THE DEVICE from the resident ones (volume_gather, no host copy of bases) and the
// run goes through the single-volume path with working-volume ids, which are mapped back on the pieces.
int mecat_b200_cns_reads_multi(mecat_b200_ctx* c, void* const* dvols, int nvols, const mecat_candidate* ec_in, size_t nec,
                               const mecat_cns_params* p, mecat_cns_piece** pieces, size_t* npieces, char** seqs, size_t* seq_bytes)
{
	if (check(c) || !dvols || nvols < 1 || (!ec_in && nec) || !p || !pieces || !npieces || !seqs || !seq_bytes) return 1;
	if (nvols == 1 && !getenv("MECAT_B200_CNS_WORK_BASES"))          // one volume is its own working volume
		return mecat_b200_cns_reads(c, dvols[0], ec_in, nec, p, pieces, npieces, seqs, seq_bytes);
	cudaSetDevice(c->device);
	*pieces = nullptr; *npieces = 0; *seqs = nullptr; *seq_bytes = 0;
	std::vector<const DVolume*> V((size_t)nvols);
	for (int v = 0; v < nvols; ++v) {
		V[(size_t)v] = (const DVolume*)dvols[v];
		if (!V[(size_t)v]) MB_FAIL(c, "cns_reads_multi: null volume %d", v);
		if (v > 0 && V[(size_t)v]->start_read_id != V[(size_t)v - 1]->start_read_id + V[(size_t)v - 1]->num_reads)
			MB_FAIL(c, "cns_reads_multi: volume %d does not continue the read ids of volume %d", v, v - 1);
	}
	const int64_t id0 = V[0]->start_read_id, idn = (int64_t)V.back()->start_read_id + V.back()->num_reads;
	auto locate = [&](int64_t id, int& vol, int& rd) {           // volumes are few: binary search over their first ids
		int lo = 0, hi = nvols - 1;
		while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (V[(size_t)mid]->start_read_id <= id) lo = mid; else hi = mid - 1; }
		vol = lo; rd = (int)(id - V[(size_t)lo]->start_read_id);
	};
	for (size_t i = 0; i < nec; ++i) {
		const mecat_candidate& e = ec_in[i];
		if (e.sid < id0 || e.sid >= idn || e.qid < id0 || e.qid >= idn) MB_FAIL(c, "cns_reads_multi: candidate %zu names a read outside the volumes", i);
		if (e.sdir != 0) MB_FAIL(c, "cns_reads_multi: candidate %zu is not normalised (sdir must be 0)", i);
		int qv, qr, sv, sr;
		locate(e.qid, qv, qr); locate(e.sid, sv, sr);
		if (e.qsize != V[(size_t)qv]->h_offsz[2 * (size_t)qr + 1] || e.ssize != V[(size_t)sv]->h_offsz[2 * (size_t)sr + 1])
			MB_FAIL(c, "cns_reads_multi: candidate %zu disagrees with the volumes about read sizes", i);
		if (e.qext < 0 || e.qext >= e.qsize || e.sext < 0 || e.sext >= e.ssize)
			MB_FAIL(c, "cns_reads_multi: candidate %zu extension point outside its read", i);
	}
	const bool debug = getenv("MECAT_CNS_DEBUG") != nullptr;
	std::vector<mecat_candidate> ec(ec_in, ec_in + nec);
	std::vector<CnsGroup> groups;
	cns_groups(ec, p, groups);
	// bases one working volume may hold (test hook: MECAT_B200_CNS_WORK_BASES forces many small working volumes)
	int64_t work_cap = 1500000000LL;
	if (const char* e = getenv("MECAT_B200_CNS_WORK_BASES")) work_cap = std::max<int64_t>(1, atoll(e));
	CnsBlob all;
	std::vector<int32_t> local((size_t)(idn - id0), -1);        // global read -> read of the working volume
	std::vector<int32_t> src_vol, src_read;
	std::vector<int64_t> global_of;
	std::vector<mecat_candidate> run;
	std::vector<CnsGroup> run_groups;
	for (size_t g0 = 0; g0 < groups.size();) {
		src_vol.clear(); src_read.clear(); global_of.clear(); run.clear(); run_groups.clear();
		int64_t bases = 0;
		size_t g1 = g0;
		auto want = [&](int64_t id, int64_t size, std::vector<int64_t>& fresh) {
			if (local[(size_t)(id - id0)] >= 0) return;
			local[(size_t)(id - id0)] = -2;                         // claimed by the group under inspection
			fresh.push_back(id);
			bases += size + 1;
		};
		while (g1 < groups.size()) {
			std::vector<int64_t> fresh;
			const int64_t before = bases;
			want(ec[groups[g1].b].sid, ec[groups[g1].b].ssize, fresh);
			for (size_t k = groups[g1].b; k < groups[g1].e; ++k) want(ec[k].qid, ec[k].qsize, fresh);
			if (g1 > g0 && bases > work_cap) {                      // does not fit any more: the run ends before this group
				for (int64_t id : fresh) local[(size_t)(id - id0)] = -1;
				bases = before;
				break;
			}
			if (bases > 0x7ff00000LL) MB_FAIL(c, "cns_reads_multi: the reads of template %d alone exceed one working volume", ec[groups[g1].b].sid);
			for (int64_t id : fresh) {
				int v, r;
				locate(id, v, r);
				local[(size_t)(id - id0)] = (int32_t)global_of.size();
				src_vol.push_back(v); src_read.push_back(r); global_of.push_back(id);
			}
			CnsGroup lg; lg.b = run.size();
			for (size_t k = groups[g1].b; k < groups[g1].e; ++k) {
				mecat_candidate e = ec[k];
				e.qid = local[(size_t)(e.qid - id0)]; e.sid = local[(size_t)(e.sid - id0)];
				run.push_back(e);
			}
			lg.e = run.size();
			run_groups.push_back(lg);
			++g1;
		}
		DVolume* W = nullptr;
		if (volume_gather(c, V.data(), src_vol.data(), src_read.data(), (int)global_of.size(), &W)) return 1;
		if (debug) fprintf(stderr, "[cns_reads_multi] templates %zu..%zu: working volume of %zu reads, %lld bases\n", g0, g1, global_of.size(), (long long)bases);
		const size_t first_rec = all.recs.size();
		const int rc = cns_core(c, W, run, run_groups, 0, run_groups.size(), p, all, debug);
		volume_release(c, W);
		if (rc) return 1;
		for (size_t k = first_rec; k < all.recs.size(); ++k) all.recs[k].id = global_of[(size_t)all.recs[k].id];
		for (int64_t id : global_of) local[(size_t)(id - id0)] = -1;
		g0 = g1;
	}
	return cns_hand_out(c, all, pieces, npieces, seqs, seq_bytes);
}

// ------------------------------------------------------------------------------------------
// One (index volume, query volume) tile.
// text != NULL: the tile's result as the lines of the reference's output file instead of records (gapped: `-g 1`).
static int pw_tile_impl(mecat_b200_ctx* c, DIndex* idx, DVolume* ref, DVolume* reads, const mecat_pw_params* p,
                        int read_begin, int read_end, void** records, size_t* n, int32_t** raw_rows, int32_t** raw_counts,
                        char** text = nullptr, size_t* text_bytes = nullptr, int gapped = 0)
{
	if (text) { *text = nullptr; *text_bytes = 0; }
	if (read_begin < 0) read_begin = 0;
	if (read_end < 0 || read_end > reads->num_reads) read_end = reads->num_reads;
	if (p->num_candidates < 1) MB_FAIL(c, "pw_tile: number of candidates must be > 0");
	if (p->tech != 0 && p->tech != 1) MB_FAIL(c, "pw_tile: technology (-x) must be 0 (pacbio) or 1 (nanopore), not %d", p->tech);
	if (p->task != 0 && p->task != 1) MB_FAIL(c, "pw_tile: task (-j) must be 0 or 1, not %d", p->task);
	const int N = reads->num_reads, maxc = p->num_candidates;
	*n = 0;
	if (records) *records = nullptr;
	if (N == 0) return 0;
	RawCand* d_cands = nullptr;
	int32_t* d_counts = nullptr;
	int64_t* d_outpos = nullptr;
	mecat_candidate* d_ec = nullptr;
	ExtendTask* d_tasks = nullptr;
	ExtendHalf* d_halves = nullptr;
	mecat_extend_result* d_res = nullptr;
	int32_t* d_scores = nullptr;
	std::vector<int32_t> h_counts(N);
	std::vector<int64_t> h_outpos(N + 1);
	auto body = [&]() -> int {
		MB_CUDA(c, c->alloc(&d_cands, (size_t)((size_t)N * maxc)));
		MB_CUDA(c, c->alloc(&d_counts, (size_t)((size_t)N)));
		MB_CUDA(c, cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * (size_t)N, c->stream));
		{
			WallTimer ts;
			if (seed_candidates(c, idx, ref, reads, p, read_begin, read_end, d_cands, d_counts)) return 1;
			c->stats.wall_seed_ms += ts.stop();
		}
		MB_CUDA(c, cudaMemcpyAsync(h_counts.data(), d_counts, sizeof(int32_t) * (size_t)N, cudaMemcpyDeviceToHost, c->stream));
		MB_CUDA(c, cudaStreamSynchronize(c->stream));
		size_t total = 0;
		for (int i = 0; i < N; ++i) { h_outpos[i] = (int64_t)total; total += (size_t)h_counts[i]; }
		h_outpos[N] = (int64_t)total;
		c->stats.num_candidates += (int64_t)total;
		c->stats.d2h_bytes += (int64_t)sizeof(int32_t) * N;
		if (raw_rows) {
			// test hook: the raw candidate_save lists
			std::vector<RawCand> all((size_t)N * maxc);
			MB_CUDA(c, cudaMemcpy(all.data(), d_cands, sizeof(RawCand) * all.size(), cudaMemcpyDeviceToHost));
			int32_t* rows = (int32_t*)malloc(sizeof(RawCand) * (total ? total : 1));
			int32_t* cnts = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
			if (!rows || !cnts) { free(rows); free(cnts); MB_FAIL(c, "pw_tile: out of host memory"); }
			size_t k = 0;
			for (int r = 0; r < N; ++r) {
				memcpy(rows + 12 * k, all.data() + (size_t)r * maxc, sizeof(RawCand) * (size_t)h_counts[r]);
				k += (size_t)h_counts[r];
				cnts[r] = h_counts[r];
			}
			*raw_rows = rows; *raw_counts = cnts; *n = total;
			return 0;
		}
		if (total == 0) return 0;
		MB_CUDA(c, c->alloc(&d_outpos, (size_t)((size_t)(N + 1))));
		MB_CUDA(c, cudaMemcpyAsync(d_outpos, h_outpos.data(), sizeof(int64_t) * (size_t)(N + 1), cudaMemcpyHostToDevice, c->stream));
		// device text -> malloc'ed host text
		auto text_out = [&](int kind, const void* d_recs, size_t nrec) -> int {
			char* d_text = nullptr;
			size_t bytes = 0;
			if (records_text_device(c, kind, gapped, d_recs, nrec, &d_text, &bytes)) return 1;
			char* h = (char*)host_out_alloc(c, bytes + 1);
			if (!h) { c->dfree(d_text); MB_FAIL(c, "pw_tile: out of host memory"); }
			WallTimer t;
			cudaError_t e = bytes ? cudaMemcpyAsync(h, d_text, bytes, cudaMemcpyDeviceToHost, c->stream) : cudaSuccess;
			if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
			c->stats.d2h_ms += t.stop();
			c->dfree(d_text);
			if (e != cudaSuccess) { host_out_release(h); MB_FAIL(c, "pw_tile: D2H: %s", cudaGetErrorString(e)); }
			h[bytes] = 0;
			c->resolve_timers();
			c->stats.d2h_bytes += (int64_t)bytes;
			c->stats.num_records += (int64_t)nrec;
			*text = h; *text_bytes = bytes; *n = nrec;
			return 0;
		};
		if (p->task == 0) {
			MB_CUDA(c, c->alloc(&d_ec, (size_t)(total)));
			{
				KScope ks(c, MECAT_K_MERGE);
				k_make_ec<<<N, 32, 0, c->stream>>>(d_cands, d_counts, d_outpos, maxc, N, reads->offsz, reads->start_read_id,
				                                   ref->offsz, ref->start_read_id, d_ec, nullptr, nullptr);
			}
			MB_CUDA(c, cudaGetLastError());
			if (text) { c->stats.h2d_bytes += (int64_t)sizeof(int64_t) * (N + 1); return text_out(0, d_ec, total); }
			mecat_candidate* h = (mecat_candidate*)malloc(sizeof(mecat_candidate) * total);
			if (!h) MB_FAIL(c, "pw_tile: out of host memory");
			WallTimer t;
			cudaError_t e = cudaMemcpyAsync(h, d_ec, sizeof(mecat_candidate) * total, cudaMemcpyDeviceToHost, c->stream);
			if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
			c->stats.d2h_ms += t.stop();
			if (e != cudaSuccess) { free(h); MB_FAIL(c, "pw_tile: D2H: %s", cudaGetErrorString(e)); }
			c->resolve_timers();
			c->stats.d2h_bytes += (int64_t)(sizeof(mecat_candidate) * total);
			c->stats.h2d_bytes += (int64_t)sizeof(int64_t) * (N + 1);
			c->stats.num_records += (int64_t)total;
			*records = h; *n = total;
			return 0;
		}
		// task 1: extend every candidate, then assemble M4 records per read
		MB_CUDA(c, c->alloc(&d_tasks, (size_t)(total)));
		MB_CUDA(c, c->alloc(&d_halves, (size_t)(2 * total)));
		MB_CUDA(c, c->alloc(&d_res, (size_t)(total)));
		MB_CUDA(c, c->alloc(&d_scores, (size_t)(total)));
		{
			KScope ks(c, MECAT_K_MERGE);
			k_make_ec<<<N, 32, 0, c->stream>>>(d_cands, d_counts, d_outpos, maxc, N, reads->offsz, reads->start_read_id,
			                                   ref->offsz, ref->start_read_id, nullptr, d_tasks, d_scores);
		}
		MB_CUDA(c, cudaGetLastError());
		// A12 on the device (default): one extension launch over all candidates, then fill_m4record / std::sort /
		// containment filter per read in kernels (records.cu); only the kept records -- or their text -- go to the host.
		// MECAT_B200_M4=host selects the earlier form (host threads behind extension chunks).
		const char* m4_mode = getenv("MECAT_B200_M4");
		if (!(m4_mode && !strcmp(m4_mode, "host"))) {
			WallTimer te;
			if (p->tech == 1) {
				if (xdrop_extend(c, reads, ref, d_tasks, total, p->min_align_size, d_res)) return 1;
			} else {
				if (extend_launch(c, reads, ref, d_tasks, total, d_halves)) return 1;
				KScope ks(c, MECAT_K_FINAL);
				k_extend_finalize<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(d_tasks, d_halves, total, p->min_align_size, d_res);
			}
			MB_CUDA(c, cudaGetLastError());
			mecat_m4* d_m4 = nullptr;
			size_t nout = 0;
			if (m4_assemble(c, reads, ref, d_tasks, d_res, d_scores, d_outpos, N, total, &d_m4, &nout)) return 1;
			c->stats.wall_extend_ms += te.stop();
			c->stats.h2d_bytes += (int64_t)sizeof(int64_t) * (N + 1);
			int rc2 = 0;
			if (text) rc2 = nout ? text_out(1, d_m4, nout) : 0;
			else {
				mecat_m4* out = (mecat_m4*)host_out_alloc(c, sizeof(mecat_m4) * (nout ? nout : 1));
				if (!out) { c->dfree(d_m4); MB_FAIL(c, "pw_tile: out of host memory"); }
				WallTimer t;
				cudaError_t e = nout ? cudaMemcpyAsync(out, d_m4, sizeof(mecat_m4) * nout, cudaMemcpyDeviceToHost, c->stream) : cudaSuccess;
				if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
				c->stats.d2h_ms += t.stop();
				if (e != cudaSuccess) { host_out_release(out); c->dfree(d_m4); MB_FAIL(c, "pw_tile: D2H: %s", cudaGetErrorString(e)); }
				c->resolve_timers();
				c->stats.d2h_bytes += (int64_t)(sizeof(mecat_m4) * nout);
				c->stats.num_records += (int64_t)nout;
				*records = out; *n = nout;
			}
			c->dfree(d_m4);
			return rc2;
		}
		// The extension runs in a few chunks of reads; while the GPU extends chunk k+1 the host threads
		// assemble the M4 records of chunk k (fill_m4record + append_m4v: sort, containment filter) -- on
		// the host like the reference, same std::sort, same comparator, so ties fall the same way.
		ExtendTask* h_tasks = nullptr;
		mecat_extend_result* h_res = nullptr;
		int32_t* h_score = nullptr;
		MB_CUDA(c, c->host_stage(0, sizeof(ExtendTask) * total, (void**)&h_tasks));
		MB_CUDA(c, c->host_stage(1, sizeof(mecat_extend_result) * total, (void**)&h_res));
		MB_CUDA(c, c->host_stage(2, sizeof(int32_t) * total, (void**)&h_score));
		const int nthreads = std::max(1, std::min(32, (int)std::thread::hardware_concurrency()));
		// chunk ends as fractions of the candidates: the last chunks are small because the assembly of the
		// final one is the only host work the GPU cannot hide
		static const double cuts8[] = {0.16, 0.32, 0.48, 0.64, 0.78, 0.89, 0.96, 1.0};
		static const double cuts3[] = {0.45, 0.85, 1.0};
		static const double cuts1[] = {1.0};
		// three chunks (measured on configs[1]: 600 ms per step against 608 ms with eight -- every launch has a tail in which
		// the last chains run alone, and only the last chunk's assembly, now 15 % of the records, is not hidden)
		int npipe = total >= 200000 ? 3 : 1;
		if (const char* e = getenv("MECAT_B200_PIPE")) npipe = atoi(e) >= 8 ? 8 : atoi(e) >= 3 ? 3 : 1;      // tuning hook
		const double* cuts = npipe == 8 ? cuts8 : npipe == 3 ? cuts3 : cuts1;
		std::vector<int> rcut((size_t)npipe + 1, N);
		rcut[0] = 0;
		for (int k = 1, r = 0; k < npipe; ++k) {
			const int64_t want = (int64_t)((double)total * cuts[k - 1]);
			while (r < N && h_outpos[r] < want) ++r;
			rcut[k] = r;
		}
		const int sub = nthreads * 2;                                   // assembly pieces per pipeline chunk
		std::vector<std::vector<mecat_m4>> parts((size_t)npipe * sub);
		auto assemble = [&](int r_lo, int r_hi, std::vector<mecat_m4>& dst) {
			std::vector<mecat_m4> loc;
			std::vector<char> valid;
			for (int r = r_lo; r < r_hi; ++r) {
				loc.clear();
				const int64_t qsize = reads->h_offsz[2 * r + 1];
				const int64_t qid = r + reads->start_read_id;
				for (int64_t k = h_outpos[r]; k < h_outpos[r + 1]; ++k) {
					const mecat_extend_result& R = h_res[k];
					if (!R.ok) continue;
					const ExtendTask& t = h_tasks[k];
					mecat_m4 m;
					memset(&m, 0, sizeof m);
					m.qid = t.sread + ref->start_read_id; m.sid = qid; m.ident = R.ident; m.vscore = h_score[k]; m.qdir = 0;
					m.qoff = R.sstart; m.qend = R.send; m.qsize = ref->h_offsz[2 * t.sread + 1]; m.ssize = qsize; m.qext = t.sstart;
					if (!t.qstrand) { m.sdir = 0; m.soff = R.qstart; m.send = R.qend; m.sext = t.qstart; }
					else { m.sdir = 1; m.soff = qsize - R.qend; m.send = qsize - R.qstart; m.sext = qsize - 1 - t.qstart; }
					loc.push_back(m);
				}
				if (loc.empty()) continue;
				std::sort(loc.begin(), loc.end(), M4Less());
				valid.assign(loc.size(), 1);
				for (size_t i = 0; i < loc.size();) {
					size_t j = i + 1;
					while (j < loc.size() && loc[j].qid == loc[i].qid) ++j;
					for (size_t a = i; a < j; ++a) {
						if (!valid[a]) continue;
						for (size_t b = a + 1; b < j; ++b) {
							if (!valid[b] || loc[a].sdir != loc[b].sdir) continue;
							if (loc[b].qoff + 100 >= loc[a].qoff && loc[b].qend - 100 <= loc[a].qend &&
							    loc[b].soff + 100 >= loc[a].soff && loc[b].send - 100 <= loc[a].send) valid[b] = 0;
						}
					}
					i = j;
				}
				for (size_t i = 0; i < loc.size(); ++i) if (valid[i]) dst.push_back(loc[i]);
			}
		};
		auto assemble_chunk = [&](int k) {
			const int r0 = rcut[k], r1 = rcut[k + 1];
			std::atomic<int> next(0);
			auto runner = [&]() {
				for (int ch; (ch = next.fetch_add(1)) < sub;) {
					const int lo = r0 + (int)((int64_t)(r1 - r0) * ch / sub), hi = r0 + (int)((int64_t)(r1 - r0) * (ch + 1) / sub);
					assemble(lo, hi, parts[(size_t)k * sub + ch]);
				}
			};
			std::vector<std::thread> pool;
			for (int t = 1; t < nthreads; ++t) pool.emplace_back(runner);
			runner();
			for (auto& th : pool) th.join();
		};
		WallTimer host_timer;
		float host_hidden = 0;
		std::thread worker;
		int rc_pipe = 0;
		for (int k = 0; k < npipe && !rc_pipe; ++k) {
			const size_t t0 = (size_t)h_outpos[rcut[k]], t1 = (size_t)h_outpos[rcut[k + 1]];
			const size_t nt = t1 - t0;
			auto gpu_part = [&]() -> int {
				if (!nt) return 0;
				WallTimer te;
				if (p->tech == 1) {
					// nanopore: XdropAligner instead of DiffAligner (pw_impl.cpp:638-642)
					if (xdrop_extend(c, reads, ref, d_tasks + t0, nt, p->min_align_size, d_res + t0)) return 1;
					c->stats.wall_extend_ms += te.stop();
				} else {
					if (extend_launch(c, reads, ref, d_tasks + t0, nt, d_halves + 2 * t0)) return 1;
					c->stats.wall_extend_ms += te.stop();
					KScope ks(c, MECAT_K_FINAL);
					k_extend_finalize<<<(unsigned)((nt + 255) / 256), 256, 0, c->stream>>>(d_tasks + t0, d_halves + 2 * t0, nt, p->min_align_size, d_res + t0);
				}
				MB_CUDA(c, cudaGetLastError());
				WallTimer t;
				MB_CUDA(c, cudaMemcpyAsync(h_tasks + t0, d_tasks + t0, sizeof(ExtendTask) * nt, cudaMemcpyDeviceToHost, c->stream));
				MB_CUDA(c, cudaMemcpyAsync(h_res + t0, d_res + t0, sizeof(mecat_extend_result) * nt, cudaMemcpyDeviceToHost, c->stream));
				MB_CUDA(c, cudaMemcpyAsync(h_score + t0, d_scores + t0, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost, c->stream));
				MB_CUDA(c, cudaStreamSynchronize(c->stream));
				c->stats.d2h_ms += t.stop();
				c->resolve_timers();
				return 0;
			};
			rc_pipe = gpu_part();
			if (worker.joinable()) worker.join();
			if (!rc_pipe) worker = std::thread(assemble_chunk, k);
		}
		if (worker.joinable()) worker.join();
		if (rc_pipe) return 1;
		(void)host_hidden;
		c->stats.d2h_bytes += (int64_t)((sizeof(ExtendTask) + sizeof(mecat_extend_result) + 4) * total);
		c->stats.h2d_bytes += (int64_t)sizeof(int64_t) * (N + 1);
		size_t nout = 0;
		for (auto& v : parts) nout += v.size();
		mecat_m4* out = (mecat_m4*)malloc(sizeof(mecat_m4) * (nout ? nout : 1));
		if (!out) MB_FAIL(c, "pw_tile: out of host memory");
		{
			// concatenate in read order, in parallel (each piece knows its offset)
			std::vector<size_t> at(parts.size() + 1, 0);
			for (size_t i = 0; i < parts.size(); ++i) at[i + 1] = at[i] + parts[i].size();
			std::atomic<size_t> next(0);
			auto runner = [&]() { for (size_t i; (i = next.fetch_add(1)) < parts.size();) if (!parts[i].empty()) memcpy(out + at[i], parts[i].data(), sizeof(mecat_m4) * parts[i].size()); };
			std::vector<std::thread> pool;
			for (int t = 1; t < std::min(nthreads, 8); ++t) pool.emplace_back(runner);
			runner();
			for (auto& th : pool) th.join();
		}
		c->stats.host_ms += host_timer.stop();
		if (text) {      // host-assembled records, device text
			mecat_m4* d_m4 = nullptr;
			int rc2 = 0;
			if (nout) {
				MB_CUDA(c, c->alloc(&d_m4, nout));
				MB_CUDA(c, cudaMemcpyAsync(d_m4, out, sizeof(mecat_m4) * nout, cudaMemcpyHostToDevice, c->stream));
				c->stats.h2d_bytes += (int64_t)(sizeof(mecat_m4) * nout);
				rc2 = text_out(1, d_m4, nout);
				c->dfree(d_m4);
			}
			free(out);
			return rc2;
		}
		c->stats.num_records += (int64_t)nout;
		*records = out; *n = nout;
		return 0;
	};
	int rc = body();
	c->dfree(d_cands); c->dfree(d_counts); c->dfree(d_outpos); c->dfree(d_ec);
	c->dfree(d_tasks); c->dfree(d_halves); c->dfree(d_res); c->dfree(d_scores);
	return rc;
}

int mecat_b200_pw_tile(mecat_b200_ctx* c, void* index, void* dvol_ref, void* dvol_reads, const mecat_pw_params* p,
                       void** records, size_t* n)
{
	if (check(c) || !index || !dvol_ref || !dvol_reads || !p || !records || !n) return 1;
	cudaSetDevice(c->device);
	WallTimer t;
	int rc = pw_tile_impl(c, (DIndex*)index, (DVolume*)dvol_ref, (DVolume*)dvol_reads, p, 0, -1, records, n, nullptr, nullptr);
	c->stats.total_ms += t.stop();
	return rc;
}

int mecat_b200_pw_tile_range(mecat_b200_ctx* c, void* index, void* dvol_ref, void* dvol_reads, const mecat_pw_params* p,
                             int read_begin, int read_end, void** records, size_t* n)
{
	if (check(c) || !index || !dvol_ref || !dvol_reads || !p || !records || !n) return 1;
	cudaSetDevice(c->device);
	WallTimer t;
	int rc = pw_tile_impl(c, (DIndex*)index, (DVolume*)dvol_ref, (DVolume*)dvol_reads, p, read_begin, read_end, records, n, nullptr, nullptr);
	c->stats.total_ms += t.stop();
	return rc;
}

int mecat_b200_volume_from_text(mecat_b200_ctx* c, const char* text, size_t text_bytes, const int64_t* src_offset, const int32_t* offset_size,
                                int32_t num_reads, int32_t num_bases, int32_t start_read_id, uint8_t* pac_out, void** dvol)
{
	if (check(c) || !dvol) return 1;
	cudaSetDevice(c->device);
	DVolume* d = nullptr;
	const int rc = volume_from_text(c, text, text_bytes, src_offset, offset_size, num_reads, num_bases, start_read_id, pac_out, &d);
	if (rc) return rc;
	*dvol = d;
	return 0;
}

int mecat_b200_pw_tile_text(mecat_b200_ctx* c, void* index, void* dvol_ref, void* dvol_reads, const mecat_pw_params* p, int gapped,
                            char** text, size_t* bytes, size_t* num_records)
{
	if (check(c) || !index || !dvol_ref || !dvol_reads || !p || !text || !bytes || !num_records) return 1;
	cudaSetDevice(c->device);
	WallTimer t;
	void* unused = nullptr;
	int rc = pw_tile_impl(c, (DIndex*)index, (DVolume*)dvol_ref, (DVolume*)dvol_reads, p, 0, -1, &unused, num_records, nullptr, nullptr, text, bytes, gapped);
	c->stats.total_ms += t.stop();
	return rc;
}

int mecat_b200_records_text(mecat_b200_ctx* c, int kind, int gapped, const void* records, size_t n, char** text, size_t* bytes)
{
	if (check(c) || (n && !records) || !text || !bytes) return 1;
	if (kind != 0 && kind != 1) MB_FAIL(c, "records_text: kind must be 0 (candidates) or 1 (m4 records), not %d", kind);
	cudaSetDevice(c->device);
	*text = nullptr; *bytes = 0;
	const size_t rec = kind == 0 ? sizeof(mecat_candidate) : sizeof(mecat_m4);
	void* d_rec = nullptr;
	char* d_text = nullptr;
	char* h = nullptr;
	auto body = [&]() -> int {
		size_t nb = 0;
		if (n) {
			MB_CUDA(c, c->dmalloc(&d_rec, rec * n));
			MB_CUDA(c, cudaMemcpyAsync(d_rec, records, rec * n, cudaMemcpyHostToDevice, c->stream));
			if (records_text_device(c, kind, gapped, d_rec, n, &d_text, &nb)) return 1;
		}
		h = (char*)malloc(nb + 1);
		if (!h) MB_FAIL(c, "records_text: out of host memory");
		if (nb) MB_CUDA(c, cudaMemcpyAsync(h, d_text, nb, cudaMemcpyDeviceToHost, c->stream));
		MB_CUDA(c, cudaStreamSynchronize(c->stream));
		h[nb] = 0;
		c->resolve_timers();
		c->stats.h2d_bytes += (int64_t)(rec * n); c->stats.d2h_bytes += (int64_t)nb;
		*bytes = nb;
		return 0;
	};
	const int rc = body();
	c->dfree(d_rec); c->dfree(d_text);
	if (rc) { free(h); return rc; }
	*text = h;
	return 0;
}

int mecat_b200_volume_from_device(mecat_b200_ctx* c, int32_t num_reads, int32_t num_bases, int32_t start_read_id,
                                  const int32_t* host_offset_size, const void* device_pac, void** dvol)
{
	if (check(c) || !dvol || !host_offset_size || !device_pac) return 1;
	cudaSetDevice(c->device);
	DVolume* d = nullptr;
	int rc = volume_from_device(c, num_reads, num_bases, start_read_id, host_offset_size, (const uint8_t*)device_pac, &d);
	if (rc) return rc;
	*dvol = d;
	return 0;
}

int mecat_b200_pw_raw_candidates(mecat_b200_ctx* c, void* index, void* dvol_ref, void* dvol_reads,
                                 const mecat_pw_params* p, int32_t** rows, int32_t** counts, size_t* n)
{
	if (check(c) || !index || !dvol_ref || !dvol_reads || !p || !rows || !counts || !n) return 1;
	cudaSetDevice(c->device);
	return pw_tile_impl(c, (DIndex*)index, (DVolume*)dvol_ref, (DVolume*)dvol_reads, p, 0, -1, nullptr, n, rows, counts);
}

static int pw_host(mecat_b200_ctx* c, const mecat_volume* ref, const mecat_volume* reads, const mecat_pw_params* p, int task,
                   void** records, size_t* n)
{
	if (check(c) || !ref || !reads || !p || !records || !n) return 1;
	cudaSetDevice(c->device);
	mecat_pw_params q = *p;
	q.task = task;
	void *dref = nullptr, *dreads = nullptr, *idx = nullptr;
	const bool same = (ref == reads) || (ref->pac == reads->pac && ref->num_bases == reads->num_bases &&
	                                     ref->start_read_id == reads->start_read_id);
	int rc = mecat_b200_volume_upload(c, ref, &dref);
	if (!rc) { if (same) dreads = dref; else rc = mecat_b200_volume_upload(c, reads, &dreads); }
	if (!rc) rc = mecat_b200_index_build(c, dref, &idx);
	if (!rc) rc = mecat_b200_pw_tile(c, idx, dref, dreads, &q, records, n);
	if (idx) mecat_b200_index_release(c, idx);
	if (dreads && dreads != dref) mecat_b200_volume_release(c, dreads);
	if (dref) mecat_b200_volume_release(c, dref);
	return rc;
}

int mecat_b200_pw_candidates(mecat_b200_ctx* c, const mecat_volume* ref, const mecat_volume* reads,
                             const mecat_pw_params* p, mecat_candidate** ec, size_t* n)
{
	return pw_host(c, ref, reads, p, 0, (void**)ec, n);
}

int mecat_b200_pw_overlaps(mecat_b200_ctx* c, const mecat_volume* ref, const mecat_volume* reads,
                           const mecat_pw_params* p, mecat_m4** m4, size_t* n)
{
	return pw_host(c, ref, reads, p, 1, (void**)m4, n);
}

}  // extern "C"
