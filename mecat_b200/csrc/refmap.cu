// mecat_b200/csrc/refmap.cu -- mecat2ref on the GPU: CUDA backend of ref_pipeline.h.
//
// The stage sequence lives in ref_pipeline.h, the per-unit bodies in ref_core.cuh; here every stage functor F becomes a
// launch of k_ref<F> (one thread per strand / clipped end), memory comes from the context's pool, and the extension
// hook is the library's own gapped aligner with strings (align.cu, row R1) on windows of the genome.  The genome is an
// ordinary device volume with one "read"; its k-mer index is the index of index.cu built over a second offset table
// that cuts the genome's ACGT runs into chunks, so that a chunk is one CTA's work and no k-mer spans another letter.
#include "common.cuh"
#include "dev_backend.cuh"
#include "ref_pipeline.h"

#include <algorithm>

namespace mb {

namespace {

template <class F>
__global__ void __launch_bounds__(128) k_ref(const F f, const int64_t n)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) f(i);
}

struct RefWarpLanes         // the lanes interface of ref_core.cuh on a real warp
{
	static constexpr int count = 32;
	__device__ static int lane() { return (int)(threadIdx.x & 31u); }
	template <class F> __device__ void each(F&& f) const { f(lane()); }
	template <class F> __device__ int sum(F&& f) const { return __reduce_add_sync(0xffffffffu, f(lane())); }
	template <class F> __device__ uint32_t ballot(F&& f) const { return __ballot_sync(0xffffffffu, f(lane())); }
	template <class F> __device__ int min_val(F&& f) const { return __reduce_min_sync(0xffffffffu, f(lane())); }
	template <class F> __device__ int lead(F&& f) const
	{
		int v = 0;
		if (lane() == 0) v = f();
		__syncwarp();                               // the leader's stores are visible to the other lanes behind this
		return __shfl_sync(0xffffffffu, v, 0);
	}
	__device__ bool leader() const { return lane() == 0; }
	__device__ void sync() const { __syncwarp(); }
};

constexpr int SEED_WARPS = 4;

// 16 CTAs of 4 warps per SM (32 registers per thread): the kernel waits on memory (ncu: 12 long-scoreboard stall cycles per
// issue at 36 warps per SM), so resident warps count for more than registers
__global__ void __launch_bounds__(SEED_WARPS * 32, 16) k_ref_seed_warp(const mbref::SeedWarpFn f, const int64_t n)
{
	__shared__ mbref::WarpScratch scratch[SEED_WARPS];
	const int64_t u = (int64_t)blockIdx.x * SEED_WARPS + (threadIdx.x >> 5);
	if (u < n) f(u, RefWarpLanes(), scratch[threadIdx.x >> 5]);
}

struct RefBackend : PoolBackend
{
	const DVolume* reads;
	const DVolume* genome;
	int tech;                      // -x: 0 = DiffAligner, 1 = XdropAligner (mecat2ref_impl_large.cpp:329-332)
	RefBackend(Ctx* ctx, const DVolume* r, const DVolume* g, int tech_) : PoolBackend(ctx, "ref", true), reads(r), genome(g), tech(tech_) {}

	template <class F> bool launch(int64_t n, const F& f, int stage)
	{
		if (n <= 0) return true;
		KScope ks(c, MECAT_K_REF_COUNT + stage);
		k_ref<F><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(f, n);
		return check(cudaGetLastError(), "launch");
	}
	bool launch_seed_warp(int64_t n, const mbref::SeedWarpFn& f, int stage)
	{
		if (n <= 0) return true;
		KScope ks(c, MECAT_K_REF_COUNT + stage);
		k_ref_seed_warp<<<(unsigned)((n + SEED_WARPS - 1) / SEED_WARPS), SEED_WARPS * 32, 0, c->stream>>>(f, n);
		return check(cudaGetLastError(), "launch");
	}
	bool align(const mecat_align_task* tasks, size_t n, bool want_strings, mecat_align_result* res, std::vector<char>& qs, std::vector<char>& ss)
	{
		static_assert(sizeof(AlignTask) == sizeof(mecat_align_task), "AlignTask mirrors mecat_align_task");
		c->stats.num_candidates += (int64_t)n;
		const int min_aln = 1000;       // extend_candidate's min_aln, mecat2ref_aux.cpp:152
		// Without strings only coordinates, columns and matches are wanted: the forward pass of the extension gives them
		// (k_extend: no traceback, no column arenas).  Opt-in until it has run on hardware next to the default.
		if (tech == 1) return align_batch(c, 2, 0.0, reads, genome, (const AlignTask*)tasks, n, min_aln, res, qs, ss, want_strings) == 0;
		if (!want_strings && forward_only()) { qs.clear(); ss.clear(); return align_forward(tasks, n, min_aln, res); }
		return align_batch(c, 0, 0.0, reads, genome, (const AlignTask*)tasks, n, min_aln, res, qs, ss, want_strings) == 0;
	}
	static bool forward_only()
	{
		const char* mode = getenv("MECAT_B200_REF_EXTEND");
		return mode && !strcmp(mode, "forward");
	}
	// k_extend addresses its subject by read; the window of task t becomes "read" t of a second offset table over the
	// genome's bases (the same bases, no copy), so the kernel runs as it is.
	bool align_forward(const mecat_align_task* tasks, size_t n, int min_aln, mecat_align_result* res)
	{
		if (!n) return true;
		std::vector<int32_t> win(2 * n);
		std::vector<ExtendTask> et(n);
		for (size_t t = 0; t < n; ++t) {
			win[2 * t] = tasks[t].swin_off; win[2 * t + 1] = tasks[t].swin_len;
			ExtendTask e;
			e.qread = tasks[t].qread; e.qstrand = tasks[t].qstrand; e.qstart = tasks[t].qstart; e.sread = (int32_t)t; e.sstart = tasks[t].sstart;
			et[t] = e;
		}
		DVolume view;
		view.num_reads = (int32_t)n; view.num_bases = genome->num_bases; view.fwd = genome->fwd; view.rev = genome->rev; view.words = genome->words;
		int32_t* d_win = alloc<int32_t>(2 * n);
		ExtendTask* d_tasks = alloc<ExtendTask>(n);
		ExtendHalf* d_halves = alloc<ExtendHalf>(2 * n);
		if (!d_win || !d_tasks || !d_halves || !upload(d_win, win.data(), 2 * n) || !upload(d_tasks, et.data(), n)) return false;
		view.offsz = (int2*)d_win;
		if (extend_launch(c, reads, &view, d_tasks, n, d_halves)) return false;
		std::vector<ExtendHalf> halves(2 * n);
		if (!download(halves.data(), d_halves, 2 * n)) return false;
		for (size_t t = 0; t < n; ++t) {       // DiffAligner::go's accessors (diff_gapalign.cpp:295-349) from the two directions
			const ExtendHalf& L = halves[2 * t];
			const ExtendHalf& R = halves[2 * t + 1];
			mecat_align_result& r = res[t];
			memset(&r, 0, sizeof r);
			r.columns = L.cols + R.cols; r.matches = L.matches + R.matches;
			r.qstart = tasks[t].qstart - L.qadv; r.qend = tasks[t].qstart + R.qadv;
			r.sstart = tasks[t].sstart - L.tadv; r.send = tasks[t].sstart + R.tadv;
			r.ok = r.columns >= min_aln;
			r.ident = r.columns ? 100.0 * r.matches / r.columns : 0.0;
			r.str_offset = -1;
		}
		return release(d_win) && release(d_tasks) && release(d_halves);
	}
	void note_hits(int64_t n) { c->stats.num_hits += n; }
};

}  // namespace

int ref_index_build(Ctx* c, const mecat_ref_genome* g, RefIndex** out)
{
	if (!g || g->num_bases < 0 || g->num_runs < 0 || (g->num_bases && !g->pac)) MB_FAIL(c, "ref_index_build: bad genome");
	if (g->num_bases >= (1ll << 31) - (1ll << 20)) MB_FAIL(c, "ref_index_build: %lld bases; this path holds positions in 32 bits (< 2^31 - 2^20)", (long long)g->num_bases);
	RefIndex* R = new RefIndex;
	const int32_t one[2] = {0, (int32_t)g->num_bases};
	mecat_volume v;
	v.num_reads = 1; v.num_bases = (int32_t)g->num_bases; v.start_read_id = 0; v.offset_size = one; v.pac = g->pac;
	if (volume_upload(c, &v, &R->genome)) { delete R; return 1; }
	// the index's view of the same bases: chunks of the ACGT runs, overlapping by the 12 bases a k-mer needs beyond its start
	std::vector<int32_t> chunks;
	if (!mbref::index_chunks(g->run_start_len, g->num_runs, g->num_bases, chunks)) { ref_index_release(c, R); MB_FAIL(c, "ref_index_build: runs out of order or out of range"); }
	DVolume view;
	view.num_reads = (int32_t)(chunks.size() / 2); view.num_bases = R->genome->num_bases; view.fwd = R->genome->fwd; view.rev = R->genome->rev;
	view.words = R->genome->words;
	auto body = [&]() -> int {
		MB_CUDA(c, c->dmalloc((void**)&view.offsz, sizeof(int2) * (chunks.size() / 2 + 1)));
		if (!chunks.empty()) MB_CUDA(c, cudaMemcpyAsync(view.offsz, chunks.data(), sizeof(int32_t) * chunks.size(), cudaMemcpyHostToDevice, c->stream));
		MB_CUDA(c, cudaStreamSynchronize(c->stream));
		return index_build(c, &view, &R->index);
	};
	const int rc = body();
	c->dfree(view.offsz);
	if (rc) { ref_index_release(c, R); return rc; }
	*out = R;
	return 0;
}

void ref_index_release(Ctx* c, RefIndex* R)
{
	if (!R) return;
	index_release(c, R->index);
	volume_release(c, R->genome);
	delete R;
}

int ref_map(Ctx* c, const RefIndex* R, const mecat_ref_reads* reads, const mecat_ref_params* p, mbref::Sink& out,
            std::vector<int32_t>* dump_counts, std::vector<int32_t>* dump_rows)
{
	if (!R || !reads || !p || !reads->vol) MB_FAIL(c, "ref_map: null argument");
	if (p->tech != 0 && p->tech != 1) MB_FAIL(c, "ref_map: technology (-x) must be 0 (pacbio) or 1 (nanopore), not %d", p->tech);
	const mecat_volume* v = reads->vol;
	for (int32_t r = 0; r < reads->num_reads; ++r) {
		const int32_t f = reads->fwd_read[r], w = reads->rev_read[r];
		if (f < 0 || f >= v->num_reads || w < 0 || w >= v->num_reads || v->offset_size[2 * f + 1] != reads->read_len[r] ||
		    v->offset_size[2 * w + 1] != reads->read_len[r])
			MB_FAIL(c, "ref_map: read %d does not match its volume reads", r);
	}
	if (reads->num_bad < 0 || (reads->num_bad && !reads->bad)) MB_FAIL(c, "ref_map: bad list of other letters");
	for (int64_t k = 1; k < reads->num_bad; ++k)
		if (reads->bad[k - 1] >= reads->bad[k]) MB_FAIL(c, "ref_map: the offsets of other letters must ascend");
	for (int32_t r = 0; r < reads->num_reads && reads->num_bad; ++r) {
		if (!reads->rev_is_rc[r]) continue;
		// a strand taken by reverse complement cannot carry letters the reference treats differently on the two strands
		const int64_t lo = v->offset_size[2 * reads->rev_read[r]], hi = lo + reads->read_len[r];
		const int64_t* it = std::lower_bound(reads->bad, reads->bad + reads->num_bad, lo);
		if (it != reads->bad + reads->num_bad && *it < hi) MB_FAIL(c, "ref_map: read %d has other letters and needs an explicit reverse strand", r);
	}
	DVolume* dv = nullptr;
	if (volume_upload(c, v, &dv)) return 1;
	int64_t* d_bad = nullptr;
	auto body = [&]() -> int {
		MB_CUDA(c, c->dmalloc((void**)&d_bad, sizeof(int64_t) * (size_t)(reads->num_bad + 1)));
		if (reads->num_bad) MB_CUDA(c, cudaMemcpyAsync(d_bad, reads->bad, sizeof(int64_t) * (size_t)reads->num_bad, cudaMemcpyHostToDevice, c->stream));
		MB_CUDA(c, cudaStreamSynchronize(c->stream));
		c->stats.h2d_bytes += (int64_t)sizeof(int64_t) * reads->num_bad;
		mbref::MapIn in;
		in.R = reads->num_reads; in.h_len = reads->read_len; in.h_fread = reads->fwd_read; in.h_rread = reads->rev_read; in.h_rrc = reads->rev_is_rc;
		in.seqcount = R->genome->num_bases;
		in.d_fwd = dv->fwd; in.d_offsz = (const int32_t*)dv->offsz; in.d_bad = d_bad; in.nbad = reads->num_bad;
		in.d_ibegin = R->index->begin; in.d_ipos = R->index->pos;
		mbref::Params P;
		P.num_candidates = p->num_candidates; P.num_output = p->num_output; P.want_strings = p->want_strings != 0;
		P.dump_counts = dump_counts; P.dump_rows = dump_rows;
		P.strings_for_printed_only = p->tech == 0 && RefBackend::forward_only();
		if (const char* e = getenv("MECAT_B200_REF_SEED")) P.seed_per_thread = !strcmp(e, "thread");      // the earlier kernel shape, for cross-checks
		if (const char* e = getenv("MECAT_B200_REF_TABLE_MB")) P.table_budget = (int64_t)atoll(e) << 20;      // test hook: force several table batches
		RefBackend be(c, dv, R->genome, p->tech);
		const int rc = mbref::map_reads(be, in, P, out);
		if (rc) be.end_batch();
		return rc;
	};
	const int rc = body();
	c->dfree(d_bad);
	volume_release(c, dv);
	if (!rc) c->stats.num_records += (int64_t)out.recs.size();
	return rc;
}

}  // namespace mb
