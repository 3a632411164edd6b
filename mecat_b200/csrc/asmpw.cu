// mecat_b200/csrc/asmpw.cu -- mecat2asmpw / mecat2trimpw on the GPU: CUDA backend of
// asm_pipeline.h.  The stage sequence lives in asm_pipeline.h, the per-unit bodies in asm_core.cuh; here every stage
// functor F becomes a launch of k_asm<F> (one thread per text position / k-mer code / strand / read), the alignment of the
// strands and of the candidates launches with a warp per unit (k_asm_seed_warp; k_asm_slots: as many warps as scratch
// slots, each taking candidates in a grid-stride loop), and memory comes from the context's pool.
// Reference: mecat2canu/src/mecat2asmpw/mecat2asmpw.c (creat_ref_index :397-497, pairwise_mapping :515-984).
#include "common.cuh"
#include "dev_backend.cuh"
#include "asm_pipeline.h"

#include <algorithm>
#include <string>

namespace mb {

struct AsmIndexDev { mbasm::AsmIndex I; };

namespace {

template <class F>
__global__ void __launch_bounds__(128) k_asm(const F f, const int64_t n)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) f(i);
}

struct AsmWarpLanes        // the lanes interface of asm_core.cuh on a real warp
{
	__device__ static int lane() { return (int)(threadIdx.x & 31u); }
	template <class F> __device__ void each(F&& f) const { f(lane()); }
	template <class F> __device__ int sum(F&& f) const { return __reduce_add_sync(0xffffffffu, f(lane())); }
	template <class F> __device__ uint32_t ballot(F&& f) const { return __ballot_sync(0xffffffffu, f(lane())); }
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

// a warp per strand: the lanes take the hits of a sampled k-mer's list, the pair tests of a block's entries and the
// entries of a voting neighbour; 1.9 KB of shared memory per warp for the entries being scored
__global__ void __launch_bounds__(SEED_WARPS * 32) k_asm_seed_warp(const mbasm::SeedWarpFn f, const int64_t n)
{
	__shared__ mbasm::WarpScratch scratch[SEED_WARPS];
	const int64_t u = (int64_t)blockIdx.x * SEED_WARPS + (threadIdx.x >> 5);
	if (u < n) f(u, AsmWarpLanes(), scratch[threadIdx.x >> 5]);
}

constexpr int SLOT_WARPS = 4;

// a warp per scratch slot, candidates in a grid-stride loop: the lanes compare 32 letters of a diagonal per step and
// write a run of columns together
template <class F>
__global__ void __launch_bounds__(SLOT_WARPS * 32) k_asm_slots(const F f, const int64_t n, const int64_t nslots)
{
	const int64_t slot = (int64_t)blockIdx.x * SLOT_WARPS + (threadIdx.x >> 5);
	if (slot >= nslots) return;
	for (int64_t i = slot; i < n; i += nslots) f(i, (int)slot, AsmWarpLanes());
}

struct AsmBackend : PoolBackend
{
	int64_t budget; int divisor, slots_per_sm;
	AsmBackend(Ctx* ctx) : PoolBackend(ctx, "asm", true), budget(0), divisor(1), slots_per_sm(32)
	{
		// block tables and record pool of the strands in flight: 4 GB hold ~10 000 strands of 4 kb reads at 32x, several
		// resident warps per scheduler, and are allocated once (a first cudaMalloc of 12 GB cost the command line 0.5 s)
		size_t free_b = 0, total_b = 0;
		budget = (int64_t)4 << 30;
		if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) budget = std::min<int64_t>(budget, (int64_t)(free_b / 5 * 2));
		if (budget < ((int64_t)64 << 20)) budget = (int64_t)64 << 20;
		if (const char* e = getenv("MECAT_B200_ASM_TABLE_MB")) budget = std::max<int64_t>(1, atoll(e)) << 20;      // test hook: force several batches
		if (const char* e = getenv("MECAT_B200_ASM_POOL_DIV")) divisor = std::max(1, atoi(e));                    // test hook: a pool that runs out
		if (const char* e = getenv("MECAT_B200_ASM_EXTEND_WARPS")) slots_per_sm = std::max(1, atoi(e));           // alignment warps per SM
	}
	template <class F> bool launch(int64_t n, const F& f, int stage)
	{
		if (n <= 0) return true;
		KScope ks(c, MECAT_K_ASM_INDEX + stage);
		k_asm<F><<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(f, n);
		return check(cudaGetLastError(), "launch");
	}
	bool launch_seed(int64_t n, const mbasm::SeedWarpFn& f, int stage)
	{
		if (n <= 0) return true;
		KScope ks(c, MECAT_K_ASM_INDEX + stage);
		k_asm_seed_warp<<<(unsigned)((n + SEED_WARPS - 1) / SEED_WARPS), SEED_WARPS * 32, 0, c->stream>>>(f, n);
		return check(cudaGetLastError(), "launch");
	}
	template <class F> bool launch_slots(int64_t n, const F& f, int64_t nslots, int stage)
	{
		if (n <= 0 || nslots <= 0) return true;
		KScope ks(c, MECAT_K_ASM_INDEX + stage);
		k_asm_slots<F><<<(unsigned)((nslots + SLOT_WARPS - 1) / SLOT_WARPS), SLOT_WARPS * 32, 0, c->stream>>>(f, n, nslots);
		return check(cudaGetLastError(), "launch");
	}
	void keep(void* p)
	{
		for (size_t i = 0; i < owned.size(); ++i) if (owned[i] == p) { owned[i] = owned.back(); owned.pop_back(); return; }
	}
	int64_t table_budget() const { return budget; }
	int pool_divisor() const { return divisor; }
	int64_t extend_slots() const { return (int64_t)c->sm_count * slots_per_sm; }      // alignment warps in flight, each with its own scratch
};

// a file of reads as the ABI describes it: every read inside the text with a NUL behind it, in order, shorter than RM
bool check_reads(Ctx* c, const mecat_asm_reads* r, const char* what)
{
	char b[256];
	if (!r->text || r->num_letters <= 0 || r->num_reads <= 0 || !r->read_start || !r->read_len) {
		snprintf(b, sizeof b, "%s: empty read set", what); c->err = b; return false;
	}
	for (int32_t i = 0; i < r->num_reads; ++i) {
		const int64_t s = r->read_start[i], l = r->read_len[i];
		if (s < 0 || l < 0 || s + l >= r->num_letters || r->text[s + l] != 0 || (i && s < (int64_t)r->read_start[i - 1] + r->read_len[i - 1] + 1)) {
			snprintf(b, sizeof b, "%s: read %d does not lie in the text with a NUL behind it", what, i); c->err = b; return false;
		}
		if (l >= mbasm::MAX_READ) {
			snprintf(b, sizeof b, "%s: read %d has %lld letters; the reference's buffers hold fewer than 100 000 (RM, mecat2asmpw.c:17)", what, i, (long long)l);
			c->err = b; return false;
		}
	}
	return true;
}

}  // namespace

int asm_index_build(Ctx* c, const mecat_asm_reads* subject, AsmIndexDev** out)
{
	if (!check_reads(c, subject, "asm_index_build")) return 1;
	AsmIndexDev* D = new AsmIndexDev;
	AsmBackend be(c);
	const bool ok = mbasm::index_build(be, subject->text, subject->num_letters, subject->read_start, subject->read_len, subject->num_reads,
	                                   subject->first_read_id, D->I);
	be.end_batch();          // after a failed build this frees the index arrays too: they are kept only at its end
	if (!ok) { delete D; return 1; }
	c->stats.index_kmers += D->I.total;
	c->stats.index_bases += subject->num_letters;
	*out = D;
	return 0;
}

void asm_index_release(Ctx* c, AsmIndexDev* D)
{
	if (!D) return;
	c->dfree(D->I.text); c->dfree(D->I.start); c->dfree(D->I.len); c->dfree(D->I.begin); c->dfree(D->I.pos);
	delete D;
}

int asm_index_export(Ctx* c, const AsmIndexDev* D, int64_t* num_positions, uint32_t* begin, int32_t* positions)
{
	*num_positions = D->I.total;
	if (begin) MB_CUDA(c, cudaMemcpyAsync(begin, D->I.begin, sizeof(uint32_t) * (size_t)(mbasm::KMERS + 1), cudaMemcpyDeviceToHost, c->stream));
	if (positions && D->I.total) MB_CUDA(c, cudaMemcpyAsync(positions, D->I.pos, sizeof(int32_t) * (size_t)D->I.total, cudaMemcpyDeviceToHost, c->stream));
	MB_CUDA(c, cudaStreamSynchronize(c->stream));
	return 0;
}

int asm_overlaps(Ctx* c, const AsmIndexDev* D, const mecat_asm_reads* query, const mecat_asm_params* p, mecat_asm_overlap** out, size_t* n)
{
	static_assert(sizeof(mbasm::Overlap) == sizeof(mecat_asm_overlap), "Overlap mirrors mecat_asm_overlap");
	if (!check_reads(c, query, "asm_overlaps")) return 1;
	AsmBackend be(c);
	std::vector<mbasm::Overlap> recs;
	mbasm::Counters cnt;
	const bool ok = mbasm::overlaps(be, D->I, query->text, query->num_letters, query->read_start, query->read_len, query->num_reads, query->first_read_id,
	                                p->variant, p->max_candidates, recs, &cnt);
	be.end_batch();
	if (!ok) return 1;
	c->stats.num_hits += cnt.hits; c->stats.num_candidates += cnt.candidates; c->stats.num_records += (int64_t)recs.size();
	mecat_asm_overlap* res = (mecat_asm_overlap*)malloc(sizeof(mecat_asm_overlap) * (recs.size() ? recs.size() : 1));
	if (!res) MB_FAIL(c, "asm_overlaps: out of host memory");
	if (!recs.empty()) memcpy(res, recs.data(), sizeof(mecat_asm_overlap) * recs.size());
	*out = res; *n = recs.size();
	return 0;
}

}  // namespace mb
