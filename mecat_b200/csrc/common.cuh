// mecat_b200/csrc/common.cuh -- shared declarations of the device library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mecat_b200.h"

namespace mb {

// ---- algorithm constants of the path (reference src/mecat2pw/pw_impl.h:10-19, pw_impl.cpp:20)
constexpr int KMER = 13;            // kmer_size
constexpr int STRIDE = 10;          // BC: every 10th k-mer of the query is looked up
constexpr int SLOTS = 40;           // SM: seeds kept per bucket
constexpr int SEGW = 2000;          // ZV: bucket width in index-volume bases
constexpr int MAX_OCC = 128;        // lookup_table.cpp:97
constexpr uint32_t NCODES = 1u << (2 * KMER);

struct Ctx;

// Device-resident volume.  `fwd` holds base i at bits 2*(i%16) of word i/16 (LSB first, so a
// walk over increasing positions is a funnel shift away); `rev` holds the same bases in
// reverse order (rev base i = base N-1-i), which turns the left extension and the reverse
// strand into forward walks too.  Both arrays carry 8 zero words of slack at the end.
struct DVolume
{
	int32_t num_reads = 0, num_bases = 0, start_read_id = 0;
	int2* offsz = nullptr;          // {offset,size} per read
	uint32_t* fwd = nullptr;
	uint32_t* rev = nullptr;
	size_t words = 0;
	int32_t max_read = 0;
	std::vector<int32_t> h_offsz;   // host copy (record assembly)
};

struct DIndex
{
	uint32_t* begin = nullptr;      // NCODES + 1 (CSR over kept k-mers)
	int32_t* pos = nullptr;         // kept k-mer start positions, ascending inside a list
	uint32_t* counts = nullptr;     // histogram / fill cursors; only alive between the two build stages
	int64_t num_kmers = 0;
};

struct Ctx
{
	int device = 0;
	int sm_count = 148;
	cudaStream_t stream = nullptr;
	std::string err;
	mecat_b200_stats stats;
	unsigned long long* d_counters = nullptr;   // 16 device counters (statistics, arena cursors)
	size_t align_arena = 3ull << 30;            // bytes per column arena of one extension-with-strings batch (two arenas); set at init
	// Device memory pool: the per-tile buffers have the same sizes call after call, so freed
	// blocks are kept and handed out again (cudaMalloc / cudaFree of multi-GB blocks cost
	// milliseconds each).  Everything is returned to the driver by trim() / destroy.
	struct Block { void* p; size_t bytes; bool used; };
	std::vector<Block> blocks;
	cudaError_t dmalloc(void** out, size_t bytes)
	{
		if (bytes == 0) bytes = 256;
		int best = -1;
		for (size_t i = 0; i < blocks.size(); ++i)
			if (!blocks[i].used && blocks[i].bytes >= bytes && blocks[i].bytes <= bytes + bytes / 4 + (1u << 20) &&
			    (best < 0 || blocks[i].bytes < blocks[best].bytes)) best = (int)i;
		if (best >= 0) { blocks[best].used = true; *out = blocks[best].p; return cudaSuccess; }
		void* p = nullptr;
		cudaError_t e = cudaMalloc(&p, bytes);
		if (e != cudaSuccess) {
			trim();                       // give cached blocks back and retry once
			(void)cudaGetLastError();
			e = cudaMalloc(&p, bytes);
			if (e != cudaSuccess) return e;
		}
		blocks.push_back({p, bytes, true});
		*out = p;
		return cudaSuccess;
	}
	void dfree(void* p)
	{
		if (!p) return;
		for (auto& b : blocks) if (b.p == p) { b.used = false; return; }
		cudaFree(p);
	}
	void trim()
	{
		size_t k = 0;
		for (size_t i = 0; i < blocks.size(); ++i) {
			if (blocks[i].used) blocks[k++] = blocks[i];
			else cudaFree(blocks[i].p);
		}
		blocks.resize(k);
	}
	template <typename T> cudaError_t alloc(T** out, size_t count) { return dmalloc((void**)out, count * sizeof(T)); }

	// Pinned host staging buffers for the per-tile result copies: grow-only, reused tile after tile
	// (a fresh pageable vector of ~70 MB costs its zero fill and page faults on every tile, and
	// pageable D2H copies are staged by the driver).
	struct HostStage { void* p = nullptr; size_t bytes = 0; };
	HostStage hstage[4];
	cudaError_t host_stage(int slot, size_t bytes, void** out)
	{
		HostStage& h = hstage[slot];
		if (h.bytes < bytes) {
			if (h.p) cudaFreeHost(h.p);
			h.p = nullptr; h.bytes = 0;
			const size_t want = bytes + bytes / 8 + 4096;
			cudaError_t e = cudaHostAlloc(&h.p, want, cudaHostAllocDefault);
			if (e != cudaSuccess) { h.p = nullptr; return e; }
			h.bytes = want;
		}
		*out = h.p;
		return cudaSuccess;
	}

	// Pinned host blocks handed OUT to the caller as results (records, text) and taken back by mecat_b200_free /
	// mecat_b200_host_free: a D2H copy into fresh pageable memory costs its page faults and a staged copy (100 MB of M4
	// records: ~30 ms against ~2 ms), and pinning a fresh block per tile costs as much, so freed blocks are kept
	// (capi.cu: host_out_alloc / host_out_release; a process-wide registry, because the binding frees without a context).
	struct HostOut { void* p; size_t bytes; bool used; };
	std::vector<HostOut> host_out;

	// per-kernel CUDA-event timing on `stream`
	struct Pending { int slot; cudaEvent_t a, b; };
	std::vector<Pending> pending;
	std::vector<cudaEvent_t> pool;
	cudaEvent_t get_event()
	{
		if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
		cudaEvent_t e;
		cudaEventCreate(&e);
		return e;
	}
	// call after the stream has been synchronised
	void resolve_timers()
	{
		for (auto& p : pending) {
			float ms = 0;
			if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) stats.kernel_ms[p.slot] += ms;
			pool.push_back(p.a); pool.push_back(p.b);
		}
		pending.clear();
	}
};

// times everything enqueued on the context stream during its lifetime into stats.kernel_ms[slot]
struct KScope
{
	Ctx* c; int slot; cudaEvent_t a; int launches;
	KScope(Ctx* c_, int slot_, int launches_ = 1) : c(c_), slot(slot_), launches(launches_) { a = c->get_event(); cudaEventRecord(a, c->stream); }
	~KScope()
	{
		cudaEvent_t b = c->get_event();
		cudaEventRecord(b, c->stream);
		c->pending.push_back({slot, a, b});
		c->stats.kernel_launches[slot] += launches;
	}
};

#define MB_CUDA(ctx, call)                                                                      \
	do {                                                                                        \
		cudaError_t e__ = (call);                                                               \
		if (e__ != cudaSuccess) {                                                               \
			char b__[512];                                                                      \
			snprintf(b__, sizeof b__, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
			(ctx)->err = b__;                                                                   \
			return 1;                                                                           \
		}                                                                                       \
	} while (0)

#define MB_FAIL(ctx, ...)                                  \
	do {                                                   \
		char b__[512];                                     \
		snprintf(b__, sizeof b__, __VA_ARGS__);            \
		(ctx)->err = b__;                                  \
		return 1;                                          \
	} while (0)

// ---- internal entry points (one per translation unit)
void* host_out_alloc(Ctx* c, size_t bytes);      // nullptr when out of memory
bool host_out_release(void* p);                   // false: not one of these blocks (a malloc'ed result)
int volume_upload(Ctx* c, const mecat_volume* v, DVolume** out);
void volume_release(Ctx* c, DVolume* v);
// a volume packed on the device from the letters of its reads (volume.cu: k_pack_text); pac_out (optional, host): the
// reference's packed bytes, for the volume file
int volume_from_text(Ctx* c, const char* text, size_t text_bytes, const int64_t* src_off, const int32_t* h_offsz, int num_reads,
                     int num_bases, int start_read_id, uint8_t* pac_out, DVolume** out);
// working volume made of n reads taken from resident volumes: read i = read h_src_read[i] of src[h_src_vol[i]]
int volume_gather(Ctx* c, const DVolume* const* src, const int32_t* h_src_vol, const int32_t* h_src_read, int n, DVolume** out);
int index_build(Ctx* c, const DVolume* v, DIndex** out);
int index_count_part(Ctx* c, const DVolume* v, uint32_t code_lo, uint32_t code_hi, DIndex** out);
int index_finish_part(Ctx* c, const DVolume* v, DIndex* I, uint32_t code_lo, uint32_t code_hi);
void index_release(Ctx* c, DIndex* i);

struct ExtendTask          // device-side extension request (global array)
{
	int32_t qread, qstrand, qstart, sread, sstart;
};
struct ExtendHalf          // result of one direction of one task
{
	int32_t cols, matches, qadv, tadv;
};
int extend_launch(Ctx* c, const DVolume* q, const DVolume* s, const ExtendTask* d_tasks, size_t ntasks,
                  ExtendHalf* d_halves /* 2*ntasks: left then right */);

struct AlignTask           // = mecat_align_task
{
	int32_t qread, qstrand, qstart, sread, sstart, swin_off, swin_len;
};
struct AlnSlot                 // where one (task, direction) writes its columns, and what it produced
{
	unsigned long long off;    // byte offset of the slot in both column arenas
	int32_t cap;               // capacity in columns
	int32_t cols, matches, qadv, tadv;
	int32_t overflow;          // bit 0: the slot was too small; bits 1-3 (nanopore extension): flags of the last column, xdrop_core.cuh
};
// nanopore (-x 1) extension, xdrop.cu: fills the slots (and, with column arenas, the columns) of 2 * nb chains ...
int xdrop_fill_slots(Ctx* c, const DVolume* q, const DVolume* s, const AlignTask* d_tasks, size_t nb, AlnSlot* d_slots,
                     char* d_colq, char* d_colt);
// ... and the string-free form for mecat2pw -j 1: XdropAligner::go's accessors per task
int xdrop_extend(Ctx* c, const DVolume* q, const DVolume* s, const ExtendTask* d_tasks, size_t ntasks, int min_aln,
                 mecat_extend_result* d_res);

// results of one arena batch of extensions with strings, still in device memory
struct AlignDev
{
	int32_t* d_info = nullptr;                  // 8 ints per task {ok, qstart, qend, sstart, send, columns, matches, -}
	char* d_packq = nullptr;                    // gapped strings, task t at d_outoff[t], NUL after each
	char* d_packt = nullptr;
	unsigned long long* d_outoff = nullptr;     // ntasks + 1
	size_t total = 0;
};
size_t align_task_columns(const DVolume* q, const DVolume* s, const AlignTask& t);
int align_batch_device(Ctx* c, int policy, double err, const DVolume* q, const DVolume* s, const AlignTask* h_tasks, size_t nb,
                       int min_aln, AlignDev* out, std::vector<int32_t>& info);
void align_dev_release(Ctx* c, AlignDev* d);
// want_strings = false: only the results come back (coordinates, columns, matches); the strings stay on the device
int align_batch(Ctx* c, int policy, double err, const DVolume* q, const DVolume* s, const AlignTask* h_tasks, size_t ntasks,
                int min_aln, mecat_align_result* h_results, std::vector<char>& qstr, std::vector<char>& sstr, bool want_strings = true);

}  // namespace mb
namespace mbcns { struct BatchIn; struct Params; }
namespace mb {
// Corrected pieces as the C ABI hands them out: records + one malloc'ed blob of bases that grows by doubling, so a
// piece is copied exactly once on its way from the pinned staging buffer to the caller.
struct CnsBlob
{
	std::vector<mecat_cns_piece> recs;
	char* buf = nullptr;
	size_t len = 0, cap = 0;
	bool oom = false;
	~CnsBlob() { free(buf); }
	void reserve(size_t want)      // room for `want` more bytes in one step (a blob that doubles its way to 400 MB copies as much again)
	{
		if (len + want + 1 <= cap) return;
		char* nb = (char*)realloc(buf, len + want + 1);
		if (nb) { buf = nb; cap = len + want + 1; }
	}
	void add(int64_t id, int64_t beg, int64_t end, const char* seq, size_t n)
	{
		if (len + n + 1 > cap) {
			size_t want = cap ? cap * 2 : (size_t)1 << 20;
			while (want < len + n + 1) want *= 2;
			char* nb = (char*)realloc(buf, want);
			if (!nb) { oom = true; return; }
			buf = nb; cap = want;
		}
		memcpy(buf + len, seq, n);
		mecat_cns_piece r;
		r.id = id; r.beg = beg; r.end = end; r.seq_offset = (int64_t)len; r.seq_len = (int64_t)n;
		recs.push_back(r);
		len += n;
	}
};
// consensus stage of mecat2cns on the extension results of one batch (cns.cu)
int cns_consensus_device(Ctx* c, const mbcns::BatchIn& in, const mbcns::Params& P, CnsBlob& out);

int device_exclusive_scan(Ctx* c, const int32_t* d_in, int64_t* d_out /* n + 1 */, int64_t n, int64_t* h_total);

// records.cu: A12 on the device (fill_m4record + append_m4v per read) and the text of the result files.
// m4_assemble: per query read r the candidates [h_outpos[r], h_outpos[r+1]) with their extension results; leaves the kept
// records in *d_m4 (pool memory, release with dfree), read by read in the reference's order.
int m4_assemble(Ctx* c, const DVolume* reads, const DVolume* ref, const ExtendTask* d_tasks, const mecat_extend_result* d_res,
                const int32_t* d_scores, const int64_t* d_outpos, int nreads, size_t total, mecat_m4** d_m4, size_t* nout);
// kind 0: mecat_candidate -> `.can` lines, kind 1: mecat_m4 -> `.m4` lines (gapped: with the two extension points).
// *d_text is pool memory (release with dfree).
int records_text_device(Ctx* c, int kind, int gapped, const void* d_records, size_t n, char** d_text, size_t* bytes);

// mecat2ref (refmap.cu): the genome as a one-read volume plus its k-mer index
struct RefIndex
{
	DVolume* genome = nullptr;
	DIndex* index = nullptr;
};
}  // namespace mb
namespace mbref { struct Sink; }
namespace mb {
int ref_index_build(Ctx* c, const mecat_ref_genome* g, RefIndex** out);
void ref_index_release(Ctx* c, RefIndex* R);
int ref_map(Ctx* c, const RefIndex* R, const mecat_ref_reads* reads, const mecat_ref_params* p, mbref::Sink& out,
            std::vector<int32_t>* dump_counts = nullptr, std::vector<int32_t>* dump_rows = nullptr);

struct AsmIndexDev;
int asm_index_build(Ctx* c, const mecat_asm_reads* subject, AsmIndexDev** out);
void asm_index_release(Ctx* c, AsmIndexDev* I);
int asm_index_export(Ctx* c, const AsmIndexDev* I, int64_t* num_positions, uint32_t* begin, int32_t* positions);
int asm_overlaps(Ctx* c, const AsmIndexDev* I, const mecat_asm_reads* query, const mecat_asm_params* p, mecat_asm_overlap** out, size_t* n);

struct RawCand             // candidate_save, pw_impl.h:21-25
{
	int32_t loc1, loc2, left1, left2, right1, right2, score, num1, num2, readno, readstart, chain;
};
// Seeding + scoring + candidate selection for every read of `reads`; fills d_cands
// (num_reads x maxc RawCand, the per-read list in reference order) and d_counts.
int seed_candidates(Ctx* c, const DIndex* idx, const DVolume* ref, const DVolume* reads,
                    const mecat_pw_params* p, int read_begin, int read_end, RawCand* d_cands, int32_t* d_counts);
int volume_from_device(Ctx* c, int num_reads, int num_bases, int start_read_id, const int32_t* h_offsz,
                       const uint8_t* d_pac, DVolume** out);

// ---- device helpers
__device__ __forceinline__ uint32_t ld_bases32(const uint32_t* __restrict__ a, uint32_t base)
{
	// 16 bases starting at `base`, LSB first
	uint32_t w = base >> 4, sh = (base & 15u) << 1;
	uint32_t lo = __ldg(a + w), hi = __ldg(a + w + 1);
	return __funnelshift_r(lo, hi, sh);
}

__device__ __forceinline__ uint32_t rev_groups2(uint32_t x)
{
	// reverse the order of the sixteen 2-bit groups of x
	x = __brev(x);
	return ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
}

// DDF consistency |dloc/(dseed*10) - 1| < 0.25 in exact integer form.  The reference
// evaluates it in float32 (pw_impl.cpp:135,165) or float64 (:412,429); for the operand
// ranges of this path the rounded quotient can never cross 0.75 / 1.25,
// so 15*b < 2*a < 25*b (b > 0), 25*b < 2*a < 15*b (b < 0), false for b == 0 is identical.
__host__ __device__ __forceinline__ bool ddf_close(int a, int b)
{
	int a2 = 2 * a;
	if (b > 0) return 15 * b < a2 && a2 < 25 * b;
	if (b < 0) return 25 * b < a2 && a2 < 15 * b;
	return false;
}

}  // namespace mb
