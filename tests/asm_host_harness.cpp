// tests/asm_host_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the product's mecat2asmpw / mecat2trimpw stage sequence (mecat_b200/csrc/asm_pipeline.h) and per-unit bodies
// (asm_core.cuh) on the host: every stage functor the CUDA backend launches as a kernel is called here in a loop, on
// memory that is never zero by luck, so that the CPU test-suite can compare the statements the GPU executes with the
// oracle and with the unmodified binaries' goldens.  Compiled by tests/util.py; never part of the product library.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../mecat_b200/csrc/asm_pipeline.h"

namespace {

struct HostBackend
{
	std::vector<void*> owned;
	std::string err;
	int64_t budget = (int64_t)1 << 30;
	int divisor = 1;
	int64_t slots = 3;

	template <class T> T* alloc(size_t n)
	{
		void* p = malloc((n ? n : 1) * sizeof(T));
		if (!p) { err = "out of memory"; return nullptr; }
		memset(p, 0xAB, (n ? n : 1) * sizeof(T));
		owned.push_back(p);
		return (T*)p;
	}
	template <class T> bool upload(T* d, const T* h, size_t n) { if (n) memcpy(d, h, n * sizeof(T)); return true; }
	template <class T> bool download(T* h, const T* d, size_t n) { if (n) memcpy(h, d, n * sizeof(T)); return true; }
	bool fill(void* d, int byte, size_t bytes) { memset(d, byte, bytes); return true; }
	template <class F> bool launch(int64_t n, const F& f, int) { for (int64_t i = 0; i < n; ++i) f(i); return true; }
	// the warp form of the seeding functor with the host lanes; MECAT_B200_ASM_SEED=thread: the one-thread form of the same work
	bool launch_seed(int64_t n, const mbasm::SeedWarpFn& f, int)
	{
		const char* e = getenv("MECAT_B200_ASM_SEED");
		if (e && !strcmp(e, "thread")) {
			mbasm::SeedFn g;
			g.q = f.q; g.sub = f.sub; g.units = f.units; g.begin = f.begin; g.pos = f.pos; g.tab = f.tab; g.gate = f.gate; g.maxc = f.maxc;
			g.cands = f.cands; g.ncand = f.ncand; g.status = f.status;
			for (int64_t i = 0; i < n; ++i) g(i);
			return true;
		}
		mbasm::WarpScratch W;
		memset(&W, 0xAB, sizeof W);
		for (int64_t i = 0; i < n; ++i) f(i, mbasm::EmuLanes(), W);
		return true;
	}
	template <class F> bool launch_slots(int64_t n, const F& f, int64_t nslots, int)
	{
		for (int64_t i = 0; i < n; ++i) f(i, (int)(i % nslots), mbasm::EmuLanes());
		return true;
	}
	bool release(void* p)
	{
		for (size_t i = 0; i < owned.size(); ++i) if (owned[i] == p) { free(p); owned[i] = owned.back(); owned.pop_back(); return true; }
		err = "release of an unknown block";
		return false;
	}
	void keep(void* p) { for (size_t i = 0; i < owned.size(); ++i) if (owned[i] == p) { owned[i] = owned.back(); owned.pop_back(); return; } }
	int64_t table_budget() const { return budget; }
	int pool_divisor() const { return divisor; }
	int64_t extend_slots() const { return slots; }
	void fail(const char* m) { err = m; }
};

}  // namespace

// One subject file, one query file.  budget / divisor: 0 = defaults; small values force the batch cuts and the pool
// splits.  stats: seed batches, candidates, hits, extension passes.
extern "C" int ah_overlaps(const char* text, int64_t n, const int32_t* starts, const int32_t* lens, int32_t nreads, int32_t first_id,
                           const char* qtext, int64_t qn, const int32_t* qstarts, const int32_t* qlens, int32_t nq, int32_t qfirst,
                           int variant, int maxc, int64_t budget, int divisor, void** out, size_t* nout, int64_t* stats, char* err, int errcap)
{
	using namespace mbasm;
	HostBackend be;
	if (budget > 0) be.budget = budget;
	if (divisor > 0) be.divisor = divisor;
	AsmIndex I;
	std::vector<Overlap> recs;
	Counters cnt;
	bool ok = index_build(be, text, n, starts, lens, nreads, first_id, I) &&
	          overlaps(be, I, qtext, qn, qstarts, qlens, nq, qfirst, variant, maxc, recs, &cnt);
	if (ok && !be.owned.empty()) { be.err = "the pipeline left blocks behind"; ok = false; }
	for (void* p : be.owned) free(p);
	free(I.text); free(I.start); free(I.len); free(I.begin); free(I.pos);
	if (!ok) { if (err && errcap > 0) { strncpy(err, be.err.c_str(), (size_t)errcap - 1); err[errcap - 1] = 0; } return 1; }
	*nout = recs.size();
	*out = malloc(recs.size() * sizeof(Overlap) + 1);
	memcpy(*out, recs.data(), recs.size() * sizeof(Overlap));
	if (stats) { stats[0] = cnt.seed_batches; stats[1] = cnt.candidates; stats[2] = cnt.hits; stats[3] = cnt.extend_passes; }
	return 0;
}

// the k-mer lists of a text: begin (4^13 + 1 entries) and positions (caller's buffers; returns the number of positions)
extern "C" int64_t ah_index(const char* text, int64_t n, const int32_t* starts, const int32_t* lens, int32_t nreads, uint32_t* begin, int32_t* pos, int64_t cap)
{
	using namespace mbasm;
	HostBackend be;
	AsmIndex I;
	if (!index_build(be, text, n, starts, lens, nreads, 1, I)) return -1;
	const int64_t total = I.total;
	memcpy(begin, I.begin, sizeof(uint32_t) * (size_t)(KMERS + 1));
	if (total <= cap) memcpy(pos, I.pos, sizeof(int32_t) * (size_t)total);
	free(I.text); free(I.start); free(I.len); free(I.begin); free(I.pos);
	for (void* p : be.owned) free(p);
	return total;
}

extern "C" void ah_free(void* p) { free(p); }

// the integer consistency test against the reference's float and double forms (tests/test_asm_host.py)
extern "C" int ah_ddf_close(int a, int b, int wide) { return (wide ? mbasm::ddf_close_d(a, b) : mbasm::ddf_close(a, b)) ? 1 : 0; }
