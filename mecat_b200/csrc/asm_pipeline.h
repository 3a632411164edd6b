// mecat_b200/csrc/asm_pipeline.h -- stage sequence of mecat2asmpw / mecat2trimpw over a
// backend: asmpw.cu runs every stage functor of asm_core.cuh as a kernel on device memory, tests/asm_host_harness.cpp
// runs the same functors in a loop on the host.  Reference: mecat2canu/src/mecat2asmpw/mecat2asmpw.c (main :1063-1166
// builds the index of one file and maps the reads of that file and of the following ones against it).
//
// Backend: alloc<T>(n) / release(p) / keep(p) / upload / download / fill / launch(n, f, stage) / launch_seed(n, f, stage) (a warp per unit) /
// launch_slots(n, f, slots, stage) (f(i, slot) with `slots` units in flight) / table_budget() / extend_slots() / fail(msg).
#pragma once
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "asm_core.cuh"

namespace mbasm {

enum { ST_INDEX = 0, ST_SEED = 1, ST_EXTEND = 2 };

struct AsmIndex            // device-side subject file: text, read table, k-mer lists
{
	char* text = nullptr; int64_t n = 0;
	int32_t* start = nullptr; int32_t* len = nullptr; int32_t nreads = 0, first_id = 0, max_len = 0;
	uint32_t* begin = nullptr; int32_t* pos = nullptr; uint32_t total = 0;
	Reads reads() const { Reads r; r.text = text; r.start = start; r.len = len; r.n = nreads; r.first_id = first_id; return r; }
};

// creat_ref_index :397-497
template <class B>
bool index_build(B& be, const char* h_text, int64_t n, const int32_t* h_start, const int32_t* h_len, int32_t nreads, int32_t first_id, AsmIndex& I)
{
	if (n <= 0 || n >= 0x7fffffff - 4 * ZV || nreads <= 0) { be.fail("asm: subject text must hold 1 .. 2^31 letters"); return false; }
	I.n = n; I.nreads = nreads; I.first_id = first_id; I.max_len = 0;
	for (int32_t r = 0; r < nreads; ++r) I.max_len = std::max(I.max_len, h_len[r]);
	I.text = be.template alloc<char>((size_t)n + 16);
	I.start = be.template alloc<int32_t>((size_t)nreads);
	I.len = be.template alloc<int32_t>((size_t)nreads);
	I.begin = be.template alloc<uint32_t>((size_t)KMERS + 1);
	uint32_t* count = be.template alloc<uint32_t>((size_t)KMERS);
	const int64_t ntiles = KMERS / TILE;
	uint32_t* tile_sum = be.template alloc<uint32_t>((size_t)ntiles + 1);
	if (!I.text || !I.start || !I.len || !I.begin || !count || !tile_sum) return false;
	if (!be.fill(I.text, 0, (size_t)n + 16) || !be.upload(I.text, h_text, (size_t)n) || !be.upload(I.start, h_start, (size_t)nreads) ||
	    !be.upload(I.len, h_len, (size_t)nreads) || !be.fill(count, 0, sizeof(uint32_t) * (size_t)KMERS)) return false;
	UpperFn fu; fu.text = I.text; fu.n = n;
	if (!be.launch((n + 15) / 16, fu, ST_INDEX)) return false;
	KmerCountFn fc; fc.text = I.text; fc.n = n; fc.count = count;
	if (!be.launch(n, fc, ST_INDEX)) return false;
	TileSumFn fs; fs.count = count; fs.tile_sum = tile_sum;
	if (!be.launch(ntiles, fs, ST_INDEX)) return false;
	{   // exclusive scan of the 2^20 tile sums on the host (4 MB each way)
		std::vector<uint32_t> h((size_t)ntiles + 1);
		if (!be.download(h.data(), tile_sum, (size_t)ntiles)) return false;
		uint32_t run = 0;
		for (int64_t t = 0; t < ntiles; ++t) { const uint32_t v = h[(size_t)t]; h[(size_t)t] = run; run += v; }
		h[(size_t)ntiles] = run;
		I.total = run;
		if (!be.upload(tile_sum, h.data(), (size_t)ntiles + 1)) return false;
	}
	TileScanFn fd; fd.count = count; fd.tile_sum = tile_sum; fd.begin = I.begin; fd.total = tile_sum + ntiles;
	if (!be.launch(ntiles, fd, ST_INDEX)) return false;
	I.pos = be.template alloc<int32_t>((size_t)I.total + 1);
	if (!I.pos) return false;
	KmerFillFn ff; ff.text = I.text; ff.n = n; ff.begin = I.begin; ff.cursor = count; ff.pos = I.pos;
	if (!be.launch(n, ff, ST_INDEX)) return false;
	ListSortFn fl; fl.begin = I.begin; fl.pos = I.pos;
	if (!be.launch(KMERS, fl, ST_INDEX)) return false;
	if (!be.release(count) || !be.release(tile_sum)) return false;
	be.keep(I.text); be.keep(I.start); be.keep(I.len); be.keep(I.begin); be.keep(I.pos);
	return true;
}

struct QuerySet            // device-side query reads of one call
{
	Reads q; const int32_t* h_len; int32_t max_len;
};

struct SeedState           // arrays over all strands of the call
{
	Cand* cands; int32_t* ncand; std::vector<int32_t> cap, hits;      // cap: records a strand's table can need at most
};

// room for a strand's table.  Its hits bound the blocks it can touch, but far fewer are touched: a chance hit makes a
// block of its own (expect text / 4^13 of them per sampled k-mer), a true overlap puts ~60 seeds into each of its blocks.
// `mult` doubles every time a range ran out and was split.
inline int64_t table_room(int64_t bound, int64_t hits, int len, int64_t text, int mult)
{
	const int64_t nk = sampled_kmers(len);
	const int64_t guess = (3 * nk * text / KMERS) / 2 + hits / 16 + 64;
	return std::min<int64_t>(bound, guess * mult);
}
// room in the second pool (full entry arrays): blocks that collect more than one k-mer -- those of true overlaps
inline int64_t entries_room(int64_t bound, int64_t hits, int mult) { return std::min<int64_t>(bound, (hits / 16 + 64) * mult); }
inline uint32_t slots_for(int64_t room) { uint32_t sl = 2; while ((int64_t)sl < 2 * room) sl <<= 1; return sl; }

// strands units[lo, hi): block tables in one allocation, records from a shared pool.  A table or a pool that runs out
// splits the range and doubles the room; a strand on its own gets its bound.
template <class B>
bool seed_range(B& be, const AsmIndex& I, const QuerySet& Q, SeedState& S, const std::vector<int32_t>& units, size_t lo, size_t hi, int gate, int maxc,
                int mult, int64_t* batches)
{
	const size_t n = hi - lo;
	if (!n) return true;
	if (n == 1) mult = 1 << 20;
	std::vector<int64_t> slot_off(n + 1, 0), list_off(n + 1, 0);
	int64_t pool_cap = 0, big_cap = 0;
	for (size_t i = 0; i < n; ++i) {
		const int32_t u = units[lo + i];
		const int64_t c = table_room(S.cap[(size_t)u], S.hits[(size_t)u], Q.h_len[u >> 1], I.n, mult);
		slot_off[i + 1] = slot_off[i] + slots_for(c); list_off[i + 1] = list_off[i] + c;
		pool_cap += c;
		big_cap += entries_room(S.cap[(size_t)u], S.hits[(size_t)u], mult);
	}
	if (big_cap < 1) big_cap = 1;
	if (big_cap > 0x7fffffff) big_cap = 0x7fffffff;
	if (n > 1) pool_cap = pool_cap / be.pool_divisor() + 64;      // few strands fill their room
	if (pool_cap < 1) pool_cap = 1;
	if (pool_cap > 0x7fffffff) pool_cap = 0x7fffffff;
	Slot* slots = be.template alloc<Slot>((size_t)slot_off[n]);
	int32_t* lists = be.template alloc<int32_t>((size_t)list_off[n]);
	Bucket* pool = be.template alloc<Bucket>((size_t)pool_cap);
	Entries* big = be.template alloc<Entries>((size_t)big_cap);
	int64_t* d_slot_off = be.template alloc<int64_t>(n + 1);
	int64_t* d_list_off = be.template alloc<int64_t>(n + 1);
	int32_t* d_units = be.template alloc<int32_t>(n);
	int32_t* d_status = be.template alloc<int32_t>(n);
	uint32_t* d_used = be.template alloc<uint32_t>(2);
	if (!slots || !lists || !pool || !big || !d_slot_off || !d_list_off || !d_units || !d_status || !d_used) return false;
	if (!be.fill(slots, 0, sizeof(Slot) * (size_t)slot_off[n]) || !be.fill(d_used, 0, 2 * sizeof(uint32_t)) ||
	    !be.upload(d_slot_off, slot_off.data(), n + 1) || !be.upload(d_list_off, list_off.data(), n + 1) || !be.upload(d_units, units.data() + lo, n)) return false;
	SeedWarpFn f;
	f.q = Q.q; f.sub = I.reads(); f.units = d_units; f.begin = I.begin; f.pos = I.pos; f.gate = gate; f.maxc = maxc;
	f.tab.slot_off = d_slot_off; f.tab.list_off = d_list_off; f.tab.slots = slots; f.tab.lists = lists; f.tab.pool = pool; f.tab.pool_used = d_used;
	f.tab.pool_cap = (uint32_t)pool_cap;
	f.tab.big = big; f.tab.big_used = d_used + 1; f.tab.big_cap = (uint32_t)big_cap;
	f.cands = S.cands; f.ncand = S.ncand; f.status = d_status;
	if (!be.launch_seed((int64_t)n, f, ST_SEED)) return false;
	std::vector<int32_t> status(n);
	if (!be.download(status.data(), d_status, n)) return false;
	++*batches;
	if (!be.release(slots) || !be.release(lists) || !be.release(pool) || !be.release(big) || !be.release(d_slot_off) || !be.release(d_list_off) || !be.release(d_units) ||
	    !be.release(d_status) || !be.release(d_used)) return false;
	bool full = false;
	for (size_t i = 0; i < n; ++i) {
		if (status[i] == 2) { be.fail("asm: a read of 100 000 letters or more (the reference's buffers hold RM - 1, mecat2asmpw.c:17)"); return false; }
		if (status[i] == 1) full = true;
	}
	if (!full) return true;
	if (n == 1) { be.fail("asm: block table of a single strand outgrew its bound"); return false; }
	const size_t mid = lo + n / 2;
	return seed_range(be, I, Q, S, units, lo, mid, gate, maxc, 2 * mult, batches) && seed_range(be, I, Q, S, units, mid, hi, gate, maxc, 2 * mult, batches);
}

struct Counters { int64_t seed_batches = 0, candidates = 0, hits = 0, extend_passes = 0; };

// pairwise_mapping :515-984 for the reads of one query file against the index
template <class B>
bool overlaps(B& be, const AsmIndex& I, const char* h_qtext, int64_t qn, const int32_t* h_qstart, const int32_t* h_qlen, int32_t nq, int32_t qfirst,
              int variant, int maxc, std::vector<Overlap>& out, Counters* cnt)
{
	out.clear();
	if (nq <= 0) return true;
	if (maxc < 1 || maxc > MAXC_MAX || (variant != 0 && variant != 1)) { be.fail("asm: variant must be 0 or 1 and max_candidates 1 .. 100"); return false; }
	char* d_text = be.template alloc<char>((size_t)qn + 16);
	int32_t* d_start = be.template alloc<int32_t>((size_t)nq);
	int32_t* d_len = be.template alloc<int32_t>((size_t)nq);
	const int64_t nstrand = 2 * (int64_t)nq;
	int32_t* d_hits = be.template alloc<int32_t>((size_t)nstrand);
	if (!d_text || !d_start || !d_len || !d_hits) return false;
	if (!be.fill(d_text, 0, (size_t)qn + 16) || !be.upload(d_text, h_qtext, (size_t)qn) || !be.upload(d_start, h_qstart, (size_t)nq) || !be.upload(d_len, h_qlen, (size_t)nq))
		return false;
	UpperFn fu; fu.text = d_text; fu.n = qn;
	if (!be.launch((qn + 15) / 16, fu, ST_SEED)) return false;
	QuerySet Q;
	Q.q.text = d_text; Q.q.start = d_start; Q.q.len = d_len; Q.q.n = nq; Q.q.first_id = qfirst; Q.h_len = h_qlen; Q.max_len = 0;
	for (int32_t r = 0; r < nq; ++r) Q.max_len = std::max(Q.max_len, h_qlen[r]);
	HitCountFn fh; fh.q = Q.q; fh.begin = I.begin; fh.hits = d_hits;
	if (!be.launch(nstrand, fh, ST_SEED)) return false;
	std::vector<int32_t> hits((size_t)nstrand);
	if (!be.download(hits.data(), d_hits, (size_t)nstrand) || !be.release(d_hits)) return false;

	SeedState S;
	S.cands = be.template alloc<Cand>((size_t)nstrand * maxc);
	S.ncand = be.template alloc<int32_t>((size_t)nstrand);
	if (!S.cands || !S.ncand) return false;
	const int64_t nblocks = I.n / ZV + 2;
	S.cap.resize((size_t)nstrand);
	S.hits = hits;
	for (int64_t u = 0; u < nstrand; ++u) { S.cap[(size_t)u] = (int32_t)std::min<int64_t>(hits[(size_t)u], nblocks); if (cnt) cnt->hits += hits[(size_t)u]; }
	// batches of strands whose tables fit the budget
	std::vector<int32_t> units((size_t)nstrand);
	for (int64_t u = 0; u < nstrand; ++u) units[(size_t)u] = (int32_t)u;
	const int64_t budget = be.table_budget();
	const int gate = variant ? 8 : 10;                         // :640
	int64_t batches = 0;
	for (size_t lo = 0; lo < units.size();) {
		size_t hi = lo;
		int64_t bytes = 0;
		while (hi < units.size()) {
			const int32_t u = units[hi];
			const int64_t c = table_room(S.cap[(size_t)u], S.hits[(size_t)u], Q.h_len[u >> 1], I.n, 1);
			const int64_t need = (int64_t)slots_for(c) * (int64_t)sizeof(Slot) + c * 4 + (c / be.pool_divisor() + 1) * (int64_t)sizeof(Bucket) +
			                     entries_room(S.cap[(size_t)u], S.hits[(size_t)u], 1) * (int64_t)sizeof(Entries) + 64;
			if (hi > lo && bytes + need > budget) break;
			bytes += need; ++hi;
		}
		if (!seed_range(be, I, Q, S, units, lo, hi, gate, maxc, 1, &batches)) return false;
		lo = hi;
	}
	if (cnt) cnt->seed_batches += batches;

	Cand* merged = be.template alloc<Cand>((size_t)nq * maxc);
	int32_t* nmerged = be.template alloc<int32_t>((size_t)nq);
	if (!merged || !nmerged) return false;
	MergeFn fm; fm.cands = S.cands; fm.ncand = S.ncand; fm.maxc = maxc; fm.merged = merged; fm.nmerged = nmerged;
	if (!be.launch(nq, fm, ST_SEED)) return false;
	if (!be.release(S.cands) || !be.release(S.ncand)) return false;

	// extension: a fixed number of units in flight, each with its own scratch
	const int64_t total = (int64_t)nq * maxc;
	Overlap* d_out = be.template alloc<Overlap>((size_t)total);
	int32_t* d_valid = be.template alloc<int32_t>((size_t)total);
	int32_t* d_over = be.template alloc<int32_t>(1);
	if (!d_out || !d_valid || !d_over) return false;
	for (int pass = 0; pass < 2; ++pass) {
		// columns of the left half: both sequences advance together, so 2.5 x the query's length covers any sane
		// alignment; a side that outgrows it raises the flag and the pass is repeated with the hard bound
		const int64_t lcap = pass == 0 ? (int64_t)Q.max_len * 5 / 2 + 64 : (int64_t)Q.max_len + I.max_len + 64;
		const int64_t per_slot = ((EXT_FIXED_BYTES + 15) / 16) * 16 + 2 * ((lcap + 15) / 16) * 16;
		int64_t slots = be.extend_slots();
		while (slots > 1 && slots * per_slot > budget) slots /= 2;
		if (slots > total) slots = total;
		char* arena = be.template alloc<char>((size_t)(slots * per_slot));
		ExtScratch* d_scr = be.template alloc<ExtScratch>((size_t)slots);
		if (!arena || !d_scr) return false;
		std::vector<ExtScratch> scr((size_t)slots);
		for (int64_t s = 0; s < slots; ++s) {
			char* p = arena + s * per_slot;
			ExtScratch& E = scr[(size_t)s];
			E.V = (int32_t*)p; p += 4 * VU_INTS;
			E.U = (int32_t*)p; p += 4 * VU_INTS;
			E.dp = (uint32_t*)p; p += 4 * DP_CELLS;
			E.row_start = (int32_t*)p; p += 4 * (MAX_D + 1);
			E.row_min = (int32_t*)p; p += 4 * (MAX_D + 1);
			E.cq = p; p += CHUNK_COLS;
			E.ct = p;
			p = arena + s * per_slot + ((EXT_FIXED_BYTES + 15) / 16) * 16;
			E.l1 = p; p += ((lcap + 15) / 16) * 16;
			E.l2 = p; E.lcap = lcap;
		}
		if (!be.upload(d_scr, scr.data(), (size_t)slots) || !be.fill(d_over, 0, 4)) return false;
		ExtendFn fe;
		fe.q = Q.q; fe.sub = I.reads(); fe.cands = merged; fe.ncand = nmerged; fe.maxc = maxc; fe.variant = variant;
		fe.scratch = d_scr; fe.out = d_out; fe.valid = d_valid; fe.overflow = d_over;
		if (!be.launch_slots(total, fe, slots, ST_EXTEND)) return false;
		int32_t over = 0;
		if (!be.download(&over, d_over, 1) || !be.release(arena) || !be.release(d_scr)) return false;
		if (cnt) cnt->extend_passes++;
		if (!over) break;
		if (pass == 1) { be.fail("asm: left half of an alignment outgrew both sequences"); return false; }
	}
	// only the printed overlaps come back: counted per read on the device, placed by a host scan of the counts
	int32_t* d_kept = be.template alloc<int32_t>((size_t)nq);
	int64_t* d_first = be.template alloc<int64_t>((size_t)nq);
	if (!d_kept || !d_first) return false;
	KeptCountFn fk; fk.valid = d_valid; fk.ncand = nmerged; fk.maxc = maxc; fk.kept = d_kept;
	if (!be.launch(nq, fk, ST_EXTEND)) return false;
	std::vector<int32_t> nm((size_t)nq), kept((size_t)nq);
	if (!be.download(kept.data(), d_kept, (size_t)nq) || !be.download(nm.data(), nmerged, (size_t)nq)) return false;
	std::vector<int64_t> first((size_t)nq);
	int64_t nout = 0;
	for (int32_t r = 0; r < nq; ++r) { first[(size_t)r] = nout; nout += kept[(size_t)r]; if (cnt) cnt->candidates += nm[(size_t)r]; }
	Overlap* d_res = be.template alloc<Overlap>((size_t)nout);
	if (!d_res || !be.upload(d_first, first.data(), (size_t)nq)) return false;
	GatherFn fg; fg.all = d_out; fg.valid = d_valid; fg.ncand = nmerged; fg.maxc = maxc; fg.first = d_first; fg.out = d_res;
	if (!be.launch(nq, fg, ST_EXTEND)) return false;
	out.resize((size_t)nout);
	if (!be.download(out.data(), d_res, (size_t)nout)) return false;
	if (!be.release(d_kept) || !be.release(d_first) || !be.release(d_res)) return false;
	return be.release(merged) && be.release(nmerged) && be.release(d_out) && be.release(d_valid) && be.release(d_over) && be.release(d_text) &&
	       be.release(d_start) && be.release(d_len);
}

}  // namespace mbasm
