// mecat_b200/csrc/ref_pipeline.h -- mecat2ref (reads against a reference genome) as a sequence of kernels and batched
// extensions (reference src/mecat2ref/mecat2ref_impl_large.cpp:274-891,
// src/mecat2ref/mecat2ref_aux.cpp:86-473).
//
// The reference maps one read at a time: seed both strands, extend the <= n best candidates, ask the block tables for
// candidates beyond clipped alignment ends, extend those, print; a read without any alignment is seeded again with a
// denser stride and wider blocks.  The same work here, batch-wise:
//
//   stage            unit                 reference
//   CountFn          strand               (layout only: index hits of the strand = room for its block table)
//   SeedFn           strand               transnum_buchang, the seeding loop with insert_loc, the candidate walk with
//                                         find_location and the neighbour votes
//   extension        candidate            extend_candidate -> GapAligner::go on a window of the reference (align.cu, row R1)
//   (host)           read                 rescue_clipped_align's sort / containment filter / choice of clipped ends
//   RescueFn         clipped end          get_left / get_right_clipped_candidate + fill_clipped_candidate
//   extension        rescue candidate     extend_candidate
//   (host)           read                 the rest of rescue_clipped_align, output_results
//
// and the reads that found nothing go round again as the second pass.  The block tables of a batch of reads stay in
// device memory from SeedFn to RescueFn; batches are cut so that they fit a memory budget.  The code is written against
// the same small backend interface as cns_pipeline.h: refmap.cu is the CUDA backend (the only one in the product
// library), tests/ref_host_harness.cpp a host backend for the CPU test-suite.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/mecat_b200.h"
#include "ref_core.cuh"

namespace mbref {

enum { ST_COUNT = 0, ST_SEED = 1, ST_RESCUE = 2, ST_NUM = 3 };      // kernel slots for the backend's per-stage timers

struct Params
{
	int num_candidates = 10;        // -n
	int num_output = 10;            // -b
	bool want_strings = true;       // the ref format prints both alignment strings; the m4 format only needs their statistics
	int64_t table_budget = 8ll << 30;
	bool seed_per_thread = false;   // SeedFn (a thread per strand) instead of SeedWarpFn (a warp per strand): the earlier form, kept for cross-checks
	// With strings wanted: extend every candidate for its coordinates only and compute alignment strings just for the
	// records that are printed (most of a read's <= -n extensions are dropped as contained in repeat-rich genomes).
	bool strings_for_printed_only = false;
	// test hook: the candidate list of every strand of the first pass (2 counts per read, 4 ints per candidate:
	// loc1 loc2 score chain), as SeedFn left them
	std::vector<int32_t>* dump_counts = nullptr;
	std::vector<int32_t>* dump_rows = nullptr;
};

struct MapIn
{
	int R = 0;                               // reads of this call
	const int32_t* h_len = nullptr;          // [R]
	const int32_t* h_fread = nullptr;        // [R] volume read holding the forward strand
	const int32_t* h_rread = nullptr;        // [R] volume read the reverse strand is taken from ...
	const int32_t* h_rrc = nullptr;          // [R] ... as its reverse complement (1) or as it is (0: the strand was packed explicitly)
	int64_t seqcount = 0;                    // bases of the reference (all sequences concatenated)
	// device side
	const uint32_t* d_fwd = nullptr;         // the read volume's forward words (volume.cu)
	const int32_t* d_offsz = nullptr;        // {offset, size} per volume read
	const int64_t* d_bad = nullptr;          // ascending volume offsets of letters that are not upper-case ACGT
	int64_t nbad = 0;
	const uint32_t* d_ibegin = nullptr;      // k-mer index of the reference: CSR over the 2^26 codes
	const int32_t* d_ipos = nullptr;         // 0-based k-mer starts, ascending inside a list
};

// A malloc'ed byte blob that grows by doubling and can be handed to the caller of the C ABI as it is (the alignment
// strings of a batch are gigabytes: every copy saved counts).
struct Blob
{
	char* p = nullptr;
	size_t len = 0, cap = 0;
	bool oom = false;
	Blob() {}
	Blob(const Blob&) = delete;
	Blob& operator=(const Blob&) = delete;
	~Blob() { free(p); }
	void append(const char* src, size_t n)
	{
		if (len + n + 1 > cap) {
			size_t want = cap ? cap * 2 : (size_t)1 << 16;
			while (want < len + n + 1) want *= 2;
			char* np = (char*)realloc(p, want);
			if (!np) { oom = true; return; }
			p = np; cap = want;
		}
		memcpy(p + len, src, n);
		len += n;
		p[len] = 0;
	}
	const char* data() const { return p ? p : ""; }
	size_t size() const { return len; }
	char* release()          // the caller owns the (NUL-terminated) bytes; never NULL
	{
		char* r = p ? p : (char*)calloc(1, 1);
		p = nullptr; len = cap = 0;
		return r;
	}
};

struct Sink
{
	std::vector<mecat_ref_result> recs;
	Blob q, s;                               // NUL-terminated alignment strings of the records (want_strings)
};

// The index's view of the genome: every maximal ACGT run cut into chunks of INDEX_CHUNK k-mer starts, each chunk
// carrying the 12 bases a k-mer needs beyond its start.  A chunk is one "read" of the index build (index.cu: one CTA's
// work); no k-mer spans a letter that is not ACGT, every k-mer start of a run lies in exactly one chunk.
constexpr int64_t INDEX_CHUNK = 32768;
inline bool index_chunks(const int64_t* run_start_len, int32_t num_runs, int64_t num_bases, std::vector<int32_t>& chunks /* {offset, size} pairs */)
{
	chunks.clear();
	int64_t prev_end = 0;
	for (int32_t r = 0; r < num_runs; ++r) {
		const int64_t s = run_start_len[2 * r], n = run_start_len[2 * r + 1];
		if (s < prev_end || n < 0 || s + n > num_bases) return false;
		prev_end = s + n;
		for (int64_t k = 0; k + SEED <= n; k += INDEX_CHUNK) {
			chunks.push_back((int32_t)(s + k));
			chunks.push_back((int32_t)std::min<int64_t>(n - k, INDEX_CHUNK + SEED - 1));
		}
	}
	return true;
}

// ---------------------------------------------------------------------------------------------- functors
struct CountFn
{
	const Unit* units; const int32_t* offsz; const uint32_t* fwd; const uint32_t* ibegin; const int64_t* bad; int64_t nbad; int32_t* hits;
	REF_HD void operator()(int64_t u) const
	{
		const Unit U = units[u];
		const int64_t h = count_hits(fwd, (uint32_t)offsz[2 * U.vread], U, ibegin, bad, nbad);
		hits[u] = (int32_t)(h > 0x7fffffff ? 0x7fffffff : h);
	}
};

struct TableRefs        // where the block table of strand u lives
{
	const int64_t* rec_off; const int64_t* slot_off; Slot* slots; Bucket* recs; int32_t* nrec;
	REF_HD Table open(int64_t u, bool fresh) const
	{
		Table T;
		const uint32_t cap = (uint32_t)(slot_off[u + 1] - slot_off[u]);
		T.slots = slots + slot_off[u]; T.mask = cap - 1; T.shift = table_shift(cap);
		T.recs = recs + rec_off[u]; T.nrec = fresh ? 0 : nrec[u];
		return T;
	}
};

struct SeedFn
{
	const Unit* units; const int32_t* offsz; const uint32_t* fwd; const uint32_t* ibegin; const int32_t* ipos; const int64_t* bad; int64_t nbad;
	TableRefs tab; int zv, gate, maxc; int64_t seqcount; RefCand* cands; int32_t* ncand;
	REF_HD void operator()(int64_t u) const
	{
		const Unit U = units[u];
		Table T = tab.open(u, true);
		seed_strand(fwd, (uint32_t)offsz[2 * U.vread], U, zv, ibegin, ipos, bad, nbad, T);
		tab.nrec[u] = T.nrec;
		ncand[u] = walk_strand(U, zv, gate, seqcount, (int)(u & 1), T, cands + u * maxc, maxc);
	}
};

struct SeedWarpFn       // the same unit of work on the 32 lanes of a warp (ref_core.cuh: seed_strand_w, walk_strand_w)
{
	const Unit* units; const int32_t* offsz; const uint32_t* fwd; const uint32_t* ibegin; const int32_t* ipos; const int64_t* bad; int64_t nbad;
	TableRefs tab; int zv, gate, maxc; int64_t seqcount; RefCand* cands; int32_t* ncand;
	template <class L>
	REF_HD void operator()(int64_t u, const L& lanes, WarpScratch& W) const
	{
		const Unit U = units[u];
		Table T = tab.open(u, true);
		seed_strand_w(lanes, fwd, (uint32_t)offsz[2 * U.vread], U, zv, ibegin, ipos, bad, nbad, T, W);
		if (lanes.leader()) tab.nrec[u] = T.nrec;
		const int n = walk_strand_w(lanes, U, zv, gate, seqcount, (int)(u & 1), T, cands + u * maxc, maxc, W);
		if (lanes.leader()) ncand[u] = n;
	}
};

struct RescueFn
{
	const RescueQuery* queries; const Unit* units; TableRefs tab; int zv; int64_t ref_size; RefCand* out; int32_t* ok;
	REF_HD void operator()(int64_t i) const
	{
		const RescueQuery q = queries[i];
		const Table T = tab.open(q.unit, false);
		RefCand c; c.loc1 = c.loc2 = c.score = c.chain = 0;
		ok[i] = rescue_candidate(q, T, zv, units[q.unit].bc, ref_size, &c) ? 1 : 0;
		out[i] = c;
	}
};

// ---------------------------------------------------------------------------------------------- host side
struct AlignInfo       // mecat2ref_aux.h:27-43
{
	int qoff, qend, parent_id, id, prev_id, next_id;
	char valid, qdir;
	int64_t soff, send;
	bool operator<(const AlignInfo& r) const { return (qend - qoff) > (r.qend - r.qoff); }
};

struct Hit             // TempResult without its strings
{
	int dir, vscore, qb, qe, columns, matches;
	int64_t sb, se;
	int64_t str_at;    // offset of the strings in the extension call's string buffers
	int call;          // 0 = primary extensions, 1 = rescue extensions
	mecat_align_task task;     // what was extended (strings_for_printed_only asks for it again, with strings)
	int64_t win;
};

struct ReadState
{
	std::vector<Hit> hits;
	std::vector<AlignInfo> alns;
	int naln = 0;
	int first_task = 0, ntask = 0;            // primary extension tasks of the read
	struct Pick { int aln, side; };           // clipped ends that were asked about, in the reference's order
	std::vector<Pick> picks;
	int first_query = 0;
};

inline bool aln_full(const AlignInfo& a, int qsize) { return a.qend - a.qoff >= qsize * 0.9; }
inline bool aln_contained(const AlignInfo& a, const AlignInfo& b)
{
	const int extra = 100;
	return a.qdir == b.qdir && b.qoff + extra >= a.qoff && b.qend <= a.qend + extra && b.soff + extra >= a.soff && b.send <= a.send + extra;
}
inline bool aln_left_of(const AlignInfo& a, const AlignInfo& b)       // is_left_clipped_align, mecat2ref_aux.cpp:256-277
{
	if (a.qdir != b.qdir) return false;
	if (abs(b.qend - a.qoff) <= 200 && a.soff - b.send > -200 && a.soff - b.send < 10000) return true;
	if (llabs((long long)(b.send - a.soff)) <= 200 && a.qoff - b.qend > -200 && a.qoff - b.qend < 10000) return true;
	return false;
}
inline bool aln_right_of(const AlignInfo& a, const AlignInfo& b)      // is_right_clipped_align, :279-303
{
	if (a.qdir != b.qdir) return false;
	if (abs(a.qend - b.qoff) <= 200 && b.soff - a.send > -200 && b.soff - a.send < 10000) return true;
	if (llabs((long long)(a.send - b.soff)) <= 200 && b.qoff - a.qend > -200 && b.qoff - a.qend < 10000) return true;
	return false;
}

// extract_sequences, mecat2ref_aux.cpp:86-121: the window of the reference a candidate is extended in
inline bool make_task(const RefCand& c, const Unit& u, int64_t ref_size, mecat_align_task* t, int64_t* win_start)
{
	const int64_t read_start = c.loc2, ref_start = (int64_t)c.loc1 - 1;
	const int64_t L1 = read_start, R1 = u.len - read_start, L2 = ref_start, R2 = ref_size - ref_start;
	const int64_t L = std::min(L1, L2), R = std::min(R1, R2);
	const int64_t left_ref = std::min(L2, (int64_t)(L * 1.2)), right_ref = std::min(R2, (int64_t)(R * 1.2));
	if (left_ref + right_ref <= 0 || read_start < 0 || read_start > u.len) return false;
	t->qread = u.vread; t->qstrand = u.rc; t->qstart = (int32_t)read_start;
	t->sread = 0; t->sstart = (int32_t)left_ref;
	t->swin_off = (int32_t)(ref_start - left_ref); t->swin_len = (int32_t)(left_ref + right_ref);
	*win_start = ref_start - left_ref;
	return true;
}

inline AlignInfo make_aln(const Hit& h, int id)
{
	AlignInfo a;
	a.qoff = h.qb; a.qend = h.qe; a.qdir = (char)(h.dir ? 'R' : 'F'); a.soff = h.sb; a.send = h.se; a.valid = 1;
	a.id = id; a.prev_id = -1; a.next_id = -1; a.parent_id = -1;
	return a;
}

// First half of rescue_clipped_align (mecat2ref_aux.cpp:305-395): order by aligned length, drop contained alignments,
// and for the three best that do not cover the read name the ends to ask about.
inline void rescue_prepare(ReadState& S, int read_len)
{
	std::vector<AlignInfo>& al = S.alns;
	int naln = (int)al.size();
	std::sort(al.begin(), al.end());
	for (int i = 0; i < naln - 1; ++i) {
		if (!al[i].valid) continue;
		for (int j = i + 1; j < naln; ++j) if (al[j].valid && aln_contained(al[i], al[j])) al[j].valid = 0;
	}
	int k = 0;
	for (int i = 0; i < naln; ++i) if (al[i].valid) al[k++] = al[i];
	naln = k;
	al.resize((size_t)naln);
	S.naln = naln;
	S.picks.clear();
	if (naln == 0 || aln_full(al[0], read_len)) return;
	// (the reference's chaining loop over pairs of existing alignments can never fire: it requires prev_id / next_id
	// to be set already, :347-366)
	const int n = std::min(naln, 3);
	for (int i = 0; i < n; ++i) {
		ReadState::Pick p;
		p.aln = i; p.side = 0; S.picks.push_back(p);
		p.side = 1; S.picks.push_back(p);
	}
}

// Second half (:396-452): a rescued alignment that continues its parent is chained to it; when anything was chained
// the list is re-sorted and, if some alignment now covers the read, only those are kept.
inline void rescue_finish(ReadState& S, int read_len, const std::vector<int>& pick_hit /* hit id or -1 per pick */)
{
	std::vector<AlignInfo>& al = S.alns;
	int naln = S.naln, k = 0;
	for (size_t p = 0; p < S.picks.size(); ++p) {
		if (pick_hit[p] < 0) continue;
		const AlignInfo par = al[(size_t)S.picks[p].aln];
		AlignInfo ai = make_aln(S.hits[(size_t)pick_hit[p]], pick_hit[p]);
		const bool chained = S.picks[p].side == 0 ? aln_left_of(par, ai) : aln_right_of(par, ai);
		if (!chained) continue;
		ai.parent_id = par.id;
		if (S.picks[p].side == 0) al[(size_t)S.picks[p].aln].prev_id = ai.id; else al[(size_t)S.picks[p].aln].next_id = ai.id;
		al.push_back(ai);
		++k;
	}
	if (!k) return;
	naln += k;
	std::sort(al.begin(), al.end());
	k = 0;
	for (int i = 0; i < naln; ++i)
		if (aln_full(al[(size_t)i], read_len)) { al[(size_t)i].parent_id = -1; al[(size_t)i].prev_id = -1; al[(size_t)i].next_id = -1; ++k; }
	if (k) al.resize((size_t)k);
}

// One pass over `reads` (indices into in.h_*): returns the reads that found no alignment.
template <class Backend>
int map_pass(Backend& be, const MapIn& in, const Params& P, int pass, const std::vector<int>& reads, std::vector<int>& unmapped, Sink& out)
{
	const int zv = pass == 0 ? ZV : ZVS, gate = pass == 0 ? 6 : 4, maxc = P.num_candidates;
	const int64_t NU = 2 * (int64_t)reads.size();
	if (!NU) return 0;
	std::vector<Unit> units((size_t)NU);
	for (size_t i = 0; i < reads.size(); ++i) {
		const int r = reads[i], len = in.h_len[r];
		int bc = pass == 0 ? 5 + len / 1000 : 5;
		if (bc > 20) bc = 20;
		Unit f; f.vread = in.h_fread[r]; f.rc = 0; f.len = len; f.bc = bc;
		Unit v; v.vread = in.h_rread[r]; v.rc = in.h_rrc[r]; v.len = len; v.bc = bc;
		units[2 * i] = f; units[2 * i + 1] = v;
	}
	Unit* d_units = be.template alloc<Unit>((size_t)NU);
	int32_t* d_hits = be.template alloc<int32_t>((size_t)NU);
	if (!d_units || !d_hits) return 1;
	if (!be.upload(d_units, units.data(), (size_t)NU)) return 1;
	{
		CountFn f{d_units, in.d_offsz, in.d_fwd, in.d_ibegin, in.d_bad, in.nbad, d_hits};
		if (!be.launch(NU, f, ST_COUNT)) return 1;
	}
	std::vector<int32_t> hits((size_t)NU);
	if (!be.download(hits.data(), d_hits, (size_t)NU)) return 1;
	{
		int64_t all = 0;
		for (int32_t h : hits) all += h;
		be.note_hits(all);          // statistics: the unit of SeedFn's algorithmic bytes
	}

	auto cap_of = [](int64_t h) { int64_t c = 4; while (c < 2 * h + 2) c <<= 1; return c; };
	size_t r0 = 0;
	while (r0 < reads.size()) {
		// ---- a batch of reads whose block tables fit the budget
		size_t r1 = r0;
		int64_t bytes = 0;
		while (r1 < reads.size()) {
			int64_t need = 0;
			for (int s = 0; s < 2; ++s) { const int64_t h = hits[2 * r1 + s]; need += h * (int64_t)sizeof(Bucket) + cap_of(h) * (int64_t)sizeof(Slot); }
			if (r1 > r0 && bytes + need > P.table_budget) break;
			bytes += need; ++r1;
		}
		const int64_t u0 = 2 * (int64_t)r0, nu = 2 * (int64_t)(r1 - r0);
		std::vector<int64_t> rec_off((size_t)nu + 1), slot_off((size_t)nu + 1);
		int64_t nrecs = 0, nslots = 0;
		for (int64_t u = 0; u < nu; ++u) { rec_off[(size_t)u] = nrecs; slot_off[(size_t)u] = nslots; nrecs += hits[(size_t)(u0 + u)]; nslots += cap_of(hits[(size_t)(u0 + u)]); }
		rec_off[(size_t)nu] = nrecs; slot_off[(size_t)nu] = nslots;
		int64_t* d_rec_off = be.template alloc<int64_t>((size_t)nu + 1);
		int64_t* d_slot_off = be.template alloc<int64_t>((size_t)nu + 1);
		Slot* d_slots = be.template alloc<Slot>((size_t)nslots);
		Bucket* d_recs = be.template alloc<Bucket>((size_t)nrecs);
		int32_t* d_nrec = be.template alloc<int32_t>((size_t)nu);
		RefCand* d_cands = be.template alloc<RefCand>((size_t)(nu * maxc));
		int32_t* d_ncand = be.template alloc<int32_t>((size_t)nu);
		if (!d_rec_off || !d_slot_off || !d_slots || !d_recs || !d_nrec || !d_cands || !d_ncand) return 1;
		if (!be.upload(d_rec_off, rec_off.data(), (size_t)nu + 1) || !be.upload(d_slot_off, slot_off.data(), (size_t)nu + 1)) return 1;
		if (!be.fill(d_slots, 0, (size_t)nslots * sizeof(Slot))) return 1;
		const TableRefs tab{d_rec_off, d_slot_off, d_slots, d_recs, d_nrec};
		if (P.seed_per_thread) {
			SeedFn f{d_units + u0, in.d_offsz, in.d_fwd, in.d_ibegin, in.d_ipos, in.d_bad, in.nbad, tab, zv, gate, maxc, in.seqcount, d_cands, d_ncand};
			if (!be.launch(nu, f, ST_SEED)) return 1;
		} else {
			SeedWarpFn f{d_units + u0, in.d_offsz, in.d_fwd, in.d_ibegin, in.d_ipos, in.d_bad, in.nbad, tab, zv, gate, maxc, in.seqcount, d_cands, d_ncand};
			if (!be.launch_seed_warp(nu, f, ST_SEED)) return 1;
		}
		std::vector<int32_t> ncand((size_t)nu);
		std::vector<RefCand> cands((size_t)(nu * maxc));
		if (!be.download(ncand.data(), d_ncand, (size_t)nu) || !be.download(cands.data(), d_cands, (size_t)(nu * maxc))) return 1;

		if (pass == 0 && P.dump_counts && P.dump_rows)
			for (int64_t u = 0; u < nu; ++u) {
				P.dump_counts->push_back(ncand[(size_t)u]);
				for (int k = 0; k < ncand[(size_t)u]; ++k) {
					const RefCand& c = cands[(size_t)(u * maxc + k)];
					P.dump_rows->push_back(c.loc1); P.dump_rows->push_back(c.loc2); P.dump_rows->push_back(c.score); P.dump_rows->push_back(c.chain);
				}
			}

		// ---- the read's list: forward-strand candidates were inserted first, so they stay ahead of equal scores
		const size_t nr = r1 - r0;
		std::vector<ReadState> st(nr);
		std::vector<mecat_align_task> tasks;
		struct TaskRef { int read, unit, vscore; int64_t win; };
		std::vector<TaskRef> tref;
		for (size_t i = 0; i < nr; ++i) {
			const RefCand* cf = &cands[(size_t)(2 * i) * maxc];
			const RefCand* cr = &cands[(size_t)(2 * i + 1) * maxc];
			const int nf = ncand[2 * i], nv = ncand[2 * i + 1];
			st[i].first_task = (int)tasks.size();
			int a = 0, b = 0;
			for (int k = 0; k < maxc && (a < nf || b < nv); ++k) {
				const bool takef = a < nf && (b >= nv || cf[a].score >= cr[b].score);
				const RefCand& c = takef ? cf[a++] : cr[b++];
				const int unit = (int)(2 * i) + (takef ? 0 : 1);
				mecat_align_task t;
				int64_t win;
				if (!make_task(c, units[(size_t)(u0 + unit)], in.seqcount, &t, &win)) continue;
				tasks.push_back(t);
				TaskRef tr; tr.read = (int)i; tr.unit = unit; tr.vscore = c.score; tr.win = win;
				tref.push_back(tr);
			}
			st[i].ntask = (int)tasks.size() - st[i].first_task;
		}
		std::vector<mecat_align_result> res(tasks.size());
		std::vector<char> q0, s0, q1, s1;
		const bool strings_now = P.want_strings && !P.strings_for_printed_only;
		if (!be.align(tasks.data(), tasks.size(), strings_now, res.data(), q0, s0)) return 1;
		for (size_t t = 0; t < tasks.size(); ++t) {
			if (!res[t].ok) continue;
			ReadState& S = st[(size_t)tref[t].read];
			Hit h;
			h.dir = tref[t].unit & 1; h.vscore = tref[t].vscore; h.qb = res[t].qstart; h.qe = res[t].qend;
			h.sb = tref[t].win + res[t].sstart; h.se = tref[t].win + res[t].send;
			h.columns = res[t].columns; h.matches = res[t].matches; h.str_at = res[t].str_offset; h.call = 0;
			h.task = tasks[t]; h.win = tref[t].win;
			S.hits.push_back(h);
			S.alns.push_back(make_aln(h, (int)S.hits.size() - 1));
		}

		// ---- clipped ends
		std::vector<RescueQuery> queries;
		for (size_t i = 0; i < nr; ++i) {
			ReadState& S = st[i];
			if (S.alns.empty()) continue;
			rescue_prepare(S, units[(size_t)(u0 + 2 * i)].len);
			S.first_query = (int)queries.size();
			for (const ReadState::Pick& p : S.picks) {
				const AlignInfo& a = S.alns[(size_t)p.aln];
				RescueQuery q;
				q.unit = (int32_t)(2 * i) + (a.qdir == 'F' ? 0 : 1); q.side = p.side; q.qoff = a.qoff; q.qend = a.qend;
				q.read_len = units[(size_t)(u0 + 2 * i)].len; q.pad_ = 0; q.soff = a.soff; q.send = a.send;
				queries.push_back(q);
			}
		}
		std::vector<int32_t> qok(queries.size());
		std::vector<RefCand> qcand(queries.size());
		if (!queries.empty()) {
			RescueQuery* d_q = be.template alloc<RescueQuery>(queries.size());
			RefCand* d_qc = be.template alloc<RefCand>(queries.size());
			int32_t* d_qok = be.template alloc<int32_t>(queries.size());
			if (!d_q || !d_qc || !d_qok || !be.upload(d_q, queries.data(), queries.size())) return 1;
			RescueFn f{d_q, d_units + u0, tab, zv, in.seqcount, d_qc, d_qok};
			if (!be.launch((int64_t)queries.size(), f, ST_RESCUE)) return 1;
			if (!be.download(qok.data(), d_qok, queries.size()) || !be.download(qcand.data(), d_qc, queries.size())) return 1;
			if (!be.release(d_q) || !be.release(d_qc) || !be.release(d_qok)) return 1;
		}
		if (!be.release(d_slots) || !be.release(d_recs) || !be.release(d_rec_off) || !be.release(d_slot_off) || !be.release(d_nrec) ||
		    !be.release(d_cands) || !be.release(d_ncand)) return 1;
		std::vector<mecat_align_task> rtasks;
		std::vector<int> rquery;
		std::vector<int64_t> rwin;
		for (size_t k = 0; k < queries.size(); ++k) {
			if (!qok[k]) continue;
			mecat_align_task t;
			int64_t win;
			if (!make_task(qcand[k], units[(size_t)(u0 + queries[k].unit)], in.seqcount, &t, &win)) continue;
			rtasks.push_back(t); rquery.push_back((int)k); rwin.push_back(win);
		}
		std::vector<mecat_align_result> rres(rtasks.size());
		if (!be.align(rtasks.data(), rtasks.size(), strings_now, rres.data(), q1, s1)) return 1;
		std::vector<int> query_task(queries.size(), -1);
		for (size_t t = 0; t < rtasks.size(); ++t) query_task[(size_t)rquery[t]] = (int)t;

		// ---- per read: chain the rescued alignments, emit (output_results, mecat2ref_aux.cpp:454-473)
		const size_t first_rec = out.recs.size();
		std::vector<mecat_align_task> ptasks;        // strings_for_printed_only: the printed records' extensions
		for (size_t i = 0; i < nr; ++i) {
			ReadState& S = st[i];
			const int r = reads[r0 + i];
			if (S.hits.empty()) { if (pass == 0) unmapped.push_back(r); continue; }
			std::vector<int> pick_hit(S.picks.size(), -1);
			for (size_t p = 0; p < S.picks.size(); ++p) {
				const int t = query_task[(size_t)S.first_query + p];
				if (t < 0 || !rres[(size_t)t].ok) continue;
				const RescueQuery& q = queries[(size_t)S.first_query + p];
				Hit h;
				h.dir = q.unit & 1; h.vscore = qcand[(size_t)S.first_query + p].score; h.qb = rres[(size_t)t].qstart; h.qe = rres[(size_t)t].qend;
				h.sb = rwin[(size_t)t] + rres[(size_t)t].sstart; h.se = rwin[(size_t)t] + rres[(size_t)t].send;
				h.columns = rres[(size_t)t].columns; h.matches = rres[(size_t)t].matches; h.str_at = rres[(size_t)t].str_offset; h.call = 1;
				h.task = rtasks[(size_t)t]; h.win = rwin[(size_t)t];
				S.hits.push_back(h);
				pick_hit[p] = (int)S.hits.size() - 1;
			}
			rescue_finish(S, in.h_len[r], pick_hit);
			int groups = 0, printed = 0;
			auto emit = [&](int id) {
				if (printed >= P.num_output) return;        // output_query_results prints at most -b records of a read
				const Hit& h = S.hits[(size_t)id];
				mecat_ref_result o;
				memset(&o, 0, sizeof o);
				o.read = r; o.dir = h.dir; o.vscore = h.vscore; o.qb = h.qb; o.qe = h.qe; o.qs = in.h_len[r];
				o.sb = h.sb; o.se = h.se; o.columns = h.columns; o.matches = h.matches; o.str_offset = -1;
				if (P.want_strings && !strings_now) ptasks.push_back(h.task);
				if (strings_now) {
					const std::vector<char>& qs = h.call ? q1 : q0;
					const std::vector<char>& ss = h.call ? s1 : s0;
					o.str_offset = (int64_t)out.q.size();
					out.q.append(qs.data() + h.str_at, (size_t)h.columns + 1);
					out.s.append(ss.data() + h.str_at, (size_t)h.columns + 1);
				}
				out.recs.push_back(o);
				++printed;
			};
			for (size_t a = 0; a < S.alns.size() && groups < P.num_output; ++a) {
				const AlignInfo& ai = S.alns[a];
				if (ai.parent_id != -1) continue;
				emit(ai.id);
				if (ai.prev_id != -1) emit(ai.prev_id);
				if (ai.next_id != -1) emit(ai.next_id);
				++groups;
			}
		}
		if (!ptasks.empty()) {
			// the same extensions again, now with strings; both routes must agree on every coordinate
			std::vector<mecat_align_result> pres(ptasks.size());
			std::vector<char> q2, s2;
			if (!be.align(ptasks.data(), ptasks.size(), true, pres.data(), q2, s2)) return 1;
			for (size_t k = 0; k < ptasks.size(); ++k) {
				mecat_ref_result& o = out.recs[first_rec + k];
				const mecat_align_result& a = pres[k];
				if (!a.ok || a.columns != o.columns || a.matches != o.matches || a.qstart != o.qb || a.qend != o.qe) {
					be.fail("mecat2ref: the extension with strings disagrees with the forward-only extension of the same candidate");
					return 1;
				}
				o.str_offset = (int64_t)out.q.size();
				out.q.append(q2.data() + a.str_offset, (size_t)a.columns + 1);
				out.s.append(s2.data() + a.str_offset, (size_t)a.columns + 1);
			}
		}
		r0 = r1;
	}
	if (!be.release(d_units) || !be.release(d_hits)) return 1;
	return 0;
}

// mecat2ref on R reads: records of a read are together and in the reference's order, reads in input order -- also the
// reads that needed the second seeding pass, whose records are produced after everyone's first pass (the reference runs
// a read's second pass right after its first; with -t 1 its output is in input order).
template <class Backend>
int map_reads(Backend& be, const MapIn& in, const Params& P, Sink& out)
{
	for (int r = 0; r < in.R; ++r)
		if (in.h_len[r] >= MAX_READ) { be.fail("mecat2ref: a read of 100 000 bases or more overflows the reference's read buffers (RM, mecat2ref_defs.h:16); not on this path"); return 1; }
	if (P.num_candidates < 1 || P.num_output < 1) { be.fail("mecat2ref: -n and -b must be > 0"); return 1; }
	std::vector<int> all((size_t)in.R), second, none;
	for (int r = 0; r < in.R; ++r) all[(size_t)r] = r;
	int rc = map_pass(be, in, P, 0, all, second, out);
	if (!rc) rc = map_pass(be, in, P, 1, second, none, out);
	be.end_batch();
	if (!rc && !second.empty())       // records keep their string offsets; a read's records stay adjacent and in order
		std::stable_sort(out.recs.begin(), out.recs.end(), [](const mecat_ref_result& a, const mecat_ref_result& b) { return a.read < b.read; });
	if (!rc && (out.q.oom || out.s.oom)) { be.fail("mecat2ref: out of host memory for the alignment strings"); rc = 1; }
	return rc;
}

}  // namespace mbref
