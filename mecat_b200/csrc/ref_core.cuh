// mecat_b200/csrc/ref_core.cuh -- per-unit bodies of the mecat2ref seeding / scoring kernels.
//
// mecat2ref maps every read against a k-mer index of the reference genome: both strands are seeded into blocks of
// 1 000 reference bases holding up to 20 (offset, seed ordinal) pairs, blocks whose running score passes a gate are
// scored by pairwise DDF consistency, neighbouring blocks vote, and the best candidates go to the gapped aligner;
// alignments that end inside the read ask the block table for a second candidate beyond their clipped end.
// Reference: reference_mapping, insert_loc, transnum_buchang   src/mecat2ref/mecat2ref_impl_large.cpp:64-130,274-891
//            find_location, fill_clipped_candidate, get_{left,right}_clipped_candidate
//                                                               src/mecat2ref/mecat2ref_aux.cpp:6-84,186-303
//
// The reference keeps a dense array of blocks per thread (genome / 1 000 entries of 92 bytes) that a read touches in
// ~1 000 places.  Here every (read, strand) owns a small open-addressing table keyed by block number whose records are
// created in first-touch order -- exactly the order the reference's index_list walks them in.  All bodies are integer
// code shared by the CUDA backend (refmap.cu) and the host harness of the CPU test-suite (tests/ref_host_harness.cpp).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define REF_HD __host__ __device__ __forceinline__
#else
#define REF_HD inline
#endif

namespace mbref {

constexpr int SEED = 13;          // seed_len of meap_ref_impl_large, mecat2ref_impl_large.cpp:952
constexpr int SM = 20;            // seeds kept per block, mecat2ref_defs.h:22
constexpr int ZV = 1000;          // block width of the first pass, :17
constexpr int ZVS = 2000;         // block width of the second, more sensitive pass, :18
constexpr int CLIPPED = 2000;     // an alignment is clipped when more than this is left on both sequences, mecat2ref_aux.cpp:215
constexpr int MAX_READ = 100000;  // RM: the reference's fixed read buffers, mecat2ref_defs.h:16

struct Unit            // one strand of one read
{
	int32_t vread;     // read of the uploaded volume holding the bases
	int32_t rc;        // 1: the strand is the reverse complement of that read
	int32_t len;
	int32_t bc;        // stride between sampled k-mers (BC): 5 + len / 1000 capped at 20, or 5 in the second pass
};

struct Bucket          // Back_List of one touched block (mecat2ref_aux.h:9-14)
{
	int32_t blk;
	int16_t score, score2, seednum, index_score;
	int16_t loczhi[SM], seedno[SM];
};

struct Slot { int32_t key, rec; };   // key = block + 1, 0 = empty

struct Table
{
	Slot* slots; uint32_t mask; int shift;
	Bucket* recs; int32_t nrec;
};

struct RefCand { int32_t loc1, loc2, score, chain; };      // the fields of `candidate_save` the extension reads

REF_HD int table_shift(uint32_t cap) { int s = 32; while (cap > 1) { cap >>= 1; --s; } return s; }

REF_HD Bucket* table_find(const Table& T, int32_t blk)
{
	if (blk < 0) return nullptr;
	uint32_t h = ((uint32_t)(blk + 1) * 2654435761u) >> T.shift;
	for (;;) {
		const Slot s = T.slots[h];
		if (s.key == blk + 1) return T.recs + s.rec;
		if (s.key == 0) return nullptr;
		h = (h + 1) & T.mask;
	}
}

REF_HD Bucket* table_touch(Table& T, int32_t blk, bool* fresh)
{
	uint32_t h = ((uint32_t)(blk + 1) * 2654435761u) >> T.shift;
	for (;;) {
		const Slot s = T.slots[h];
		if (s.key == blk + 1) { *fresh = false; return T.recs + s.rec; }
		if (s.key == 0) break;
		h = (h + 1) & T.mask;
	}
	Slot s; s.key = blk + 1; s.rec = T.nrec;
	T.slots[h] = s;
	Bucket* b = T.recs + T.nrec++;
	b->blk = blk; b->score = 0; b->score2 = 0; b->seednum = 0; b->index_score = 0;
	*fresh = true;
	return b;
}

// |dloc / (dseed * BC) - 1| < 0.25 in exact integers.  The reference evaluates it in float32 (insert_loc, find_location)
// and float64 (the neighbour votes); |dseed * BC| < 2^18 here, so a quotient that is not exactly 0.75 / 1.25 lies at
// least 2^-20 from them -- many ulps -- and rounding is monotone, so both float forms agree with 3 b BC < 4 a < 5 b BC
// (mirrored for b < 0; b == 0 divides by zero: inf or NaN, never close).  tests/test_ref_host.py checks the float forms.
REF_HD bool ddf_close(int64_t a, int64_t b, int bc)
{
	const int64_t a4 = 4 * a, d = b * bc;
	if (d > 0) return 3 * d < a4 && a4 < 5 * d;
	if (d < 0) return 5 * d < a4 && a4 < 3 * d;
	return false;
}

// The volume's forward words hold base p at bits 2 (p mod 16) of word p / 16 (volume.cu), A0 C1 G2 T3.
// the sixteen 2-bit groups of x in reverse order
REF_HD uint32_t reverse_groups(uint32_t x)
{
	x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
	x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
	x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
	return (x >> 16) | (x << 16);
}

// 13-mer starting at base i of the strand, first base most significant -- the code the index is addressed by.  Two word
// loads: the 13 bases are 26 consecutive bits of the volume (first base in the low bits); the forward strand reverses
// their order, the reverse strand reads the same bits downwards, so they only need complementing.
REF_HD uint32_t strand_kmer(const uint32_t* fwd, uint32_t off, const Unit& u, int i)
{
	const uint32_t p = u.rc ? off + (uint32_t)(u.len - SEED - i) : off + (uint32_t)i;
	const uint32_t w = p >> 4, sh = (p & 15u) << 1;
	const uint32_t x = (uint32_t)((((uint64_t)fwd[w + 1] << 32) | fwd[w]) >> sh);
	const uint32_t mask = (1u << (2 * SEED)) - 1;
	return u.rc ? ~x & mask : reverse_groups(x) >> (32 - 2 * SEED);
}

// does [lo, hi) hold a letter that is not upper-case ACGT?  (`bad`: ascending volume offsets of such letters;
// transnum_buchang gives the k-mer the value -1, mecat2ref_impl_large.cpp:64-90)
REF_HD bool covers_bad(const int64_t* bad, int64_t nbad, int64_t lo, int64_t hi)
{
	int64_t a = 0, b = nbad;
	while (a < b) { const int64_t m = (a + b) >> 1; if (bad[m] < lo) a = m + 1; else b = m; }
	return a < nbad && bad[a] < hi;
}

REF_HD int sampled_kmers(int len, int bc) { return len < SEED ? 0 : (len - SEED) / bc + 1; }

// insert_loc, mecat2ref_impl_large.cpp:92-130: the 21st seed of a block evicts the seed that agrees with the fewest others
REF_HD void insert_loc(Bucket* b, int loc, int seedn, int bc)
{
	int lloc[SM + 1], lseed[SM + 1], lscore[SM + 1];
	for (int i = 0; i < SM; ++i) { lloc[i] = b->loczhi[i]; lseed[i] = b->seedno[i]; lscore[i] = 0; }
	lloc[SM] = loc; lseed[SM] = seedn; lscore[SM] = 0;
	for (int i = 0; i < SM; ++i)
		for (int j = i + 1; j <= SM; ++j)
			if (lseed[j] - lseed[i] > 0 && lloc[j] - lloc[i] > 0 && ddf_close(lloc[j] - lloc[i], lseed[j] - lseed[i], bc)) { ++lscore[i]; ++lscore[j]; }
	int mini = -1, minval = 10000;
	for (int i = 0; i <= SM; ++i) if (minval > lscore[i]) { minval = lscore[i]; mini = i; }
	if (minval == SM) { b->loczhi[SM - 1] = (int16_t)loc; b->seedno[SM - 1] = (int16_t)seedn; }
	else if (minval < SM && mini < SM) {
		for (int i = mini; i < SM; ++i) { b->loczhi[i] = (int16_t)lloc[i + 1]; b->seedno[i] = (int16_t)lseed[i + 1]; }
		--b->score;
	}
}

// find_location, mecat2ref_aux.cpp:6-84.  loc = {offset, seed ordinal} of the anchor (and of a second seed the caller
// ignores); rep_loc = which entry the anchor is.
REF_HD int find_location(const int* t_loc, const int* t_seedn, int* t_score, int* loc, int k, int* rep_loc, int bc, int read_len)
{
	int maxval = 0, maxi = 0, rep = 0, lasti = 0;
	for (int i = 0; i < k; ++i) t_score[i] = 0;
	for (int i = 0; i < k - 1; ++i)
		for (int j = i + 1; j < k; ++j) {
			const int dl = t_loc[j] - t_loc[i], ds = t_seedn[j] - t_seedn[i];
			if (ds > 0 && dl > 0 && dl < read_len && ddf_close(dl, ds, bc)) { ++t_score[i]; ++t_score[j]; }
		}
	for (int i = 0; i < k; ++i) {
		if (maxval < t_score[i]) { maxval = t_score[i]; maxi = i; rep = 0; }
		else if (maxval == t_score[i]) { ++rep; lasti = i; }
	}
	loc[0] = loc[1] = loc[2] = loc[3] = 0;
	if (maxval < 5) return 0;
	if (rep == maxval) {
		loc[0] = t_loc[maxi]; loc[1] = t_seedn[maxi];
		*rep_loc = maxi;
		loc[2] = t_loc[lasti]; loc[3] = t_seedn[lasti];
		return 1;
	}
	// the first consistent partner on the left of the best entry (else the best entry itself) becomes the anchor
	for (int j = 0; j <= maxi; ++j) {
		bool take = j == maxi;
		if (!take) {
			const int dl = t_loc[maxi] - t_loc[j], ds = t_seedn[maxi] - t_seedn[j];
			take = ds > 0 && dl > 0 && dl < read_len && ddf_close(dl, ds, bc);
		}
		if (take) {
			if (loc[0] == 0) { loc[0] = t_loc[j]; loc[1] = t_seedn[j]; *rep_loc = j; }
			else { loc[2] = t_loc[j]; loc[3] = t_seedn[j]; }
		}
	}
	for (int j = maxi + 1; j < k; ++j) {
		const int dl = t_loc[j] - t_loc[maxi], ds = t_seedn[j] - t_seedn[maxi];
		if (ds > 0 && dl > 0 && dl <= read_len && ddf_close(dl, ds, bc)) {
			if (loc[0] == 0) { loc[0] = t_loc[j]; loc[1] = t_seedn[j]; *rep_loc = j; }
			else { loc[2] = t_loc[j]; loc[3] = t_seedn[j]; }
		}
	}
	return 1;
}

// Number of index positions the strand's sampled k-mers hit: sizes the strand's table (one record per hit at most).
REF_HD int64_t count_hits(const uint32_t* fwd, uint32_t off, const Unit& u, const uint32_t* ibegin, const int64_t* bad, int64_t nbad)
{
	const int n = sampled_kmers(u.len, u.bc);
	int64_t hits = 0;
	for (int k = 0; k < n; ++k) {
		const int i = k * u.bc;
		if (nbad && !u.rc && covers_bad(bad, nbad, (int64_t)off + i, (int64_t)off + i + SEED)) continue;
		const uint32_t code = strand_kmer(fwd, off, u, i);
		hits += ibegin[code + 1] - ibegin[code];
	}
	return hits;
}

// The seeding loop of reference_mapping (mecat2ref_impl_large.cpp:407-462): every hit of every sampled k-mer, in the
// reference's order (k-mers ascending, positions ascending).  A block accepts one hit per seed ordinal.
REF_HD void seed_strand(const uint32_t* fwd, uint32_t off, const Unit& u, int zv, const uint32_t* ibegin, const int32_t* ipos,
                        const int64_t* bad, int64_t nbad, Table& T)
{
	const int n = sampled_kmers(u.len, u.bc);
	for (int k = 0; k < n; ++k) {
		const int i = k * u.bc;
		if (nbad && !u.rc && covers_bad(bad, nbad, (int64_t)off + i, (int64_t)off + i + SEED)) continue;
		const uint32_t code = strand_kmer(fwd, off, u, i);
		const uint32_t e = ibegin[code + 1];
		for (uint32_t h = ibegin[code]; h < e; ++h) {
			const uint32_t pos = (uint32_t)ipos[h] + 1u;         // the reference's positions are 1-based (:265); below 2^31 here
			const int32_t blk = (int32_t)(pos / (uint32_t)zv);
			const int offs = (int)(pos - (uint32_t)blk * (uint32_t)zv);
			bool fresh;
			Bucket* b = table_touch(T, blk, &fresh);
			if (b->score == 0 || b->seednum < k + 1) {
				const int loc = ++b->score;
				if (loc <= SM) { b->loczhi[loc - 1] = (int16_t)offs; b->seedno[loc - 1] = (int16_t)(k + 1); }
				else insert_loc(b, offs, k + 1, u.bc);
				int s_k = b->score;
				if (blk > 0) { const Bucket* p = table_find(T, blk - 1); if (p) s_k += p->score; }
				b->index_score = (int16_t)s_k;
				b->score2 = b->score;
			}
			b->seednum = (int16_t)(k + 1);
		}
	}
}

// The candidate walk of reference_mapping (:463-614) over the touched blocks in first-touch order.  Writes at most maxc
// candidates, best first, equal scores in walk order; returns their number.
REF_HD int walk_strand(const Unit& u, int zv, int gate, int64_t seqcount, int chain, Table& T, RefCand* cand, int maxc)
{
	int ncand = 0;
	int t_loc[2 * SM], t_seed[2 * SM], t_score[2 * SM];
	const int nrec = T.nrec;
	for (int i = 0; i < nrec; ++i) {
		Bucket* spr = T.recs + i;
		if (!(spr->index_score > gate) || spr->score == 0) continue;
		const int blk = spr->blk;
		const int s_k = spr->score;
		const Bucket* prev = blk > 0 ? table_find(T, blk - 1) : nullptr;
		const int pscore = prev ? prev->score : 0;
		int64_t start_loc = (int64_t)blk * zv;
		int n = 0;
		if (pscore > 0) {
			start_loc = (int64_t)(blk - 1) * zv;
			for (int q = 0; q < pscore && q < SM; ++q) { t_loc[n] = prev->loczhi[q]; t_seed[n] = prev->seedno[q]; ++n; }
			for (int q = 0; q < s_k && q < SM; ++q) { t_loc[n] = spr->loczhi[q] + zv; t_seed[n] = spr->seedno[q]; ++n; }
		} else {
			for (int q = 0; q < s_k && q < SM; ++q) { t_loc[n] = spr->loczhi[q]; t_seed[n] = spr->seedno[q]; ++n; }
		}
		int loc[4], rep = 0;
		if (!find_location(t_loc, t_seed, t_score, loc, n, &rep, u.bc, u.len)) continue;
		if (t_score[rep] < 6) continue;
		int score = t_score[rep];
		const int loc_seed = t_seed[rep];
		const int64_t loc_list = start_loc + loc[0];
		const int qoff = (loc[1] - 1) * u.bc;
		const int64_t left1 = loc_list + SEED - 1, right1 = seqcount - loc_list;
		const int64_t left2 = qoff + SEED - 1, right2 = u.len - qoff;
		const int num1 = (int)(left1 >= left2 ? left2 : left1), num2 = (int)(right1 >= right2 ? right2 : right1);
		// votes of the blocks further left (from blk - 2 down) and right; a block that mostly agrees is consumed
		{
			int64_t bk = (int64_t)blk - 2;
			for (int k = num1 / zv; bk >= 0 && k >= 0; --k, --bk) {
				Bucket* p = table_find(T, (int32_t)bk);
				if (!p || !(p->score > 0)) continue;
				const int64_t sl = bk * zv;
				const int scnt = p->score < SM ? p->score : SM;
				int s = 0;
				for (int q = 0; q < scnt; ++q) if (ddf_close(loc_list - sl - p->loczhi[q], loc_seed - p->seedno[q], u.bc)) ++s;
				score += s;
				if (5 * s > 2 * scnt) p->score = 0;           // s / scnt > 0.4
			}
		}
		{
			int64_t bk = (int64_t)blk + 1;
			for (int k = num2 / zv; k > 0; --k, ++bk) {
				Bucket* p = table_find(T, (int32_t)bk);
				if (!p || !(p->score > 0)) continue;
				const int64_t sl = bk * zv;
				const int scnt = p->score < SM ? p->score : SM;
				int s = 0;
				for (int q = 0; q < scnt; ++q) if (ddf_close(sl + p->loczhi[q] - loc_list, p->seedno[q] - loc_seed, u.bc)) ++s;
				score += s;
				if (5 * s > 2 * scnt) p->score = 0;
			}
		}
		// sorted insert (binary search on the score, equal scores behind), list capped at maxc (:589-611)
		int low = 0, high = ncand - 1;
		while (low <= high) {
			const int mid = (low + high) / 2;
			if (cand[mid].score < score) high = mid - 1; else low = mid + 1;
		}
		const int at = high + 1;
		if (ncand < maxc || at < maxc) {
			const int last = ncand < maxc ? ncand : maxc - 1;
			for (int q = last; q > at; --q) cand[q] = cand[q - 1];
			RefCand c; c.loc1 = (int32_t)loc_list; c.loc2 = qoff; c.score = score; c.chain = chain;
			cand[at] = c;
			if (ncand < maxc) ++ncand;
		}
	}
	return ncand;
}

// ------------------------------------------------------------------------------------------ a warp per strand
// The same seeding loop and candidate walk with the 32 lanes of a warp on one strand (SeedWarpFn).  What is sequential
// in the reference stays sequential -- hits in (k-mer, position) order, blocks in first-touch order -- and is done by
// the leading lane; the lanes share the parts that are wide:
//   * the k-mers of the next 32 ordinals, their index lists and the table slots / records their first hits will touch
//     are fetched by 32 lanes at once, so the leader's dependent chain (slot -> record -> left neighbour) finds them in
//     cache;
//   * insert_loc's 21 x 20 / 2 pair tests (every second hit of a true locus: a block of 1 000 bases receives ~50 seeds,
//     20 are kept) and find_location's k (k - 1) / 2: a lane per entry;
//   * the neighbour votes of the walk: a lane per stored seed.
// `lanes` is the interface of cns_core.cuh plus lead / min_val / max_val; per-warp scratch (WarpScratch) carries what
// the lanes hand each other.  Results are identical to seed_strand + walk_strand (checked strand by strand in the CPU
// suite, tests/test_ref_host.py, and on the device against the thread-per-strand form).
struct WarpScratch
{
	uint32_t lb[32], le[32];                 // index list bounds of the 32 k-mers being fetched
	int t_loc[2 * SM], t_seed[2 * SM], t_score[2 * SM];
	int sink;                                // keeps the prefetching loads alive
};

struct EmuLanes        // the host stand-in of a warp: the lanes run one after the other
{
	static constexpr int count = 32;
	template <class F> void each(F&& f) const { for (int l = 0; l < count; ++l) f(l); }
	template <class F> int sum(F&& f) const { int s = 0; for (int l = 0; l < count; ++l) s += f(l); return s; }
	template <class F> uint32_t ballot(F&& f) const { uint32_t m = 0; for (int l = 0; l < count; ++l) if (f(l)) m |= 1u << l; return m; }
	template <class F> int min_val(F&& f) const { int m = f(0); for (int l = 1; l < count; ++l) { const int v = f(l); if (v < m) m = v; } return m; }
	template <class F> int lead(F&& f) const { return f(); }
	bool leader() const { return true; }
	void sync() const {}
};

REF_HD int low_bit(uint32_t x)       // x != 0
{
#if defined(__CUDA_ARCH__)
	return __ffs((int)x) - 1;
#else
	return __builtin_ctz(x);
#endif
}

// do entries i < j of a block agree?  (the test of insert_loc and find_location without its distance cap)
REF_HD bool pair_agrees(int loc_i, int seed_i, int loc_j, int seed_j, int bc)
{
	return seed_j - seed_i > 0 && loc_j - loc_i > 0 && ddf_close(loc_j - loc_i, seed_j - seed_i, bc);
}

// insert_loc with a lane per entry.  Entry i < SM is the block's i-th seed, entry SM the new one.
template <class L>
REF_HD void insert_loc_w(const L& lanes, Bucket* b, int loc, int seedn, int bc)
{
	auto eloc = [&](int i) { return i < SM ? (int)b->loczhi[i] : loc; };
	auto eseed = [&](int i) { return i < SM ? (int)b->seedno[i] : seedn; };
	auto score_of = [&](int l) {
		if (l > SM) return 10000;
		int s = 0;
		const int li = eloc(l), si = eseed(l);
		for (int j = 0; j < l; ++j) s += pair_agrees(eloc(j), eseed(j), li, si, bc) ? 1 : 0;
		for (int j = l + 1; j <= SM; ++j) s += pair_agrees(li, si, eloc(j), eseed(j), bc) ? 1 : 0;
		return s;
	};
	const int minval = lanes.min_val(score_of);
	const int mini = low_bit(lanes.ballot([&](int l) { return score_of(l) == minval; }));     // the first entry with the lowest score
	lanes.sync();
	if (lanes.leader()) {
		if (minval == SM) { b->loczhi[SM - 1] = (int16_t)loc; b->seedno[SM - 1] = (int16_t)seedn; }
		else if (minval < SM && mini < SM) {
			for (int i = mini; i < SM - 1; ++i) { b->loczhi[i] = b->loczhi[i + 1]; b->seedno[i] = b->seedno[i + 1]; }
			b->loczhi[SM - 1] = (int16_t)loc; b->seedno[SM - 1] = (int16_t)seedn;
			--b->score;
		}
	}
	lanes.sync();
}

template <class L>
REF_HD void seed_strand_w(const L& lanes, const uint32_t* fwd, uint32_t off, const Unit& u, int zv, const uint32_t* ibegin, const int32_t* ipos,
                          const int64_t* bad, int64_t nbad, Table& T, WarpScratch& W)
{
	const int n = sampled_kmers(u.len, u.bc);
	for (int k0 = 0; k0 < n; k0 += 32) {
		// the next 32 k-mers: list bounds, and a first look at what their hits will touch (pulls it into cache)
		lanes.each([&](int l) {
			const int k = k0 + l;
			uint32_t b = 0, e = 0;
			if (k < n) {
				const int i = k * u.bc;
				if (!(nbad && !u.rc && covers_bad(bad, nbad, (int64_t)off + i, (int64_t)off + i + SEED))) {
					const uint32_t code = strand_kmer(fwd, off, u, i);
					b = ibegin[code]; e = ibegin[code + 1];
				}
			}
			W.lb[l] = b; W.le[l] = e;
			int acc = 0;
			for (uint32_t h = b; h < e && h < b + 4; ++h) {
				const uint32_t pos = (uint32_t)ipos[h] + 1u;
				const int32_t blk = (int32_t)(pos / (uint32_t)zv);
				const uint32_t hs = ((uint32_t)(blk + 1) * 2654435761u) >> T.shift;
				const Slot sl = T.slots[hs];
				acc += sl.key;
				if (sl.key == blk + 1) acc += T.recs[sl.rec].score;
				// the left neighbour, whose score goes into the block's index_score
				const uint32_t hn = ((uint32_t)blk * 2654435761u) >> T.shift;
				const Slot sn = T.slots[hn];
				if (blk > 0 && sn.key == blk) acc += T.recs[sn.rec].score;
			}
			if (acc == 0x7fffffff) W.sink = acc;
		});
		lanes.sync();
		for (int j = 0; j < 32 && k0 + j < n; ++j) {
			const int k = k0 + j;
			const uint32_t e = W.le[j];
			for (uint32_t h = W.lb[j]; h < e; ++h) {
				const uint32_t pos = (uint32_t)ipos[h] + 1u;
				const int32_t blk = (int32_t)(pos / (uint32_t)zv);
				const int offs = (int)(pos - (uint32_t)blk * (uint32_t)zv);
				// the leader touches the block and takes the hit if the block has room; 2 * record + 1 asks for an eviction
				const int r = lanes.lead([&]() {
					bool fresh;
					Bucket* b = table_touch(T, blk, &fresh);
					const int rec = (int)(b - T.recs);
					if (!(b->score == 0 || b->seednum < k + 1)) { b->seednum = (int16_t)(k + 1); return 2 * rec; }
					const int loc = ++b->score;
					if (loc > SM) return 2 * rec + 1;
					b->loczhi[loc - 1] = (int16_t)offs; b->seedno[loc - 1] = (int16_t)(k + 1);
					int s_k = b->score;
					if (blk > 0) { const Bucket* p = table_find(T, blk - 1); if (p) s_k += p->score; }
					b->index_score = (int16_t)s_k; b->score2 = b->score; b->seednum = (int16_t)(k + 1);
					return 2 * rec;
				});
				if ((r >> 1) >= T.nrec) T.nrec = (r >> 1) + 1;
				if (r & 1) {
					Bucket* b = T.recs + (r >> 1);
					insert_loc_w(lanes, b, offs, k + 1, u.bc);
					if (lanes.leader()) {
						int s_k = b->score;
						if (blk > 0) { const Bucket* p = table_find(T, blk - 1); if (p) s_k += p->score; }
						b->index_score = (int16_t)s_k; b->score2 = b->score; b->seednum = (int16_t)(k + 1);
					}
					lanes.sync();
				}
			}
		}
		lanes.sync();
	}
}

// find_location over the entries in W.t_loc / W.t_seed: the pair counts with a lane per entry, the rest as it is
template <class L>
REF_HD int find_location_w(const L& lanes, WarpScratch& W, int* loc, int k, int* rep_loc, int bc, int read_len)
{
	lanes.each([&](int l) {
		for (int i = l; i < k; i += L::count) {
			int s = 0;
			const int li = W.t_loc[i], si = W.t_seed[i];
			for (int j = 0; j < k; ++j) {
				if (j == i) continue;
				const int dl = j > i ? W.t_loc[j] - li : li - W.t_loc[j], ds = j > i ? W.t_seed[j] - si : si - W.t_seed[j];
				if (ds > 0 && dl > 0 && dl < read_len && ddf_close(dl, ds, bc)) ++s;
			}
			W.t_score[i] = s;
		}
	});
	lanes.sync();
	// selection of the anchor: find_location's own lines on the finished counts (every lane, same result)
	int maxval = 0, maxi = 0, rep = 0, lasti = 0;
	const int* t_loc = W.t_loc; const int* t_seedn = W.t_seed; const int* t_score = W.t_score;
	for (int i = 0; i < k; ++i) {
		if (maxval < t_score[i]) { maxval = t_score[i]; maxi = i; rep = 0; }
		else if (maxval == t_score[i]) { ++rep; lasti = i; }
	}
	loc[0] = loc[1] = loc[2] = loc[3] = 0;
	if (maxval < 5) return 0;
	if (rep == maxval) {
		loc[0] = t_loc[maxi]; loc[1] = t_seedn[maxi];
		*rep_loc = maxi;
		loc[2] = t_loc[lasti]; loc[3] = t_seedn[lasti];
		return 1;
	}
	for (int j = 0; j <= maxi; ++j) {
		bool take = j == maxi;
		if (!take) {
			const int dl = t_loc[maxi] - t_loc[j], ds = t_seedn[maxi] - t_seedn[j];
			take = ds > 0 && dl > 0 && dl < read_len && ddf_close(dl, ds, bc);
		}
		if (take) {
			if (loc[0] == 0) { loc[0] = t_loc[j]; loc[1] = t_seedn[j]; *rep_loc = j; }
			else { loc[2] = t_loc[j]; loc[3] = t_seedn[j]; }
		}
	}
	for (int j = maxi + 1; j < k; ++j) {
		const int dl = t_loc[j] - t_loc[maxi], ds = t_seedn[j] - t_seedn[maxi];
		if (ds > 0 && dl > 0 && dl <= read_len && ddf_close(dl, ds, bc)) {
			if (loc[0] == 0) { loc[0] = t_loc[j]; loc[1] = t_seedn[j]; *rep_loc = j; }
			else { loc[2] = t_loc[j]; loc[3] = t_seedn[j]; }
		}
	}
	return 1;
}

template <class L>
REF_HD int walk_strand_w(const L& lanes, const Unit& u, int zv, int gate, int64_t seqcount, int chain, Table& T, RefCand* cand, int maxc,
                         WarpScratch& W)
{
	int ncand = 0;
	const int nrec = T.nrec;
	for (int base = 0; base < nrec; base += 32) {
		// 32 records at a time: which of them pass the gate NOW (a vote of an earlier one may have emptied a later one:
		// the score is read again when its turn comes)
		uint32_t pass = lanes.ballot([&](int l) { const int i = base + l; return i < nrec && T.recs[i].index_score > gate; });
		while (pass) {
			const int i = base + low_bit(pass);
			pass &= pass - 1;
			Bucket* spr = T.recs + i;
			if (spr->score == 0) continue;
			const int blk = spr->blk;
			const int s_k = spr->score;
			const Bucket* prev = blk > 0 ? table_find(T, blk - 1) : nullptr;
			const int pscore = prev ? prev->score : 0;
			int64_t start_loc = (int64_t)blk * zv;
			const int np = pscore > 0 ? (pscore < SM ? pscore : SM) : 0, ns = s_k < SM ? s_k : SM;
			if (pscore > 0) start_loc = (int64_t)(blk - 1) * zv;
			const int n = np + ns;
			lanes.each([&](int l) {
				for (int q = l; q < n; q += L::count) {
					if (q < np) { W.t_loc[q] = prev->loczhi[q]; W.t_seed[q] = prev->seedno[q]; }
					else { W.t_loc[q] = spr->loczhi[q - np] + (np ? zv : 0); W.t_seed[q] = spr->seedno[q - np]; }
				}
			});
			lanes.sync();
			int loc[4], rep = 0;
			if (!find_location_w(lanes, W, loc, n, &rep, u.bc, u.len)) { lanes.sync(); continue; }
			const int rep_score = W.t_score[rep], loc_seed = W.t_seed[rep];
			lanes.sync();
			if (rep_score < 6) continue;
			int score = rep_score;
			const int64_t loc_list = start_loc + loc[0];
			const int qoff = (loc[1] - 1) * u.bc;
			const int64_t left1 = loc_list + SEED - 1, right1 = seqcount - loc_list;
			const int64_t left2 = qoff + SEED - 1, right2 = u.len - qoff;
			const int num1 = (int)(left1 >= left2 ? left2 : left1), num2 = (int)(right1 >= right2 ? right2 : right1);
			// neighbour votes: a lane per stored seed of the neighbour
			{
				int64_t bk = (int64_t)blk - 2;
				for (int k = num1 / zv; bk >= 0 && k >= 0; --k, --bk) {
					Bucket* p = table_find(T, (int32_t)bk);
					if (!p || !(p->score > 0)) continue;
					const int64_t sl = bk * zv;
					const int scnt = p->score < SM ? p->score : SM;
					const int s = lanes.sum([&](int l) { return l < scnt && ddf_close(loc_list - sl - p->loczhi[l], loc_seed - p->seedno[l], u.bc) ? 1 : 0; });
					score += s;
					lanes.sync();
					if (5 * s > 2 * scnt && lanes.leader()) p->score = 0;
					lanes.sync();
				}
			}
			{
				int64_t bk = (int64_t)blk + 1;
				for (int k = num2 / zv; k > 0; --k, ++bk) {
					Bucket* p = table_find(T, (int32_t)bk);
					if (!p || !(p->score > 0)) continue;
					const int64_t sl = bk * zv;
					const int scnt = p->score < SM ? p->score : SM;
					const int s = lanes.sum([&](int l) { return l < scnt && ddf_close(sl + p->loczhi[l] - loc_list, p->seedno[l] - loc_seed, u.bc) ? 1 : 0; });
					score += s;
					lanes.sync();
					if (5 * s > 2 * scnt && lanes.leader()) p->score = 0;
					lanes.sync();
				}
			}
			int low = 0, high = ncand - 1;
			while (low <= high) {
				const int mid = (low + high) / 2;
				if (cand[mid].score < score) high = mid - 1; else low = mid + 1;
			}
			const int at = high + 1;
			if (ncand < maxc || at < maxc) {
				const int last = ncand < maxc ? ncand : maxc - 1;
				if (lanes.leader()) {
					for (int q = last; q > at; --q) cand[q] = cand[q - 1];
					RefCand c; c.loc1 = (int32_t)loc_list; c.loc2 = qoff; c.score = score; c.chain = chain;
					cand[at] = c;
				}
				if (ncand < maxc) ++ncand;
				lanes.sync();
			}
		}
	}
	return ncand;
}

// One end of one alignment that stops inside the read (get_left / get_right_clipped_candidate + fill_clipped_candidate,
// mecat2ref_aux.cpp:186-303): the best-filled block beyond the clipped end proposes one more candidate.
struct RescueQuery { int32_t unit, side, qoff, qend, read_len, pad_; int64_t soff, send; };

REF_HD bool rescue_candidate(const RescueQuery& q, const Table& T, int zv, int bc, int64_t ref_size, RefCand* out)
{
	int max_score = 0;
	const Bucket* block = nullptr;
	int64_t bid = -1;
	if (q.side == 0) {
		if (q.qoff <= CLIPPED || q.soff <= CLIPPED) return false;
		const int na = q.qoff / zv, nb = (int)(q.soff / zv);
		int n = na < nb ? na : nb;
		for (int64_t n2 = q.soff / zv - 1; n >= 0 && n2 >= 0; --n, --n2) {
			const Bucket* p = table_find(T, (int32_t)n2);
			if (p && p->score2 > max_score) { max_score = p->score2; block = p; bid = n2; }
		}
	} else {
		if (q.read_len - q.qend <= CLIPPED || ref_size - q.send <= CLIPPED) return false;
		const int na = (q.read_len - q.qend) / zv, nb = (int)((ref_size - q.send) / zv);
		int n = na < nb ? na : nb;
		for (int64_t k = q.send / zv + 1; n >= 0; --n, ++k) {
			const Bucket* p = table_find(T, (int32_t)k);
			if (p && p->score2 > max_score) { max_score = p->score2; block = p; bid = k; }
		}
	}
	if (!block || !(block->score2 > 4)) return false;
	int seedn[SM], boff[SM], score[SM], loc[4], rep = 0;
	const int n = block->score2 < SM ? block->score2 : SM;
	for (int i = 0; i < n; ++i) { seedn[i] = block->seedno[i]; boff[i] = block->loczhi[i]; }
	if (!find_location(boff, seedn, score, loc, n, &rep, bc, q.read_len)) return false;
	out->score = score[rep];
	out->loc1 = (int32_t)(bid * zv + loc[0]);
	out->loc2 = (loc[1] - 1) * bc;
	out->chain = 0;
	return true;
}

}  // namespace mbref
