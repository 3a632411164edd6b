// mecat_b200/csrc/asm_core.cuh -- per-unit bodies of mecat2asmpw / mecat2trimpw: the
// overlapper mecat2canu runs on corrected reads (mecat2canu/src/mecat2asmpw/mecat2asmpw.c; mecat2trimpw.c differs in the
// score gate :640 and the printed score :942-943, the *50.c programs in MAXC :23).
//
// The program works on letters, not on 2-bit volumes: the subject file is one text (reads joined by NUL), its 13-mers
// (A0 T1 C2 G3, lists longer than 256 dropped) index 1-based start positions; a query strand samples a 13-mer every 10
// letters, every hit lands in the block of 1 000 text positions it falls in (up to 60 (offset, ordinal) pairs, one per
// sampled k-mer and block), blocks whose running score passes a gate are scored by pairwise distance consistency,
// neighbouring blocks vote and are retired, and the best MAXC candidates of a read are aligned to both sides of the seed
// in chunks of 500 letters with a banded O(nd) aligner whose edits are gaps only.
//
//   creat_ref_index :397-497   transnum_buchang :316-333   pairwise_mapping :515-984   find_location :338-366
//   align :108-197             string_check :199-281       binary :283-296
//
// The reference keeps a dense array of blocks per thread; here a strand owns an open-addressing table keyed by block
// number over records taken from a shared pool in first-touch order -- the order the reference's index_list walks.
// Where the reference reads block memory that no seed of the strand wrote (the neighbour votes run to a block's score,
// which keeps counting past SM because insert_loc is commented out, :605, :699-712), this code reads zero (DESIGN.md
// section 4.11 says why).  All bodies are shared by the CUDA backend (asmpw.cu) and the host harness
// of the CPU test-suite (tests/asm_host_harness.cpp).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ASM_HD __host__ __device__ __forceinline__
#define ASM_HDN __host__ __device__ __noinline__
#else
#define ASM_HD inline
#define ASM_HDN inline
#endif

namespace mbasm {

constexpr int SEED = 13;            // seed_len, :1080
constexpr int ZV = 1000;            // block width, :18
constexpr int DN = 500;             // letters per alignment chunk, :19
constexpr int BC = 10;              // stride between sampled k-mers, :20
constexpr int SM = 60;              // seeds kept per block, :21
constexpr int MAX_OCC = 256;        // sumvalue_x :307-314
constexpr int MAX_READ = 100000;    // RM: the reference's fixed read buffers, :17
constexpr int MAXC_MAX = 100;       // MAXC :23 (50 in the *50 programs)
constexpr int64_t KMERS = (int64_t)1 << (2 * SEED);
constexpr int TILE = 64;            // k-mer codes per scan tile: two 128-byte lines of counters per unit

ASM_HD uint32_t atomic_inc(uint32_t* p)
{
#if defined(__CUDA_ARCH__)
	return atomicAdd(p, 1u);
#else
	return (*p)++;
#endif
}

ASM_HD int code_of(char c)          // atcttrans :298-304 on upper-cased letters
{
	return c == 'A' ? 0 : c == 'T' ? 1 : c == 'C' ? 2 : c == 'G' ? 3 : 4;
}

// 13-mer whose first letter is text[i]; -1 when a letter is not ACGT (NUL between reads included)
ASM_HD int32_t text_kmer(const char* text, int64_t i, int64_t n)
{
	if (i + SEED > n) return -1;
	int32_t code = 0;
	for (int j = 0; j < SEED; ++j) {
		const int t = code_of(text[i + j]);
		if (t == 4) return -1;
		code = (code << 2) + t;
	}
	return code;
}

// the loaders' view of a file: letters from 'a' up go through toupper (load_read :355, load_fastq :1012), which in the C
// locale changes a .. z only
struct UpperFn
{
	char* text; int64_t n;
	ASM_HD void operator()(int64_t i) const
	{
		const int64_t b = i * 16, e = b + 16 < n ? b + 16 : n;
		for (int64_t j = b; j < e; ++j) { const char c = text[j]; if (c >= 'a' && c <= 'z') text[j] = (char)(c - 32); }
	}
};

// ---------------------------------------------------------------------------------------------- index of the subject text
struct KmerCountFn
{
	const char* text; int64_t n; uint32_t* count;
	ASM_HD void operator()(int64_t i) const
	{
		const int32_t c = text_kmer(text, i, n);
		if (c >= 0) atomic_inc(count + c);
	}
};

ASM_HD uint32_t kept(uint32_t c) { return c > (uint32_t)MAX_OCC ? 0u : c; }

struct TileSumFn
{
	const uint32_t* count; uint32_t* tile_sum;
	ASM_HD void operator()(int64_t t) const
	{
		uint32_t s = 0;
		for (int i = 0; i < TILE; ++i) s += kept(count[t * TILE + i]);
		tile_sum[t] = s;
	}
};

struct TileScanFn          // begin[c] = first slot of code c's list; count[] is cleared for the fill's cursors
{
	uint32_t* count; const uint32_t* tile_sum; uint32_t* begin; const uint32_t* total;
	ASM_HD void operator()(int64_t t) const
	{
		uint32_t run = tile_sum[t];
		for (int i = 0; i < TILE; ++i) {
			const int64_t c = t * TILE + i;
			begin[c] = run; run += kept(count[c]); count[c] = 0;
		}
		if (t == KMERS / TILE - 1) begin[KMERS] = *total;
	}
};

struct KmerFillFn
{
	const char* text; int64_t n; const uint32_t* begin; uint32_t* cursor; int32_t* pos;
	ASM_HD void operator()(int64_t i) const
	{
		const int32_t c = text_kmer(text, i, n);
		if (c < 0 || begin[c + 1] == begin[c]) return;
		pos[begin[c] + atomic_inc(cursor + c)] = (int32_t)(i + 1);      // :485 i + 2 - seed_len with i the last letter
	}
};

struct ListSortFn          // ascending positions inside a list, as a sequential fill leaves them
{
	const uint32_t* begin; int32_t* pos;
	ASM_HD void operator()(int64_t c) const
	{
		const uint32_t b = begin[c], e = begin[c + 1];
		for (uint32_t i = b + 1; i < e; ++i) {
			const int32_t v = pos[i];
			uint32_t j = i;
			while (j > b && pos[j - 1] > v) { pos[j] = pos[j - 1]; --j; }
			pos[j] = v;
		}
	}
};

// ---------------------------------------------------------------------------------------------- query strands
struct Reads               // a set of reads as one text: read r = text[start[r] .. start[r] + len[r]), NUL behind it
{
	const char* text; const int32_t* start; const int32_t* len; int32_t n, first_id;
};

ASM_HD char complement(char c) { return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c; }      // :578-585

struct Strand              // letters of one strand of a query read
{
	const char* fwd; int32_t len, rc;
	ASM_HD char at(int i) const { return rc ? complement(fwd[len - 1 - i]) : fwd[i]; }
	ASM_HD int32_t kmer(int i) const
	{
		if (i + SEED > len) return -1;
		int32_t code = 0;
		for (int j = 0; j < SEED; ++j) {
			const int t = code_of(at(i + j));
			if (t == 4) return -1;
			code = (code << 2) + t;
		}
		return code;
	}
};

ASM_HD int sampled_kmers(int len) { return len < SEED ? 0 : (len - SEED) / BC + 1; }      // transnum_buchang :316-333

struct HitCountFn          // index hits of a strand: bounds its table
{
	Reads q; const uint32_t* begin; int32_t* hits;
	ASM_HD void operator()(int64_t u) const
	{
		const int r = (int)(u >> 1);
		Strand s; s.fwd = q.text + q.start[r]; s.len = q.len[r]; s.rc = (int)(u & 1);
		int64_t h = 0;
		if (s.len < MAX_READ) {
			const int nk = sampled_kmers(s.len);
			for (int k = 0; k < nk; ++k) {
				const int32_t c = s.kmer(k * BC);
				if (c >= 0) h += begin[c + 1] - begin[c];
			}
		}
		hits[u] = (int32_t)(h > 0x7fffffff ? 0x7fffffff : h);
	}
};

struct Entries { int16_t loczhi[SM], seedno[SM]; };      // the 60 (offset, ordinal) pairs of a Back_List :67-70

// One touched block.  Most blocks a strand touches hold a single chance hit, so the record carries its first pair itself
// (24 bytes) and gets the 240 bytes of a full entry array only when a second k-mer arrives.
struct Bucket
{
	int32_t blk, index, ext;          // ext: its Entries in the second pool, -1 while one pair is enough
	int16_t score, seednum, index_score, loc0, seed0, pad_;
};

struct Slot { int32_t key, rec; };    // key = block + 1, 0 = empty

struct Table
{
	Slot* slots; uint32_t mask; int shift;
	Bucket* pool; uint32_t* pool_used; uint32_t pool_cap;
	Entries* big; uint32_t* big_used; uint32_t big_cap;
	int32_t* list; int32_t nrec, cap; // records of this strand in first-touch order; cap: room in list (slots hold twice as many)
	bool full;
};

ASM_HD int table_shift(uint32_t cap) { int s = 32; while (cap > 1) { cap >>= 1; --s; } return s; }

ASM_HD Bucket* table_find(const Table& T, int32_t blk)
{
	if (blk < 0) return nullptr;
	uint32_t h = ((uint32_t)(blk + 1) * 2654435761u) >> T.shift;
	for (;;) {
		const Slot s = T.slots[h];
		if (s.key == blk + 1) return T.pool + s.rec;
		if (s.key == 0) return nullptr;
		h = (h + 1) & T.mask;
	}
}

ASM_HD Bucket* table_touch(Table& T, int32_t blk)
{
	uint32_t h = ((uint32_t)(blk + 1) * 2654435761u) >> T.shift;
	for (;;) {
		const Slot s = T.slots[h];
		if (s.key == blk + 1) return T.pool + s.rec;
		if (s.key == 0) break;
		h = (h + 1) & T.mask;
	}
	if (T.nrec >= T.cap) { T.full = true; return nullptr; }
	const uint32_t id = atomic_inc(T.pool_used);
	if (id >= T.pool_cap) { T.full = true; return nullptr; }
	Slot s; s.key = blk + 1; s.rec = (int32_t)id;
	T.slots[h] = s;
	Bucket* b = T.pool + id;
	b->blk = blk; b->index = T.nrec; b->ext = -1; b->score = 0; b->seednum = 0; b->index_score = 0; b->loc0 = 0; b->seed0 = 0; b->pad_ = 0;
	T.list[T.nrec++] = (int32_t)id;
	return b;
}

ASM_HD int score_of(const Table& T, int32_t blk) { const Bucket* b = table_find(T, blk); return b ? b->score : 0; }

// pair j of a block (j < SM); what was never stored reads zero
ASM_HD int ent_loc(const Table& T, const Bucket* b, int j) { return b->ext >= 0 ? T.big[b->ext].loczhi[j] : j == 0 ? b->loc0 : 0; }
ASM_HD int ent_seed(const Table& T, const Bucket* b, int j) { return b->ext >= 0 ? T.big[b->ext].seedno[j] : j == 0 ? b->seed0 : 0; }
ASM_HD void ent_set_loc(const Table& T, Bucket* b, int j, int v)
{
	if (b->ext >= 0) T.big[b->ext].loczhi[j] = (int16_t)v;
	else if (j == 0) b->loc0 = (int16_t)v;
}
// the pair a k-mer adds as the block's `loc`-th (1-based; only the first SM are kept); false when the second pool ran out
ASM_HD bool ent_store(const Table& T, Bucket* b, int loc, int u, int seedn)
{
	if (loc > SM) return true;
	if (loc == 1 && b->ext < 0) { b->loc0 = (int16_t)u; b->seed0 = (int16_t)seedn; return true; }
	if (b->ext < 0) {
		const uint32_t id = atomic_inc(T.big_used);
		if (id >= T.big_cap) return false;
		Entries* e = T.big + id;
		for (int i = 0; i < SM; ++i) { e->loczhi[i] = 0; e->seedno[i] = 0; }
		e->loczhi[0] = b->loc0; e->seedno[0] = b->seed0;
		b->ext = (int32_t)id;
	}
	T.big[b->ext].loczhi[loc - 1] = (int16_t)u; T.big[b->ext].seedno[loc - 1] = (int16_t)seedn;
	return true;
}

// The 124 shorts of a Back_List as the reference's overflowing loops see them: score, loczhi[60], seedno[60], seednum,
// the two halves of index; `field` may run into the following blocks.  A block without a record is all zero, index -1.
ASM_HD int raw_short(const Table& T, int32_t blk, int field)
{
	blk += field / 124; field %= 124;
	const Bucket* b = table_find(T, blk);
	if (!b) return field >= 122 ? -1 : 0;
	if (field == 0) return b->score;
	if (field <= SM) return ent_loc(T, b, field - 1);
	if (field <= 2 * SM) return ent_seed(T, b, field - 1 - SM);
	if (field == 121) return b->seednum;
	return field == 122 ? (int16_t)(b->index & 0xffff) : (int16_t)(b->index >> 16);
}
ASM_HD int loczhi_at(const Table& T, const Bucket* b, int j) { return j < SM ? ent_loc(T, b, j) : raw_short(T, b->blk, 1 + j); }
ASM_HD int seedno_at(const Table& T, const Bucket* b, int j) { return j < SM ? ent_seed(T, b, j) : raw_short(T, b->blk, 1 + SM + j); }

// |a / (10 b) - 1| < 0.10 in exact integers.  find_location evaluates it with a float quotient, the neighbour votes with
// a double one; |10 b| < 2^18, so a quotient other than exactly 0.9 or 1.1 is many ulps from them and rounding is
// monotone: only equality needs care.  1.1 rounds up in both widths: not close.  0.9 rounds down in float (not close)
// but up in double, where 1 - fl(0.9) < fl(0.10): close.  b == 0 divides by zero: inf or NaN, never close.
// tests/test_asm_host.py compares both with the literal expressions.
ASM_HD bool ddf_close(int a, int b)             // float quotient, :340 :349 :355
{
	if (b > 0) return 9 * b < a && a < 11 * b;
	if (b < 0) return 11 * b < a && a < 9 * b;
	return false;
}
ASM_HD bool ddf_close_d(int a, int b)           // double quotient, :702 :709
{
	if (b > 0) return 9 * b <= a && a < 11 * b;
	if (b < 0) return 11 * b < a && a <= 9 * b;
	return false;
}

// seeding of one strand, :590-627
ASM_HD void seed_strand(const Strand& s, const uint32_t* begin, const int32_t* pos, Table& T)
{
	const int nk = sampled_kmers(s.len);
	for (int k = 0; k < nk && !T.full; ++k) {
		const int32_t c = s.kmer(k * BC);
		if (c < 0) continue;
		const uint32_t e = begin[c + 1];
		for (uint32_t i = begin[c]; i < e; ++i) {
			const int32_t p = pos[i], blk = p / ZV, u = p % ZV;
			Bucket* b = table_touch(T, blk);
			if (!b) return;
			if (b->score == 0 || b->seednum < k + 1) {
				const int loc = ++b->score;
				if (!ent_store(T, b, loc, u, k + 1)) { T.full = true; return; }
				b->index_score = (int16_t)(loc + (blk > 0 ? score_of(T, blk - 1) : 0));
			}
			b->seednum = (int16_t)(k + 1);
		}
	}
}

// find_location :338-366 over the entries of a block and its left neighbour
ASM_HDN int find_location(const int* t_loc, const int* t_seedn, int* t_score, int* loc, int k, int* rep_loc, int read_len)
{
	for (int i = 0; i < k; ++i) t_score[i] = 0;
	for (int i = 0; i < k - 1; ++i) {
		int last = t_seedn[i];
		const int li = t_loc[i], si = t_seedn[i];
		for (int j = i + 1; j < k; ++j) {
			const int ds = t_seedn[j] - si, dl = t_loc[j] - li;
			if (last != t_seedn[j] && ds > 0 && dl > 0 && dl < read_len && ddf_close(dl, ds)) { t_score[i]++; t_score[j]++; last = t_seedn[j]; }
		}
	}
	int maxval = 0, maxi = 0, rep = 0, lasti = 0;
	for (int i = 0; i < k; ++i) {
		if (maxval < t_score[i]) { maxval = t_score[i]; maxi = i; rep = 0; }
		else if (maxval == t_score[i]) { rep++; lasti = i; }
	}
	loc[0] = loc[1] = loc[2] = loc[3] = 0;
	if (maxval < 5) return 0;
	if (rep == maxval) {
		loc[0] = t_loc[maxi]; loc[1] = t_seedn[maxi]; *rep_loc = maxi; loc[2] = t_loc[lasti]; loc[3] = t_seedn[lasti];
		return 1;
	}
	for (int j = 0; j < k; ++j) {
		bool take;
		if (j == maxi) take = true;
		else if (j < maxi) {
			const int ds = t_seedn[maxi] - t_seedn[j], dl = t_loc[maxi] - t_loc[j];
			take = ds > 0 && dl > 0 && dl < read_len && ddf_close(dl, ds);
		} else {
			const int ds = t_seedn[j] - t_seedn[maxi], dl = t_loc[j] - t_loc[maxi];
			take = ds > 0 && dl > 0 && dl <= read_len && ddf_close(dl, ds);
		}
		if (!take) continue;
		if (loc[0] == 0) { loc[0] = t_loc[j]; loc[1] = t_seedn[j]; *rep_loc = j; }
		else { loc[2] = t_loc[j]; loc[3] = t_seedn[j]; }
	}
	return 1;
}

struct Cand                // canidate_save :61-64
{
	int32_t loc1, loc2, left1, left2, right1, right2, score, num1, num2, readno, readstart, chain;
};

ASM_HD int read_of(const int32_t* start, int n, int key)      // binary :283-296: the last read starting at or before key
{
	int lo = 0, hi = n;
	while (lo < hi) { const int m = (lo + hi) >> 1; if (start[m] <= key) lo = m + 1; else hi = m; }
	return lo - 1;
}

// candidate walk of one strand, :638-726.  `out` holds up to maxc candidates in order of score, a newcomer behind its equals.
ASM_HDN int walk_strand(const Strand& s, int read_name, int chain, const Reads& sub, Table& T, int gate, Cand* out, int maxc)
{
	int ncand = 0;
	int t_list[2 * SM], t_seedn[2 * SM], t_score[2 * SM];
	for (int e = 0; e < T.nrec; ++e) {
		Bucket* b = T.pool + T.list[e];
		if (!(b->index_score > gate) || b->score == 0) continue;
		const int blk = b->blk, s_k = b->score;
		int start_loc = blk * ZV, u_k = 0;
		const Bucket* a = blk > 0 ? table_find(T, blk - 1) : nullptr;
		const int loc = a ? a->score : 0;
		if (loc > 0) {
			start_loc = (blk - 1) * ZV;
			for (int j = 0; j < loc && j < SM; ++j) { t_list[u_k] = ent_loc(T, a, j); t_seedn[u_k] = ent_seed(T, a, j); u_k++; }
			for (int j = 0; j < s_k && j < SM; ++j) { t_list[u_k] = ent_loc(T, b, j) + ZV; t_seedn[u_k] = ent_seed(T, b, j); u_k++; }
		} else
			for (int j = 0; j < s_k && j < SM; ++j) { t_list[u_k] = ent_loc(T, b, j); t_seedn[u_k] = ent_seed(T, b, j); u_k++; }
		int location[4], rep_loc = 0;
		if (!find_location(t_list, t_seedn, t_score, location, u_k, &rep_loc, s.len)) continue;
		if (t_score[rep_loc] < 6) continue;
		Cand c;
		c.score = t_score[rep_loc];
		const int loc_seed = t_seedn[rep_loc];
		location[0] += start_loc;
		const int loc_list = location[0];
		const int readno = read_of(sub.start, sub.n, loc_list);
		const int readstart = sub.start[readno], readend = readno + 1 < sub.n ? sub.start[readno + 1] : 0;      // llocation[n] is never written, :360-372
		if (sub.first_id + readno > read_name) continue;
		if (sub.first_id + readno == read_name) {              // :666-673: the read's own letters leave the table
			int u = readstart / ZV, sk = readstart % ZV, k = 0;
			Bucket* t = table_find(T, u);
			if (t) { for (int j = 0; j < t->score && j < SM; ++j) if (ent_loc(T, t, j) < sk) ent_set_loc(T, t, k++, ent_loc(T, t, j)); t->score = (int16_t)k; }
			const int kend = readend / ZV;
			for (++u; u < kend; ++u) { t = table_find(T, u); if (t) t->score = 0; }
			t = table_find(T, u);
			k = 0; sk = readend % ZV;
			if (t) { for (int j = 0; j < t->score && j < SM; ++j) if (ent_loc(T, t, j) > sk) ent_set_loc(T, t, k++, ent_loc(T, t, j)); t->score = (int16_t)k; }
			continue;
		}
		c.readno = readno; c.readstart = readstart;
		location[1] = (location[1] - 1) * BC;
		c.left1 = location[0] - readstart + SEED - 1; c.right1 = readend - location[0];
		c.left2 = location[1] + SEED - 1; c.right2 = s.len - location[1];
		c.num1 = c.left1 >= c.left2 ? c.left2 : c.left1;
		c.num2 = c.right1 >= c.right2 ? c.right2 : c.right1;
		if (c.num1 + c.num2 < 400) continue;
		c.loc1 = location[0]; c.loc2 = location[1];
		int seedcount = 0;
		{   // neighbour votes :699-712: blocks further out on the same diagonal add to the score and are retired
			int u = blk - 2;
			for (int k = c.num1 / ZV; u >= 0 && k >= 0; --u, --k) {
				Bucket* t = table_find(T, u);
				if (!t || t->score <= 0) continue;
				const int st = u * ZV;
				int hit = 0;
				for (int j = 0; j < t->score; ++j) if (ddf_close_d(loc_list - st - loczhi_at(T, t, j), loc_seed - seedno_at(T, t, j))) hit++;
				seedcount += hit;
				if (5 * hit > 2 * t->score) t->score = 0;       // hit * 1.0 / score > 0.4
			}
			u = blk + 1;
			for (int k = c.num2 / ZV; k > 0; ++u, --k) {
				Bucket* t = table_find(T, u);
				if (!t || t->score <= 0) continue;
				const int st = u * ZV;
				int hit = 0;
				for (int j = 0; j < t->score; ++j) if (ddf_close_d(st + loczhi_at(T, t, j) - loc_list, seedno_at(T, t, j) - loc_seed)) hit++;
				seedcount += hit;
				if (5 * hit > 2 * t->score) t->score = 0;
			}
		}
		c.score += seedcount;
		c.chain = chain;
		int at = ncand;
		while (at > 0 && out[at - 1].score < c.score) --at;
		if (at < maxc) {
			const int last = ncand < maxc ? ncand : maxc - 1;
			for (int i = last; i > at; --i) out[i] = out[i - 1];
			out[at] = c;
			if (ncand < maxc) ++ncand;
		}
	}
	return ncand;
}

struct TableRefs           // where the table of strand u lives
{
	const int64_t* slot_off; const int64_t* list_off; Slot* slots; int32_t* lists; Bucket* pool; uint32_t* pool_used; uint32_t pool_cap;
	Entries* big; uint32_t* big_used; uint32_t big_cap;
	ASM_HD Table open(int64_t u) const
	{
		Table T;
		const uint32_t cap = (uint32_t)(slot_off[u + 1] - slot_off[u]);
		T.slots = slots + slot_off[u]; T.mask = cap - 1; T.shift = table_shift(cap);
		T.pool = pool; T.pool_used = pool_used; T.pool_cap = pool_cap;
		T.big = big; T.big_used = big_used; T.big_cap = big_cap;
		T.list = lists + list_off[u]; T.nrec = 0; T.cap = (int32_t)(list_off[u + 1] - list_off[u]); T.full = false;
		return T;
	}
};

struct SeedFn              // one strand: block table, candidate walk
{
	Reads q, sub; const int32_t* units; const uint32_t* begin; const int32_t* pos; TableRefs tab; int gate, maxc;
	Cand* cands; int32_t* ncand; int32_t* status;
	ASM_HD void operator()(int64_t i) const
	{
		const int64_t u = units[i];
		const int r = (int)(u >> 1);
		Strand s; s.fwd = q.text + q.start[r]; s.len = q.len[r]; s.rc = (int)(u & 1);
		ncand[u] = 0;
		if (s.len >= MAX_READ) { status[i] = 2; return; }
		Table T = tab.open(i);
		seed_strand(s, begin, pos, T);
		if (T.full) { status[i] = 1; return; }
		status[i] = 0;
		ncand[u] = walk_strand(s, q.first_id + r, s.rc, sub, T, gate, cands + u * maxc, maxc);
	}
};

// ---------------------------------------------------------------------------------------------- a warp per strand
// The same strand on the 32 lanes of a warp.  `lanes` is the warp (asmpw.cu) or its host stand-in (EmuLanes): each(f)
// runs f(lane) on every lane, ballot / sum combine a value over the lanes, lead(f) has one lane compute a value for all,
// sync() orders the lanes' memory accesses.  Code outside these calls is the same for every lane (on the device all lanes
// execute it with the same values; stores there are the leader's).
struct WarpScratch
{
	int32_t blk[32], off[32], rec[32];
	int t_loc[2 * SM], t_seedn[2 * SM], t_score[2 * SM];
};

struct EmuLanes            // the host stand-in of a warp: the lanes run one after the other
{
	template <class F> void each(F&& f) const { for (int l = 0; l < 32; ++l) f(l); }
	template <class F> int sum(F&& f) const { int s = 0; for (int l = 0; l < 32; ++l) s += f(l); return s; }
	template <class F> uint32_t ballot(F&& f) const { uint32_t m = 0; for (int l = 0; l < 32; ++l) if (f(l)) m |= 1u << l; return m; }
	template <class F> int lead(F&& f) const { return f(); }
	bool leader() const { return true; }
	void sync() const {}
};

ASM_HD int popcount32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
	return __popc(x);
#else
	return __builtin_popcount(x);
#endif
}

ASM_HD uint32_t atomic_add(uint32_t* p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
	return atomicAdd(p, v);
#else
	const uint32_t o = *p; *p += v; return o;
#endif
}

ASM_HD void add_int(int* p, int v)
{
#if defined(__CUDA_ARCH__)
	atomicAdd(p, v);
#else
	*p += v;
#endif
}

ASM_HD bool slot_claim(Slot* s, int32_t key)      // does the empty slot become this key's?
{
#if defined(__CUDA_ARCH__)
	return atomicCAS(&s->key, 0, key) == 0;
#else
	if (s->key) return false;
	s->key = key;
	return true;
#endif
}

// seed_strand with a lane per hit of the sampled k-mer's list (ascending positions).  Only the first hit of a k-mer in a
// block changes the block (:599), so the acting lanes of a round hold distinct blocks: new records are numbered in lane
// order -- first-touch order -- and the running score of a block and its left neighbour (:608) is read after every
// lane's increment, which is what the sequential walk sees because the neighbour's hit, if any, lies at a lower position.
template <class L>
ASM_HD void seed_strand_w(const L& lanes, const Strand& s, const uint32_t* begin, const int32_t* pos, Table& T, WarpScratch& W)
{
	const int nk = sampled_kmers(s.len);
	for (int k = 0; k < nk; ++k) {
		const int32_t c = s.kmer(k * BC);
		if (c < 0) continue;
		const uint32_t e = begin[c + 1];
		int32_t carry = -1;
		for (uint32_t base = begin[c]; base < e; base += 32) {
			lanes.each([&](int l) {
				const uint32_t i = base + (uint32_t)l;
				if (i < e) { const int32_t p = pos[i]; W.blk[l] = p / ZV; W.off[l] = p % ZV; }
				else W.blk[l] = -1;
			});
			lanes.sync();
			const uint32_t act = lanes.ballot([&](int l) { return W.blk[l] >= 0 && W.blk[l] != (l ? W.blk[l - 1] : carry); });
			lanes.each([&](int l) {
				if (!(act >> l & 1u)) return;
				const Bucket* b = table_find(T, W.blk[l]);
				W.rec[l] = b ? (int32_t)(b - T.pool) : -1;
			});
			lanes.sync();
			const uint32_t fresh = lanes.ballot([&](int l) { return (act >> l & 1u) && W.rec[l] < 0; });
			if (fresh) {
				const int n = popcount32(fresh);
				if (T.nrec + n > T.cap) { T.full = true; return; }
				const uint32_t first = (uint32_t)lanes.lead([&]() { return (int)atomic_add(T.pool_used, (uint32_t)n); });
				if (first + (uint32_t)n > T.pool_cap) { T.full = true; return; }
				lanes.each([&](int l) {
					if (!(fresh >> l & 1u)) return;
					const int rank = popcount32(fresh & ((1u << l) - 1u));
					const int32_t id = (int32_t)first + rank, blk = W.blk[l];
					Bucket* b = T.pool + id;
					b->blk = blk; b->index = T.nrec + rank; b->ext = -1; b->score = 0; b->seednum = 0; b->index_score = 0; b->loc0 = 0; b->seed0 = 0; b->pad_ = 0;
					T.list[T.nrec + rank] = id;
					uint32_t h = ((uint32_t)(blk + 1) * 2654435761u) >> T.shift;
					while (!slot_claim(T.slots + h, blk + 1)) h = (h + 1) & T.mask;
					T.slots[h].rec = id;
					W.rec[l] = id;
				});
				T.nrec += n;
				lanes.sync();
			}
			lanes.each([&](int l) {
				if (!(act >> l & 1u)) return;
				Bucket* b = T.pool + W.rec[l];
				const int loc = ++b->score;
				if (!ent_store(T, b, loc, W.off[l], k + 1)) W.rec[l] = -2;
				b->seednum = (int16_t)(k + 1);
			});
			lanes.sync();
			if (lanes.ballot([&](int l) { return (act >> l & 1u) && W.rec[l] == -2; })) { T.full = true; return; }
			lanes.each([&](int l) {
				if (!(act >> l & 1u)) return;
				Bucket* b = T.pool + W.rec[l];
				b->index_score = (int16_t)(b->score + (W.blk[l] > 0 ? score_of(T, W.blk[l] - 1) : 0));
			});
			carry = W.blk[31];
			lanes.sync();
		}
	}
}

// walk_strand on a warp: the pair tests of find_location with a lane per first entry, the neighbour votes with a lane
// per entry; the order of the blocks, the retiring of neighbours and the candidate list stay one lane's.
template <class L>
ASM_HD int walk_strand_w(const L& lanes, const Strand& s, int read_name, int chain, const Reads& sub, Table& T, int gate, Cand* out, int maxc, WarpScratch& W)
{
	int ncand = 0;
	for (int e = 0; e < T.nrec; ++e) {
		Bucket* b = T.pool + T.list[e];
		if (!(b->index_score > gate) || b->score == 0) continue;
		const int blk = b->blk, s_k = b->score;
		const Bucket* a = blk > 0 ? table_find(T, blk - 1) : nullptr;
		const int loc = a ? a->score : 0;
		const int na = loc > 0 ? (loc < SM ? loc : SM) : 0, nb = s_k < SM ? s_k : SM, u_k = na + nb;
		const int start_loc = loc > 0 ? (blk - 1) * ZV : blk * ZV;
		lanes.sync();
		lanes.each([&](int l) {
			for (int i = l; i < u_k; i += 32) {
				if (i < na) { W.t_loc[i] = ent_loc(T, a, i); W.t_seedn[i] = ent_seed(T, a, i); }
				else { W.t_loc[i] = ent_loc(T, b, i - na) + (na ? ZV : 0); W.t_seedn[i] = ent_seed(T, b, i - na); }
				W.t_score[i] = 0;
			}
		});
		lanes.sync();
		lanes.each([&](int l) {
			for (int i = l; i < u_k - 1; i += 32) {
				int last = W.t_seedn[i], mine = 0;
				const int li = W.t_loc[i], si = W.t_seedn[i];
				for (int j = i + 1; j < u_k; ++j) {
					const int sj = W.t_seedn[j], ds = sj - si, dl = W.t_loc[j] - li;
					if (last != sj && ds > 0 && dl > 0 && dl < s.len && ddf_close(dl, ds)) { mine++; add_int(W.t_score + j, 1); last = sj; }
				}
				if (mine) add_int(W.t_score + i, mine);
			}
		});
		lanes.sync();
		// the choice of find_location :343-365 over the scores
		int location[4] = {0, 0, 0, 0}, rep_loc = 0;
		{
			int maxval = 0, maxi = 0, rep = 0, lasti = 0;
			for (int i = 0; i < u_k; ++i) {
				const int v = W.t_score[i];
				if (maxval < v) { maxval = v; maxi = i; rep = 0; }
				else if (maxval == v) { rep++; lasti = i; }
			}
			if (maxval < 5) continue;
			if (rep == maxval) {
				location[0] = W.t_loc[maxi]; location[1] = W.t_seedn[maxi]; rep_loc = maxi; location[2] = W.t_loc[lasti]; location[3] = W.t_seedn[lasti];
			} else {
				const int lm = W.t_loc[maxi], sm = W.t_seedn[maxi];
				for (int j = 0; j < u_k; ++j) {
					bool take;
					if (j == maxi) take = true;
					else if (j < maxi) { const int ds = sm - W.t_seedn[j], dl = lm - W.t_loc[j]; take = ds > 0 && dl > 0 && dl < s.len && ddf_close(dl, ds); }
					else { const int ds = W.t_seedn[j] - sm, dl = W.t_loc[j] - lm; take = ds > 0 && dl > 0 && dl <= s.len && ddf_close(dl, ds); }
					if (!take) continue;
					if (location[0] == 0) { location[0] = W.t_loc[j]; location[1] = W.t_seedn[j]; rep_loc = j; }
					else { location[2] = W.t_loc[j]; location[3] = W.t_seedn[j]; }
				}
			}
		}
		if (W.t_score[rep_loc] < 6) continue;
		Cand c;
		c.score = W.t_score[rep_loc];
		const int loc_seed = W.t_seedn[rep_loc];
		location[0] += start_loc;
		const int loc_list = location[0];
		const int readno = read_of(sub.start, sub.n, loc_list);
		const int readstart = sub.start[readno], readend = readno + 1 < sub.n ? sub.start[readno + 1] : 0;
		if (sub.first_id + readno > read_name) continue;
		if (sub.first_id + readno == read_name) {
			if (lanes.leader()) {
				int u = readstart / ZV, sk = readstart % ZV, k = 0;
				Bucket* t = table_find(T, u);
				if (t) { for (int j = 0; j < t->score && j < SM; ++j) if (ent_loc(T, t, j) < sk) ent_set_loc(T, t, k++, ent_loc(T, t, j)); t->score = (int16_t)k; }
				const int kend = readend / ZV;
				for (++u; u < kend; ++u) { t = table_find(T, u); if (t) t->score = 0; }
				t = table_find(T, u);
				k = 0; sk = readend % ZV;
				if (t) { for (int j = 0; j < t->score && j < SM; ++j) if (ent_loc(T, t, j) > sk) ent_set_loc(T, t, k++, ent_loc(T, t, j)); t->score = (int16_t)k; }
			}
			lanes.sync();
			continue;
		}
		c.readno = readno; c.readstart = readstart;
		location[1] = (location[1] - 1) * BC;
		c.left1 = location[0] - readstart + SEED - 1; c.right1 = readend - location[0];
		c.left2 = location[1] + SEED - 1; c.right2 = s.len - location[1];
		c.num1 = c.left1 >= c.left2 ? c.left2 : c.left1;
		c.num2 = c.right1 >= c.right2 ? c.right2 : c.right1;
		if (c.num1 + c.num2 < 400) continue;
		c.loc1 = location[0]; c.loc2 = location[1];
		int seedcount = 0;
		for (int side = 0; side < 2; ++side) {
			int u = side ? blk + 1 : blk - 2;
			for (int k = (side ? c.num2 : c.num1) / ZV; side ? k > 0 : (u >= 0 && k >= 0); u += side ? 1 : -1, --k) {
				Bucket* t = table_find(T, u);
				const int sc = t ? t->score : 0;
				if (sc <= 0) continue;
				const int st = u * ZV;
				const int hit = lanes.sum([&](int l) {
					int h = 0;
					for (int j = l; j < sc; j += 32) {
						const int lo = loczhi_at(T, t, j), se = seedno_at(T, t, j);
						if (side ? ddf_close_d(st + lo - loc_list, se - loc_seed) : ddf_close_d(loc_list - st - lo, loc_seed - se)) h++;
					}
					return h;
				});
				seedcount += hit;
				if (5 * hit > 2 * sc) { if (lanes.leader()) t->score = 0; lanes.sync(); }
			}
		}
		c.score += seedcount;
		c.chain = chain;
		ncand = lanes.lead([&]() {
			int at = ncand;
			while (at > 0 && out[at - 1].score < c.score) --at;
			if (at >= maxc) return ncand;
			const int last = ncand < maxc ? ncand : maxc - 1;
			for (int i = last; i > at; --i) out[i] = out[i - 1];
			out[at] = c;
			return ncand < maxc ? ncand + 1 : ncand;
		});
	}
	return ncand;
}

struct SeedWarpFn          // SeedFn on a warp
{
	Reads q, sub; const int32_t* units; const uint32_t* begin; const int32_t* pos; TableRefs tab; int gate, maxc;
	Cand* cands; int32_t* ncand; int32_t* status;
	template <class L>
	ASM_HD void operator()(int64_t i, const L& lanes, WarpScratch& W) const
	{
		const int64_t u = units[i];
		const int r = (int)(u >> 1);
		Strand s; s.fwd = q.text + q.start[r]; s.len = q.len[r]; s.rc = (int)(u & 1);
		if (s.len >= MAX_READ) { if (lanes.leader()) { ncand[u] = 0; status[i] = 2; } return; }
		Table T = tab.open(i);
		seed_strand_w(lanes, s, begin, pos, T, W);
		if (T.full) { if (lanes.leader()) { ncand[u] = 0; status[i] = 1; } return; }
		const int n = walk_strand_w(lanes, s, q.first_id + r, s.rc, sub, T, gate, cands + u * maxc, maxc, W);
		if (lanes.leader()) { ncand[u] = n; status[i] = 0; }
	}
};

struct MergeFn             // :716-726 across both strands: forward candidates first among equals, cut at MAXC
{
	const Cand* cands; const int32_t* ncand; int maxc; Cand* merged; int32_t* nmerged;
	ASM_HD void operator()(int64_t r) const
	{
		const Cand* f = cands + (2 * r) * maxc; const Cand* v = cands + (2 * r + 1) * maxc;
		const int nf = ncand[2 * r], nv = ncand[2 * r + 1];
		int i = 0, j = 0, n = 0;
		Cand* o = merged + r * maxc;
		while (n < maxc && (i < nf || j < nv)) {
			if (j >= nv || (i < nf && f[i].score >= v[j].score)) o[n++] = f[i++];
			else o[n++] = v[j++];
		}
		nmerged[r] = n;
	}
};

// ---------------------------------------------------------------------------------------------- alignment of a candidate
struct Overlap             // one printed line, :944-945
{
	int32_t sread, qread; float score; int32_t sbeg, send, slen, strand, qbeg, qend, qlen;
};

constexpr int MAX_CHUNK = 600;                        // a side is cut into chunks of DN while more than 600 letters are left
constexpr int MAX_D = 2 * MAX_CHUNK / 10;             // (int)(ErrorRate * (q_len + t_len))
constexpr int VU_INTS = 2 * MAX_D + 4;
constexpr int DP_ROW = MAX_D + 4;                     // diagonals of a row: the band is at most 2 * (int)(0.1 n) + 2 wide
constexpr int DP_CELLS = MAX_D * (DP_ROW / 2 + 1);
constexpr int CHUNK_COLS = 2 * MAX_CHUNK + 8;

struct ExtScratch          // per resident warp
{
	int32_t* V; int32_t* U;            // VU_INTS each
	uint32_t* dp;                      // DP_CELLS: x1 | x2 << 10 | (came from k + 1) << 20
	int32_t* row_start; int32_t* row_min;      // MAX_D + 1 each
	char* cq; char* ct;                // CHUNK_COLS each: the columns of one chunk
	char* l1; char* l2; int64_t lcap;  // the left half, all chunks
};
constexpr int64_t EXT_FIXED_BYTES = 4 * (2 * VU_INTS + DP_CELLS + 2 * (MAX_D + 1)) + 2 * CHUNK_COLS;

struct Side                // letters of both sequences walking away from the seed
{
	const char* text; int64_t p1; Strand q; int p2; int step;
	ASM_HD char s(int i) const { return text[p1 + (int64_t)step * i]; }
	ASM_HD char r(int i) const { return q.at(p2 + step * i); }
};

ASM_HD int trailing_ones(uint32_t m)      // how many low bits of m are set
{
	const uint32_t z = ~m;
	if (!z) return 32;
#if defined(__CUDA_ARCH__)
	return __ffs((int)z) - 1;
#else
	return __builtin_ctz(z);
#endif
}

// align :108-197 on n letters of both sequences, a warp per alignment: the diagonals of a row are walked one after the
// other by every lane alike (the leader stores), the slide along a diagonal compares 32 letters per step, the columns of a
// run of agreeing letters are written by a lane each.  Returns 1 when an end was reached; *cols columns in S.cq / S.ct
// (subject row, query row), *qe / *te letters of each consumed.
template <class L>
ASM_HD int align_chunk(const L& lanes, const Side& W, int n, ExtScratch& S, int* cols, int* qe, int* te)
{
	const int max_d = (int)(0.10 * (n + n));
	const int band_tol = (int)(0.10 * n), band = 2 * band_tol, off = max_d;
	*cols = 0; *qe = 0; *te = 0;
	lanes.sync();
	lanes.each([&](int l) { for (int i = l; i < 2 * max_d + 3; i += 32) { S.V[i] = 0; S.U[i] = 0; } });
	lanes.sync();
	int best_m = -1, min_k = 0, max_k = 0, ncell = 0;
	for (int d = 0; d < max_d; ++d) {
		if (max_k - min_k > band) break;
		if (lanes.leader()) { S.row_start[d] = ncell; S.row_min[d] = min_k; }
		int x = 0, y = 0, k;
		bool aligned = false;
		for (k = min_k; k <= max_k; k += 2) {
			uint32_t down;
			if (k == min_k || (k != max_k && S.V[k - 1 + off] < S.V[k + 1 + off])) { down = 1; x = S.V[k + 1 + off]; }
			else { down = 0; x = S.V[k - 1 + off] + 1; }
			y = x - k;
			const int x1 = x;
			for (;;) {
				const uint32_t m = lanes.ballot([&](int l) { return x + l < n && y + l < n && W.s(x + l) == W.r(y + l); });
				const int run = trailing_ones(m);
				x += run; y += run;
				if (run < 32) break;
			}
			if (lanes.leader()) {
				S.dp[ncell] = (uint32_t)x1 | ((uint32_t)x << 10) | (down << 20);
				S.V[k + off] = x; S.U[k + off] = x + y;      // this row writes one parity of diagonals and reads the other
			}
			++ncell;
			if (x + y > best_m) best_m = x + y;
			if (x >= n || y >= n) { aligned = true; break; }
		}
		lanes.sync();
		if (aligned) {
			*qe = x; *te = y;
			int pos = (x + y + d) / 2;
			*cols = pos;
			int ck = k;
			for (int cd = d; cd >= 0; --cd) {
				const uint32_t e = S.dp[S.row_start[cd] + (ck - S.row_min[cd]) / 2];
				const int x1 = (int)(e & 1023u), x2 = (int)((e >> 10) & 1023u);
				const int len = x2 - x1, first = pos - len;
				lanes.each([&](int l) { for (int i = l; i < len; i += 32) { S.cq[first + i] = W.s(x1 + i); S.ct[first + i] = W.r(x1 + i - ck); } });
				pos = first;
				if (cd == 0) break;
				--pos;
				if ((e >> 20) & 1u) { if (lanes.leader()) { S.cq[pos] = '-'; S.ct[pos] = W.r(x1 - ck - 1); } ck += 1; }
				else { if (lanes.leader()) { S.cq[pos] = W.s(x1 - 1); S.ct[pos] = '-'; } ck -= 1; }
			}
			lanes.sync();
			return 1;
		}
		int new_min = max_k, new_max = min_k;
		for (int k2 = min_k; k2 <= max_k; k2 += 2)
			if (S.U[k2 + off] >= best_m - band_tol) { if (k2 < new_min) new_min = k2; if (k2 > new_max) new_max = k2; }
		max_k = new_max + 1; min_k = new_min - 1;
	}
	return n == 0 ? 1 : 0;       // two empty strings count as aligned, :196
}

struct SideResult { int done1, done2, cols, gaps, gaps_head; bool overflow; };

// one side of the seed, :733-789 / :791-846.  keep: the columns are appended to S.l1 / S.l2 (left half); otherwise only
// counted (gap columns, and those among the first SEED columns of the side).
template <class L>
ASM_HD SideResult extend_side(const L& lanes, Side W, int num, int len1, int len2, ExtScratch& S, bool keep)
{
	SideResult R; R.done1 = R.done2 = R.cols = R.gaps = R.gaps_head = 0; R.overflow = false;
	bool more = true;
	while (more) {
		int n;
		if (num > MAX_CHUNK) n = DN; else { more = false; n = num < 0 ? 0 : num; }
		int cols, qe, te;
		const int ok = align_chunk(lanes, W, n, S, &cols, &qe, &te);
		if (!ok) break;
		int take = cols, adv1, adv2;
		if (more) {
			int k, loc = 0, sci = 0, run = 0;          // back to the last run of four agreeing columns, :748-753
			for (k = cols - 1; k > -1 && run < 4; --k) {
				const char a = S.cq[k], b = S.ct[k];
				if (a != '-') loc++;
				if (b != '-') sci++;
				if (a == b) run++; else run = 0;
			}
			loc = DN - qe + loc; sci = DN - te + sci;
			if (loc == DN) break;
			take = k + 1; adv1 = DN - loc; adv2 = DN - sci;
		} else {
			if (num - qe == num) break;
			adv1 = qe; adv2 = te;
		}
		if (keep && R.cols + take > S.lcap) { R.overflow = true; break; }
		const int at = R.cols;
		const int packed = lanes.sum([&](int l) {
			int g = 0, gh = 0;
			for (int i = l; i < take; i += 32) {
				const char a = S.cq[i], b = S.ct[i];
				if (a != b) { g++; if (at + i < SEED) gh++; }
				if (keep) { S.l1[at + i] = a; S.l2[at + i] = b; }
			}
			return g | (gh << 16);
		});
		R.gaps += packed & 0xffff; R.gaps_head += packed >> 16;
		R.cols += take; R.done1 += adv1; R.done2 += adv2;
		W.p1 += (int64_t)W.step * adv1; W.p2 += W.step * adv2;
		num = len1 - R.done1 >= len2 - R.done2 ? len2 - R.done2 : len1 - R.done1;
	}
	return R;
}

// string_check :199-281 on the left half (str1 / str2, n columns), walking the columns from their far end: a gap column
// whose pending letters agree pulls the run of agreeing letters over.  The letters without gaps are the sequences
// themselves walking left from the seed, seq1[i] = W.s(i), seq2[i] = W.r(i), last indices len1 / len2.
struct GapShifter
{
	Side W; char* str1; char* str2;
	// letters a, a-1, ... of seq1 against b, b-1, ... of seq2: when the first pair agrees the run moves into the columns
	// col, col-1, ... and as many letters left of them are blanked in both rows
	ASM_HD bool pull(int col, int a, int b) const
	{
		if (a < 0 || b < 0 || W.s(a) != W.r(b)) return false;      // :257 may look one place before seq1: never a letter
		int k = 1;
		while (a - k >= 0 && b - k >= 0 && W.s(a - k) == W.r(b - k)) k++;
		int s = 0, j = col;
		while (s < k && j >= 0) { if (str1[j] != '-') { str1[j] = '-'; s++; } j--; }
		s = 0; j = col;
		while (s < k && j >= 0) { if (str2[j] != '-') { str2[j] = '-'; s++; } j--; }
		for (s = 0, j = col; s < k && j >= 0; --j, ++s) { str1[j] = W.s(a - s); str2[j] = W.r(b - s); }
		return true;
	}
	// one column of the walk; only a column with a gap can change anything but the two counters
	ASM_HD void column(int col, int len1, int len2, int& loc1, int& loc2) const
	{
		if (str1[col] != '-') loc1++;
		else if (pull(col, len1 - loc1, len2 - loc2)) { if (str1[col] != '-') loc1++; }
		if (str2[col] != '-') loc2++;
		else if (str1[col] != '-') { if (pull(col, len1 - loc1 + 1, len2 - loc2) && str2[col] != '-') loc2++; }
		else if (pull(col, len1 - loc1, len2 - loc2) && str2[col] != '-') loc2++;
	}
};

// The lanes look at 32 columns at a time; columns without a gap only advance the counters, the first column with one is
// walked by the leader (it may rewrite columns to its left, so the next look starts right behind it).
template <class L>
ASM_HD void shift_gaps(const L& lanes, const Side& W, int len1, int len2, char* str1, char* str2, int n)
{
	GapShifter G; G.W = W; G.str1 = str1; G.str2 = str2;
	int loc1 = 0, loc2 = 0, col = n - 1;
	lanes.sync();
	while (col > -1) {
		const uint32_t plain = lanes.ballot([&](int l) { return col - l > -1 && str1[col - l] != '-' && str2[col - l] != '-'; });
		const int run = trailing_ones(plain);
		loc1 += run; loc2 += run; col -= run;
		if (run == 32 || col < 0) continue;
		const int packed = lanes.lead([&]() {
			int a = loc1, b = loc2;
			G.column(col, len1, len2, a, b);
			return (a - loc1) | ((b - loc2) << 8);
		});
		loc1 += packed & 0xff; loc2 += packed >> 8;
		--col;
	}
}

struct ExtendFn            // one candidate on a warp: both sides, the left half's gap shifting, the record (:728-953)
{
	Reads q, sub; const Cand* cands; const int32_t* ncand; int maxc, variant;
	ExtScratch* scratch; Overlap* out; int32_t* valid; int32_t* overflow;
	template <class L>
	ASM_HD void operator()(int64_t i, int slot, const L& lanes) const
	{
		const int r = (int)(i / maxc), ci = (int)(i % maxc);
		if (ci >= ncand[r]) { if (lanes.leader()) valid[i] = 0; return; }
		const Cand c = cands[i];
		ExtScratch S = scratch[slot];
		Strand st; st.fwd = q.text + q.start[r]; st.len = q.len[r]; st.rc = c.chain;
		Side Lf; Lf.text = sub.text; Lf.p1 = (int64_t)c.loc1 + SEED - 2; Lf.q = st; Lf.p2 = c.loc2 + SEED - 1; Lf.step = -1;
		const SideResult A = extend_side(lanes, Lf, c.num1, c.left1, c.left2, S, true);
		if (A.overflow) { if (lanes.leader()) { *overflow = 1; valid[i] = 0; } return; }
		Side Rt; Rt.text = sub.text; Rt.p1 = (int64_t)c.loc1 - 1; Rt.q = st; Rt.p2 = c.loc2; Rt.step = 1;
		const SideResult B = extend_side(lanes, Rt, c.num2, c.right1, c.right2, S, false);
		shift_gaps(lanes, Lf, A.done1 - 1, A.done2 - 1, S.l1, S.l2, A.cols);
		lanes.sync();
		const int loc = lanes.sum([&](int l) { int n = 0; for (int j = l; j < A.cols; j += 32) if (S.l1[j] != '-') n++; return n; });
		const int eit = lanes.sum([&](int l) { int n = 0; for (int j = l; j < A.cols; j += 32) if (S.l2[j] != '-') n++; return n; });
		const int mism_left = lanes.sum([&](int l) {
			int n = 0;
			for (int j = l; j < A.cols; j += 32) { const char a = S.l1[j], b = S.l2[j]; if (!(a == b && b != '-')) n++; }
			return n;
		});
		const int u_k = A.cols, s_k = B.cols;
		int left_loc1, left_loc, right_loc1, right_loc;
		if (u_k == SEED - 1) { left_loc1 = c.loc1 + SEED - loc - 1; left_loc = c.loc2 + SEED - eit; }
		else if (u_k > 0) { left_loc1 = c.loc1 + SEED - loc; left_loc = c.loc2 + SEED - eit + 1; }
		else { left_loc1 = c.loc1; left_loc = c.loc2 + 1; }
		if (s_k > 0) { right_loc1 = c.loc1 + B.done1 - 1; right_loc = c.loc2 + B.done2; }
		else { right_loc1 = c.loc1 + SEED - 1; right_loc = c.loc2 + SEED; }
		int n, mism;
		if (s_k >= SEED && u_k >= SEED) { n = u_k + s_k - SEED; mism = mism_left + B.gaps - B.gaps_head; }
		else if (u_k < SEED) { n = s_k; mism = B.gaps; }
		else { n = u_k; mism = mism_left; }
		left_loc1 -= c.readstart; right_loc1 -= c.readstart;
		if (!lanes.leader()) return;
		valid[i] = 0;
		if (!(right_loc1 - left_loc1 > 450)) return;
		float js;
		if (variant == 0) { js = (float)(2 * n - mism); js = js * 30 * 4 / n; }       // :942-943
		else { js = (float)mism; js = js / (4 * n); }                                  // mecat2trimpw.c:942-943
		Overlap o;
		o.sread = sub.first_id + c.readno; o.qread = q.first_id + r; o.score = js; o.sbeg = left_loc1 - 1; o.send = right_loc1;
		o.slen = sub.len[c.readno]; o.qlen = st.len;
		if (c.chain == 0) { o.strand = 0; o.qbeg = left_loc - 1; o.qend = right_loc; }
		else { o.strand = 1; o.qbeg = st.len - right_loc; o.qend = st.len - left_loc + 1; }
		out[i] = o; valid[i] = 1;
	}
};

// the printed overlaps of a read, counted and then gathered into the read's place of the result (read order, candidate order)
struct KeptCountFn
{
	const int32_t* valid; const int32_t* ncand; int maxc; int32_t* kept;
	ASM_HD void operator()(int64_t r) const
	{
		int n = 0;
		for (int i = 0; i < ncand[r]; ++i) n += valid[r * maxc + i] ? 1 : 0;
		kept[r] = n;
	}
};

struct GatherFn
{
	const Overlap* all; const int32_t* valid; const int32_t* ncand; int maxc; const int64_t* first; Overlap* out;
	ASM_HD void operator()(int64_t r) const
	{
		int64_t at = first[r];
		for (int i = 0; i < ncand[r]; ++i) if (valid[r * maxc + i]) out[at++] = all[r * maxc + i];
	}
};

}  // namespace mbasm
