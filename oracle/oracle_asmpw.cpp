// oracle/oracle_asmpw.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// CPU restatement of mecat2asmpw / mecat2trimpw: the overlapper mecat2canu runs on the
// corrected reads.  One C file in the reference, mecat2canu/src/mecat2asmpw/mecat2asmpw.c (1 166 lines); mecat2trimpw.c
// differs in the score gate (:640, 8 instead of 10) and the score it prints (:942-943); the *50.c files in MAXC (:23).
// Written from the program's behaviour: index of the subject file's text (creat_ref_index :397-497), per strand the
// block table and the candidate walk (pairwise_mapping :515-725), per candidate the chunked O(nd) alignment to both
// sides of the seed (align :108-197, :728-846), gap shifting of the left half (string_check :199-281) and the printed
// line (:848-953).  Pinned against the unmodified binaries: tests/golden/asm*.gz (tests/golden/make_golden.py asm).
//
// Where the reference reads memory it never wrote, this restatement reads zero -- what a fresh process gets from the
// allocator for the sizes involved:
//   * llocation[n], the end of the last read of the subject file (:691; load_read fills n entries of n + 1, :360-372);
//   * entries of a block beyond its score, and blocks no seed of this strand touched, when the neighbour votes walk a
//     block whose score passed SM = 60 (:699-712 read loczhi[j] / seedno[j] for j < score, i.e. into seedno[],
//     seednum, index and the next block; insert_loc is commented out at :605, so the score keeps counting).  The
//     reference would find there what earlier reads of the same thread left; its result then depends on the thread
//     schedule.  `history` = 1 reproduces one thread's memory instead, to pin the restatement against the binary where
//     the two conventions differ (tests/golden/asmdeep.*: 6 of 15 976 lines).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

namespace {

constexpr int ZV = 1000, DN = 500, BC = 10, SM = 60, SEED = 13;     // :18-22, seed_len :1080
constexpr double ERR = 0.10;                                         // ErrorRate :27

struct Block                 // Back_List :67-70 -- same layout: the overflow reads below depend on it
{
	short score, loczhi[SM], seedno[SM], seednum;
	int index;
};
static_assert(sizeof(Block) == 248, "Back_List layout");

struct Cand { int loc1, loc2, left1, left2, right1, right2, score, num1, num2, readno, readstart; char chain; };   // :61-64

struct Rec { int32_t sread, qread; float score; int32_t sbeg, send, slen, strand, qbeg, qend, qlen; };

int code_of(char c)          // atcttrans :298-304 ('t' is not accepted; the loaders upper-case)
{
	switch (c) { case 'A': case 'a': return 0; case 'T': return 1; case 'C': case 'c': return 2; case 'G': case 'g': return 3; }
	return 4;
}

struct Index
{
	std::vector<int> count, begin, pos;       // countin / databaseindex / allloc
	void build(const char* seq, int n)        // creat_ref_index :397-497: lists of more than 256 are dropped (sumvalue_x :307-314)
	{
		const int K = 1 << (2 * SEED);
		count.assign(K, 0);
		begin.assign(K, -1);
		auto walk = [&](auto&& f) {
			unsigned eit = 0; int run = 0;
			for (int i = 0; i < n; ++i) {
				const int t = code_of(seq[i]);
				if (seq[i] == 'N' || t == 4) { eit = 0; run = 0; continue; }
				eit = (eit << 2) + t;
				if (++run >= SEED) { f(eit, i + 2 - SEED); eit &= (1u << (2 * (SEED - 1))) - 1; }
			}
		};
		walk([&](unsigned c, int) { ++count[c]; });
		int sum = 0;
		for (int c = 0; c < K; ++c) {
			if (count[c] > 256) count[c] = 0;
			if (count[c] > 0) { begin[c] = sum; sum += count[c]; count[c] = 0; }
		}
		pos.assign(sum, 0);
		walk([&](unsigned c, int p) { if (begin[c] >= 0) pos[begin[c] + count[c]++] = p; });
	}
};

// |a / (b * len) - 1| < 0.10 in the two forms the reference uses (insert_loc, a third form, is never called: :605)
bool close_f1(int a, int b, float len) { return fabs(a / (b * len) - 1) < 0.10; }                         // find_location form
bool close_d(int a, int b) { return fabs(a / (b * BC * 1.0) - 1.0) < 0.10; }                              // neighbour votes, double quotient

// find_location :338-366
int find_location(const int* t_loc, const int* t_seedn, int* t_score, int* loc, int k, int* rep_loc, float len, int read_len1)
{
	int maxval = 0, maxi = 0, rep = 0, lasti = 0;
	for (int i = 0; i < k; ++i) t_score[i] = 0;
	for (int i = 0; i < k - 1; ++i) {
		int tempi = t_seedn[i];
		for (int j = i + 1; j < k; ++j)
			if (tempi != t_seedn[j] && t_seedn[j] - t_seedn[i] > 0 && t_loc[j] - t_loc[i] > 0 && t_loc[j] - t_loc[i] < read_len1 &&
			    close_f1(t_loc[j] - t_loc[i], t_seedn[j] - t_seedn[i], len)) {
				t_score[i]++; t_score[j]++; tempi = t_seedn[j];
			}
	}
	for (int i = 0; i < k; ++i) {
		if (maxval < t_score[i]) { maxval = t_score[i]; maxi = i; rep = 0; }
		else if (maxval == t_score[i]) { rep++; lasti = i; }
	}
	for (int i = 0; i < 4; ++i) loc[i] = 0;
	if (maxval >= 5 && rep == maxval) {
		loc[0] = t_loc[maxi]; loc[1] = t_seedn[maxi]; *rep_loc = maxi; loc[2] = t_loc[lasti]; loc[3] = t_seedn[lasti];
		return 1;
	}
	if (maxval >= 5) {
		auto take = [&](int j) {
			if (loc[0] == 0) { loc[0] = t_loc[j]; loc[1] = t_seedn[j]; *rep_loc = j; }
			else { loc[2] = t_loc[j]; loc[3] = t_seedn[j]; }
		};
		for (int j = 0; j < maxi; ++j)
			if (t_seedn[maxi] - t_seedn[j] > 0 && t_loc[maxi] - t_loc[j] > 0 && t_loc[maxi] - t_loc[j] < read_len1 &&
			    close_f1(t_loc[maxi] - t_loc[j], t_seedn[maxi] - t_seedn[j], len)) take(j);
		take(maxi);
		for (int j = maxi + 1; j < k; ++j)
			if (t_seedn[j] - t_seedn[maxi] > 0 && t_loc[j] - t_loc[maxi] > 0 && t_loc[j] - t_loc[maxi] <= read_len1 &&
			    close_f1(t_loc[j] - t_loc[maxi], t_seedn[j] - t_seedn[maxi], len)) take(j);
		return 1;
	}
	return 0;
}

struct Aln { int size, dist, qs, qe, ts, te; char q[2500], t[2500]; };
struct DP { int d, k, pre_k, x1, y1, x2, y2; };

// align :108-197: furthest-reaching O(nd) with a band that follows the best anti-diagonal; edits are gaps only
int align(const char* q, const char* t, int band_tol, Aln* A, int* V, int* U, std::vector<DP>& dp, std::vector<int>& row_start)
{
	const int q_len = (int)strlen(q), t_len = (int)strlen(t);
	const int max_d = (int)(ERR * (q_len + t_len));
	const int band = band_tol * 2, off = max_d;
	A->size = A->qs = A->qe = A->ts = A->te = 0;
	int best_m = -1, min_k = 0, max_k = 0, x = 0, y = 0, k = 0, d;
	bool aligned = false;
	dp.clear(); row_start.clear();
	for (d = 0; d < max_d; ++d) {
		if (max_k - min_k > band) break;
		row_start.push_back((int)dp.size());
		for (k = min_k; k <= max_k; k += 2) {
			int pre_k;
			if (k == min_k || (k != max_k && V[k - 1 + off] < V[k + 1 + off])) { pre_k = k + 1; x = V[k + 1 + off]; }
			else { pre_k = k - 1; x = V[k - 1 + off] + 1; }
			y = x - k;
			DP e; e.d = d; e.k = k; e.x1 = x; e.y1 = y;
			while (x < q_len && y < t_len && q[x] == t[y]) { ++x; ++y; }
			e.x2 = x; e.y2 = y; e.pre_k = pre_k;
			dp.push_back(e);
			V[k + off] = x; U[k + off] = x + y;
			if (x + y > best_m) best_m = x + y;
			if (x >= q_len || y >= t_len) { aligned = true; break; }
		}
		int new_min = max_k, new_max = min_k;
		for (int k2 = min_k; k2 <= max_k; k2 += 2)
			if (U[k2 + off] >= best_m - band_tol) { if (k2 < new_min) new_min = k2; if (k2 > new_max) new_max = k2; }
		max_k = new_max + 1; min_k = new_min - 1;
		if (aligned) {
			A->qe = x; A->te = y; A->dist = d;
			// walk back row by row, then forward writing the columns
			std::vector<int> px, py;
			int cd = d, ck = k;
			while (cd >= 0 && (int)px.size() < q_len + t_len + 1) {
				const int rs = row_start[cd];
				const DP& e = dp[rs + (ck - dp[rs].k) / 2];
				px.push_back(e.x2); py.push_back(e.y2);
				px.push_back(e.x1); py.push_back(e.y1);
				ck = e.pre_k; --cd;
			}
			int idx = (int)px.size() - 1, cx = px[idx], cy = py[idx], pos = 0;
			A->qs = cx; A->ts = cy;
			while (idx > 0) {
				--idx;
				const int nx = px[idx], ny = py[idx];
				if (cx == nx && cy == ny) continue;
				if (nx == cx) { for (int i = 0; i < ny - cy; ++i) { A->q[pos + i] = '-'; A->t[pos + i] = t[cy + i]; } pos += ny - cy; }
				else if (ny == cy) { for (int i = 0; i < nx - cx; ++i) { A->q[pos + i] = q[cx + i]; A->t[pos + i] = '-'; } pos += nx - cx; }
				else { for (int i = 0; i < nx - cx; ++i) A->q[pos + i] = q[cx + i]; for (int i = 0; i < ny - cy; ++i) A->t[pos + i] = t[cy + i]; pos += ny - cy; }
				cx = nx; cy = ny;
			}
			A->size = pos;
			break;
		}
	}
	return (A->qe == q_len || A->te == t_len) ? 1 : 0;
}

// string_check :199-281: walking the gapped strings from their end, a gap column whose pending letters agree pulls the
// run of agreeing letters over.  seq1/seq2 are the strings without gaps; a read one place before seq1 (:257) gives 0.
void string_check(const std::string& s1, const std::string& s2, char* str1, char* str2)
{
	const int len1 = (int)s1.size() - 1, len2 = (int)s2.size() - 1;
	auto c1 = [&](int i) -> char { return i < 0 ? 0 : s1[i]; };
	auto c2 = [&](int i) -> char { return i < 0 ? 0 : s2[i]; };
	auto pull = [&](int col, int k, int base1) {       // base1: index in seq1 of the first letter moved
		int s = 0, j = col;
		while (s < k) { if (str1[j] != '-') { str1[j] = '-'; s++; } j--; }
		s = 0; j = col;
		while (s < k) { if (str2[j] != '-') { str2[j] = '-'; s++; } j--; }
		return base1;
	};
	int loc1 = 0, loc2 = 0;
	for (int col = (int)strlen(str1) - 1; col > -1; --col) {
		if (str1[col] != '-') loc1++;
		else if (loc1 <= len1 && loc2 <= len2 && c1(len1 - loc1) == c2(len2 - loc2)) {
			int k = 1;
			while (loc1 + k <= len1 && loc2 + k <= len2 && c1(len1 - loc1 - k) == c2(len2 - loc2 - k)) k++;
			pull(col, k, 0);
			for (int s = 0, j = col; s < k; --j, ++s) { str1[j] = c1(len1 - loc1 - s); str2[j] = c2(len2 - loc2 - s); }
			if (str1[col] != '-') loc1++;
		}
		if (str2[col] != '-') loc2++;
		else if (str1[col] != '-' && (loc1 - 1 <= len1 && loc2 <= len2 && c1(len1 - loc1 + 1) == c2(len2 - loc2))) {
			int k = 1;
			while (loc1 + k - 1 <= len1 && loc2 + k <= len2 && c1(len1 - loc1 + 1 - k) == c2(len2 - loc2 - k)) k++;
			pull(col, k, 0);
			for (int s = 0, j = col; s < k; --j, ++s) { str1[j] = c1(len1 - loc1 + 1 - s); str2[j] = c2(len2 - loc2 - s); }
			if (str2[col] != '-') loc2++;
		}
		else if (str1[col] == '-' && (loc1 - 1 <= len1 && loc2 <= len2 && c1(len1 - loc1) == c2(len2 - loc2))) {
			int k = 1;
			while (loc1 + k <= len1 && loc2 + k <= len2 && c1(len1 - loc1 - k) == c2(len2 - loc2 - k)) k++;
			pull(col, k, 0);
			for (int s = 0, j = col; s < k; --j, ++s) { str1[j] = c1(len1 - loc1 - s); str2[j] = c2(len2 - loc2 - s); }
			if (str2[col] != '-') loc2++;
		}
	}
}

struct Mapper
{
	const char* seq; int seqcount; const int* ll; const int* slen; int nsub, sfirst;
	Index idx;
	std::vector<Block> db;
	int variant, maxc, history;
	std::vector<Rec> out;
	// scratch of align
	int V[2000], U[2000];
	std::vector<DP> dp; std::vector<int> rows;
	Aln A;

	short raw(int blk, int field) const       // field: index of a short inside the block array, as the overflow reads see it
	{
		return ((const short*)db.data())[(size_t)blk * (sizeof(Block) / 2) + field];
	}
	short loczhi_at(int blk, int j) const { return raw(blk, 1 + j); }
	short seedno_at(int blk, int j) const { return raw(blk, 1 + SM + j); }

	int read_of(int key) const     // binary :283-296: the last read that starts at or before key
	{
		int lo = 0, hi = nsub;
		while (lo < hi) { const int m = (lo + hi) / 2; if (ll[m] <= key) lo = m + 1; else hi = m; }
		return lo - 1;
	}

	// one side of the seed: chunks of DN letters while more than 600 are left (:735-789, :791-846)
	void extend_side(const char* p1, const char* p2, int step, int num, int len1, int len2, std::string& st1, std::string& st2)
	{
		char seq1[2500], seq2[2500];
		int done1 = 0, done2 = 0;
		bool more = true;
		st1.clear(); st2.clear();
		while (more) {
			int n;
			if (num > 600) n = DN; else { more = false; n = num; }
			for (int i = 0; i < n; ++i) { seq1[i] = p1[step * i]; seq2[i] = p2[step * i]; }
			if (n < 0) n = 0;
			seq1[n] = seq2[n] = 0;
			p1 += step * n; p2 += step * n;
			memset(V, 0, sizeof V); memset(U, 0, sizeof U);
			int ok = align(seq1, seq2, (int)(ERR * n), &A, V, U, dp, rows);
			if (ok) {
				A.q[A.size] = A.t[A.size] = 0;
				int k, loc = 0, sci = 0, run = 0;
				for (k = A.size - 1; k > -1 && run < 4; --k) {
					if (A.q[k] != '-') loc++;
					if (A.t[k] != '-') sci++;
					if (A.q[k] == A.t[k]) run++; else run = 0;
				}
				if (more) {
					loc = DN - A.qe + loc; sci = DN - A.te + sci;
					if (loc == DN) ok = 0;
					p1 -= step * loc; p2 -= step * sci;
					A.q[k + 1] = A.t[k + 1] = 0;
					done1 += DN - loc; done2 += DN - sci;
				} else {
					loc = num - A.qe; sci = num - A.te;
					if (loc == num) ok = 0;
					done1 += num - loc; done2 += num - sci;
				}
				if (ok) {
					st1 += A.q; st2 += A.t;
					num = len1 - done1 >= len2 - done2 ? len2 - done2 : len1 - done1;
				}
			}
			if (!ok) break;
		}
	}

	void map_read(const char* read, int read_len, int read_name)
	{
		std::string strand[2];
		strand[0].assign(read, read_len);
		strand[1].assign(read, read_len);
		for (int i = 0, j = read_len - 1; i < read_len; ++i, --j) {      // :575-586: reversed, upper-case ACGT complemented
			char c = read[j];
			switch (c) { case 'A': c = 'T'; break; case 'T': c = 'A'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; }
			strand[1][i] = c;
		}
		std::vector<Cand> cands;
		std::vector<int> index_list; std::vector<short> index_score;
		for (int ii = 0; ii < 2; ++ii) {
			const std::string& s = strand[ii];
			index_list.clear(); index_score.clear();
			const int cleave = (read_len - SEED) / BC + 1;                    // transnum_buchang :316-333
			for (int k = 0; k < cleave; ++k) {
				int code = 0;
				for (int j = 0; j < SEED; ++j) {
					const int t = (k * BC + j < read_len) ? code_of(s[k * BC + j]) : 4;
					if (t == 4) { code = -1; break; }
					code = (code << 2) + t;
				}
				if (code < 0 || idx.begin[code] < 0) continue;
				for (int i = 0; i < idx.count[code]; ++i) {                     // :594-627
					const int p = idx.pos[idx.begin[code] + i], blk = p / ZV, u = p % ZV;
					Block& b = db[blk];
					if (b.score == 0 || b.seednum < k + 1) {
						const int loc = ++b.score;
						if (loc <= SM) { b.loczhi[loc - 1] = (short)u; b.seedno[loc - 1] = (short)(k + 1); }
						const int s_k = blk > 0 ? b.score + db[blk - 1].score : b.score;
						if (b.index == -1) { b.index = (int)index_list.size(); index_list.push_back(blk); index_score.push_back((short)s_k); }
						else index_score[b.index] = (short)s_k;
					}
					b.seednum = (short)(k + 1);
				}
			}
			const int gate = variant ? 8 : 10;                                   // :640 (mecat2trimpw.c:640)
			for (size_t e = 0; e < index_list.size(); ++e) {
				if (!(index_score[e] > gate)) continue;
				const int blk = index_list[e];
				Block* b = &db[blk];
				if (b->score == 0) continue;
				int s_k = b->score, start_loc = blk * ZV, loc = 0;
				if (blk > 0) { loc = db[blk - 1].score; if (loc > 0) start_loc = (blk - 1) * ZV; }
				int t_list[150], t_seedn[150], t_score[150], u_k = 0, location[4], rep_loc = 0;
				if (loc == 0) for (int j = 0; j < s_k && j < SM; ++j) { t_list[u_k] = b->loczhi[j]; t_seedn[u_k] = b->seedno[j]; u_k++; }
				else {
					const Block* a = &db[blk - 1];
					for (int j = 0; j < loc && j < SM; ++j) { t_list[u_k] = a->loczhi[j]; t_seedn[u_k] = a->seedno[j]; u_k++; }
					for (int j = 0; j < s_k && j < SM; ++j) { t_list[u_k] = b->loczhi[j] + ZV; t_seedn[u_k] = b->seedno[j]; u_k++; }
				}
				if (!find_location(t_list, t_seedn, t_score, location, u_k, &rep_loc, BC, read_len)) continue;
				if (t_score[rep_loc] < 6) continue;
				Cand c;
				c.score = t_score[rep_loc];
				const int loc_seed = t_seedn[rep_loc];
				location[0] += start_loc;
				const int loc_list = location[0];
				const int readno = read_of(location[0]);
				const int readstart = ll[readno], readend = ll[readno + 1];
				if (sfirst + readno > read_name) continue;
				if (sfirst + readno == read_name) {                                 // :666-673: the read's own letters leave the table
					int u = readstart / ZV;
					Block* t = &db[u];
					int sk = readstart % ZV, k = 0;
					for (int j = 0; j < t->score && j < SM; ++j) if (t->loczhi[j] < sk) { t->loczhi[k] = t->loczhi[j]; k++; }
					t->score = (short)k;
					int kend = readend / ZV;
					for (++t, ++u; u < kend; ++u, ++t) t->score = 0;
					k = 0; sk = readend % ZV;
					for (int j = 0; j < t->score && j < SM; ++j) if (t->loczhi[j] > sk) { t->loczhi[k] = t->loczhi[j]; k++; }
					t->score = (short)k;
					continue;
				}
				c.readno = readno; c.readstart = readstart;
				location[1] = (location[1] - 1) * BC;
				const int left1 = location[0] - readstart + SEED - 1, right1 = readend - location[0];
				const int left2 = location[1] + SEED - 1, right2 = read_len - location[1];
				const int num1 = left1 >= left2 ? left2 : left1, num2 = right1 >= right2 ? right2 : right1;
				if (num1 + num2 < 400) continue;
				c.loc1 = location[0]; c.num1 = num1; c.loc2 = location[1]; c.num2 = num2;
				c.left1 = left1; c.left2 = left2; c.right1 = right1; c.right2 = right2;
				int seedcount = 0;
				{   // neighbour votes :699-712; the loops run to the block's score, not to SM
					int u = blk - 2;
					for (int k = num1 / ZV; u >= 0 && k >= 0; --u, --k) if (db[u].score > 0) {
						const int st = u * ZV; int hit = 0;
						for (int j = 0; j < db[u].score; ++j)
							if (close_d(loc_list - st - loczhi_at(u, j), loc_seed - seedno_at(u, j))) { seedcount++; hit++; }
						if (hit * 1.0 / db[u].score > 0.4) db[u].score = 0;
					}
					u = blk + 1;
					for (int k = num2 / ZV; k > 0; ++u, --k) if (db[u].score > 0) {
						const int st = u * ZV; int hit = 0;
						for (int j = 0; j < db[u].score; ++j)
							if (close_d(st + loczhi_at(u, j) - loc_list, seedno_at(u, j) - loc_seed)) { seedcount++; hit++; }
						if (hit * 1.0 / db[u].score > 0.4) db[u].score = 0;
					}
				}
				c.score += seedcount;
				c.chain = ii == 0 ? 'F' : 'R';
				// :716-726: kept in order of score, a newcomer behind its equals, the list cut at MAXC
				size_t at = cands.size();
				while (at > 0 && cands[at - 1].score < c.score) --at;
				if (at < (size_t)maxc) { cands.insert(cands.begin() + at, c); if ((int)cands.size() > maxc) cands.pop_back(); }
			}
			for (int blk : index_list) {                   // :727 resets score and index; `history` keeps the rest like the reference
				if (!history) memset(&db[blk], 0, sizeof(Block));
				db[blk].score = 0; db[blk].index = -1;
			}
		}
		for (const Cand& c : cands) emit(c, strand[c.chain == 'F' ? 0 : 1], read_len, read_name);
	}

	void emit(const Cand& c, const std::string& q, int read_len, int read_name)
	{
		std::string l1, l2, r1, r2;
		extend_side(seq + c.loc1 + SEED - 2, q.data() + c.loc2 + SEED - 1, -1, c.num1, c.left1, c.left2, l1, l2);
		extend_side(seq + c.loc1 - 1, q.data() + c.loc2, +1, c.num2, c.right1, c.right2, r1, r2);
		std::string g1, g2;
		for (size_t j = 0; j < l1.size(); ++j) { if (l1[j] != '-') g1 += l1[j]; if (l2[j] != '-') g2 += l2[j]; }
		std::vector<char> b1(l1.begin(), l1.end()), b2(l2.begin(), l2.end());
		b1.push_back(0); b2.push_back(0);
		string_check(g1, g2, b1.data(), b2.data());
		const int u_k = (int)l1.size();
		std::string o1, o2;
		int loc = 0, eit = 0;
		for (int j = u_k - 1; j > -1; --j) { o1 += b1[j]; if (b1[j] != '-') loc++; o2 += b2[j]; if (b2[j] != '-') eit++; }
		int left_loc1, left_loc, right_loc1, right_loc;
		if (u_k == SEED - 1) { left_loc1 = c.loc1 + SEED - loc - 1; left_loc = c.loc2 + SEED - eit; }
		else if (u_k > 0) { left_loc1 = c.loc1 + SEED - loc; left_loc = c.loc2 + SEED - eit + 1; }
		else { left_loc1 = c.loc1; left_loc = c.loc2 + 1; }
		const int s_k = (int)r1.size();
		loc = eit = 0;
		for (int k = 0; k < s_k; ++k) { if (r1[k] != '-') loc++; if (r2[k] != '-') eit++; }
		if (s_k > 0) { right_loc1 = c.loc1 + loc - 1; right_loc = c.loc2 + eit; }
		else { right_loc1 = c.loc1 + SEED - 1; right_loc = c.loc2 + SEED; }
		if (s_k >= SEED && u_k >= SEED) { o1 += r1.substr(SEED); o2 += r2.substr(SEED); }
		else if (u_k < SEED) { o1 = r1; o2 = r2; }
		left_loc1 -= c.readstart; right_loc1 -= c.readstart;
		if (!(right_loc1 - left_loc1 > 450)) return;
		int mism = 0;
		const int n = (int)o1.size();
		for (int j = 0; j < n; ++j) if (!(o1[j] == o2[j] && o2[j] != '-')) mism++;
		float js;
		if (variant == 0) { js = 2 * n - mism; js = js * 30 * 4 / (n); }          // :942-943
		else { js = mism; js = js / (4 * n); }                                     // mecat2trimpw.c:942-943
		Rec r;
		r.sread = sfirst + c.readno; r.qread = read_name; r.score = js; r.sbeg = left_loc1 - 1; r.send = right_loc1; r.slen = slen[c.readno];
		if (c.chain == 'F') { r.strand = 0; r.qbeg = left_loc - 1; r.qend = right_loc; }
		else { r.strand = 1; r.qbeg = read_len - right_loc; r.qend = read_len - left_loc + 1; }
		r.qlen = read_len;
		out.push_back(r);
	}
};

}  // namespace

// Subject file: text = reads joined with a NUL behind each (seqcount letters), read i at starts[i], numbered first_id + i.
// Query reads likewise.  variant 0 = mecat2asmpw, 1 = mecat2trimpw; maxc = MAXC (100, the *50 programs 50).
// history 0: every strand starts from zeroed blocks (the header's convention, what the product implements); 1: blocks
// keep what earlier reads of this call wrote, like one reference thread that maps all reads of the file in order (-T1,
// or any -T for files of at most PLL = 500 reads); 2: the same, but the memory is fresh at every chunk of PLL = 500 reads
// (:26, :556-567), what the binary does when every chunk is taken by a thread of its own (-T at least the number of
// chunks).  Modes 1 and 2 must reproduce the unmodified binary byte for byte.
// Records in read order, within a read in candidate order; *out is malloc'ed (10 x 4 bytes per record).
extern "C" int orc_asm_overlaps(const char* text, int seqcount, const int32_t* starts, const int32_t* lens, int n, int first_id,
                                const char* qtext, const int32_t* qstarts, const int32_t* qlens, int nq, int qfirst_id, int variant, int maxc,
                                int history, void** out, size_t* nout)
{
	Mapper* m = new Mapper;
	std::vector<int> ll(starts, starts + n);
	ll.push_back(0);
	m->seq = text; m->seqcount = seqcount; m->ll = ll.data(); m->slen = lens; m->nsub = n; m->sfirst = first_id;
	m->variant = variant; m->maxc = maxc; m->history = history;
	m->idx.build(text, seqcount);
	m->db.assign(seqcount / ZV + 5 + 256, Block());
	for (Block& b : m->db) { memset(&b, 0, sizeof b); b.index = -1; }
	for (int r = 0; r < nq; ++r) {
		if (history == 2 && r % 500 == 0)
			for (Block& b : m->db) { memset(&b, 0, sizeof b); b.index = -1; }
		m->map_read(qtext + qstarts[r], qlens[r], qfirst_id + r);
	}
	*nout = m->out.size();
	*out = malloc(m->out.size() * sizeof(Rec) + 1);
	memcpy(*out, m->out.data(), m->out.size() * sizeof(Rec));
	delete m;
	return 0;
}
