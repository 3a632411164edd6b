// oracle/oracle_pw.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the mecat2pw hot path: k-mer index, block seeding, DDF scoring,
// candidate selection, O(nd) diff extension and the per-read record assembly.
// Written from the behaviour documented in SURVEY.md Appendix A; each routine names the
// reference lines whose observable behaviour it reproduces.
#include "oracle.h"
#include "oracle_align.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr int KMER = 13;          // pw_impl.cpp:20
constexpr int STRIDE = 10;        // BC, pw_impl.h:12
constexpr int SLOTS = 40;         // SM, pw_impl.h:13
constexpr int SEGW = 2000;        // ZV, pw_impl.h:16
constexpr int MAX_OCC = 128;      // lookup_table.cpp:97
constexpr uint32_t NCODES = 1u << (2 * KMER);

inline int base_at(const uint8_t* pac, int64_t i)  // packed_db.h:103-107
{
	return (pac[i >> 2] >> (((~i) & 3) << 1)) & 3;
}

// ------------------------------------------------------------------ A1 index
struct Index
{
	std::vector<uint32_t> begin;   // NCODES + 1, CSR over kept k-mers
	std::vector<int32_t> pos;
};

// ------------------------------------------------------------------ bucket state (Back_List, pw_impl.h:29-33)
struct Bucket
{
	int16_t score = 0;
	int16_t loc[SLOTS];
	int16_t seed[SLOTS];
	int16_t last_seed = 0;
	int32_t order = -1;
};

inline bool ddf_close_f32(int dloc, int dseed)  // float32 DDF test of pw_impl.cpp:135,165
{
	float ratio = dloc / (dseed * (float)STRIDE);
	return std::fabs((double)ratio - 1.0) < 0.25;
}

// pw_impl.cpp:121-159
void insert_loc_impl(Bucket& b, int loc, int seedn)
{
	int l[SLOTS + 1], s[SLOTS + 1], votes[SLOTS + 1];
	for (int i = 0; i < SLOTS; ++i) { l[i] = b.loc[i]; s[i] = b.seed[i]; votes[i] = 0; }
	l[SLOTS] = loc; s[SLOTS] = seedn; votes[SLOTS] = 0;
	for (int i = 0; i < SLOTS; ++i)
		for (int j = i + 1; j <= SLOTS; ++j)
			if (s[j] - s[i] > 0 && l[j] - l[i] > 0 && ddf_close_f32(l[j] - l[i], s[j] - s[i])) { ++votes[i]; ++votes[j]; }
	int worst = -1, worst_votes = 10000;
	for (int i = 0; i <= SLOTS; ++i)
		if (votes[i] < worst_votes) { worst_votes = votes[i]; worst = i; }
	if (worst_votes == SLOTS) {
		b.loc[SLOTS - 1] = (int16_t)loc;
		b.seed[SLOTS - 1] = (int16_t)seedn;
	} else if (worst_votes < SLOTS && worst < SLOTS) {
		for (int i = worst; i < SLOTS; ++i) { b.loc[i] = (int16_t)l[i + 1]; b.seed[i] = (int16_t)s[i + 1]; }
		--b.score;
	}
}

// pw_impl.cpp:161-239
int find_location_impl(const int* t_loc, const int* t_seedn, int* t_score, int* loc, int k, int* rep_loc, int read_len)
{
	for (int i = 0; i < k; ++i) t_score[i] = 0;
	for (int i = 0; i + 1 < k; ++i) {
		int last_seed = t_seedn[i];
		for (int j = i + 1; j < k; ++j) {
			int ds = t_seedn[j] - t_seedn[i], dl = t_loc[j] - t_loc[i];
			if (last_seed != t_seedn[j] && ds > 0 && dl > 0 && dl < read_len && ddf_close_f32(dl, ds)) {
				++t_score[i]; ++t_score[j];
				last_seed = t_seedn[j];
			}
		}
	}
	int best = 0, besti = 0, ties = 0, last_tie = 0;
	for (int i = 0; i < k; ++i) {
		if (best < t_score[i]) { best = t_score[i]; besti = i; ties = 0; }
		else if (best == t_score[i]) { ++ties; last_tie = i; }
	}
	loc[0] = loc[1] = loc[2] = loc[3] = 0;
	if (best < 5) return 0;
	if (ties == best) {
		loc[0] = t_loc[besti]; loc[1] = t_seedn[besti]; *rep_loc = besti;
		loc[2] = t_loc[last_tie]; loc[3] = t_seedn[last_tie];
		return 1;
	}
	auto take = [&](int j) {
		if (loc[0] == 0) { loc[0] = t_loc[j]; loc[1] = t_seedn[j]; *rep_loc = j; }
		else { loc[2] = t_loc[j]; loc[3] = t_seedn[j]; }
	};
	for (int j = 0; j < besti; ++j) {
		int ds = t_seedn[besti] - t_seedn[j], dl = t_loc[besti] - t_loc[j];
		if (ds > 0 && dl > 0 && dl < read_len && ddf_close_f32(dl, ds)) take(j);
	}
	take(besti);
	for (int j = besti + 1; j < k; ++j) {
		int ds = t_seedn[j] - t_seedn[besti], dl = t_loc[j] - t_loc[besti];
		if (ds > 0 && dl > 0 && dl <= read_len && ddf_close_f32(dl, ds)) take(j);
	}
	return 1;
}

// split_database.cpp:16-35
int read_of_offset(const orc_volume* v, int offset)
{
	const int32_t* a = v->offset_size;
	int n = v->num_reads;
	int left = 0, right = n - 1, mid = (left + right) / 2;
	if (a[2 * right] < offset) return right;
	while (left <= right) {
		int o = a[2 * mid], sz = a[2 * mid + 1];
		if (o <= offset && o + sz > offset) return mid;
		if (o + sz <= offset) left = mid + 1;
		else right = mid - 1;
		mid = (left + right) / 2;
	}
	return mid;
}

struct Seeder
{
	std::vector<Bucket> db;          // SeedingBK::database, pw_impl.cpp:99-111
	std::vector<int32_t> touched;    // index_list
	std::vector<int16_t> snap;       // index_score
	explicit Seeder(int ref_bases) : db(ref_bases / SEGW + 5) {}
};

void unpack_read(const orc_volume* v, int rid, int strand, std::vector<char>& out)
{
	int off = v->offset_size[2 * rid], sz = v->offset_size[2 * rid + 1];
	out.resize(sz);
	if (!strand) for (int i = 0; i < sz; ++i) out[i] = (char)base_at(v->pac, off + i);
	else for (int i = 0; i < sz; ++i) out[sz - 1 - i] = (char)(3 - base_at(v->pac, off + i));  // pw_impl.cpp:69-81
}

// pw_impl.cpp:241-286 (+ extract_kmers :83-97)
int seed_strand(const Index& idx, const char* read, int len, Seeder& S)
{
	S.touched.clear(); S.snap.clear();
	int nk = (len - KMER) / STRIDE + 1;
	for (int km = 0; km < nk; ++km) {
		uint32_t code = 0;
		for (int j = 0; j < KMER; ++j) code = (code << 2) | (uint32_t)read[km * STRIDE + j];
		for (uint32_t h = idx.begin[code]; h < idx.begin[code + 1]; ++h) {
			int p = idx.pos[h], seg = p / SEGW, off = p % SEGW;
			Bucket& b = S.db[seg];
			if (b.score == 0 || b.last_seed < km + 1) {
				int n = ++b.score;
				if (n <= SLOTS) { b.loc[n - 1] = (int16_t)off; b.seed[n - 1] = (int16_t)(km + 1); }
				else insert_loc_impl(b, off, km + 1);
				int s = b.score + (seg > 0 ? S.db[seg - 1].score : 0);
				if (b.order < 0) {
					b.order = (int)S.touched.size();
					S.touched.push_back(seg);
					S.snap.push_back((int16_t)s);
				} else
					S.snap[b.order] = (int16_t)s;
			}
			b.last_seed = (int16_t)(km + 1);
		}
	}
	return (int)S.touched.size();
}

struct Cand
{
	int loc1, loc2, left1, left2, right1, right2, score, num1, num2, readno, readstart, chain;
};

inline bool ddf_close_f64(int dloc, int dseed)   // pw_impl.cpp:412,429
{
	double r = dloc / (dseed * STRIDE * 1.0);
	return std::fabs(r - 1.0) < 0.25;
}

// pw_impl.cpp:288-465
int collect_candidates(const orc_volume* ref, Seeder& S, int qid, int qlen, int chain, const orc_pw_params* P,
                       std::vector<Cand>& list, int have)
{
	const int maxc = P->num_candidates;
	const int kmin = P->min_kmer_match;
	const int min_span = (P->tech == 0) ? 1800 : 400;   // pw_impl.cpp:845,848
	int w_loc[2 * SLOTS + 10], w_seed[2 * SLOTS + 10], w_score[2 * SLOTS + 10];
	for (size_t t = 0; t < S.touched.size(); ++t) {
		if (S.snap[t] < 2 * kmin) continue;
		const int seg = S.touched[t];
		Bucket* b = &S.db[seg];
		if (b->score == 0) continue;
		int cur = b->score, prev = 0, origin = seg * SEGW;
		if (seg > 0) { prev = S.db[seg - 1].score; if (prev > 0) origin = (seg - 1) * SEGW; }
		int n = 0;
		if (prev) {
			const Bucket& pb = S.db[seg - 1];
			for (int j = 0; j < prev && j < SLOTS; ++j) { w_loc[n] = pb.loc[j]; w_seed[n] = pb.seed[j]; ++n; }
			for (int j = 0; j < cur && j < SLOTS; ++j) { w_loc[n] = b->loc[j] + SEGW; w_seed[n] = b->seed[j]; ++n; }
		} else
			for (int j = 0; j < cur && j < SLOTS; ++j) { w_loc[n] = b->loc[j]; w_seed[n] = b->seed[j]; ++n; }
		int anchor[4], rep = 0;
		if (!find_location_impl(w_loc, w_seed, w_score, anchor, n, &rep, qlen)) continue;
		if (w_score[rep] < 2 * kmin + 2) continue;
		Cand c;
		c.score = w_score[rep];
		c.chain = chain;
		const int anchor_seed = w_seed[rep];
		const int gpos = origin + anchor[0];
		int sidx = read_of_offset(ref, gpos);
		const int sstart = ref->offset_size[2 * sidx], ssize = ref->offset_size[2 * sidx + 1];
		const int send = sstart + ssize + 1;
		const int sid = sidx + ref->start_read_id;
		if (sid > qid) continue;
		if (sid == qid) {
			// purge the read's own span (:371-383); only the position column is compacted
			int u = sstart / SEGW, cut = sstart % SEGW, k = 0;
			Bucket* p = &S.db[u];
			for (int j = 0; j < p->score && j < SLOTS; ++j) if (p->loc[j] < cut) p->loc[k++] = p->loc[j];
			p->score = (int16_t)k;
			++p; ++u;
			for (int last = send / SEGW; u < last; ++u, ++p) p->score = 0;
			cut = send % SEGW; k = 0;
			for (int j = 0; j < p->score && j < SLOTS; ++j) if (p->loc[j] > cut) p->loc[k++] = p->loc[j];
			p->score = (int16_t)k;
			continue;
		}
		c.readno = sid; c.readstart = sstart;
		const int qoff = (anchor[1] - 1) * STRIDE;
		c.left1 = gpos - sstart + KMER - 1; c.right1 = send - gpos;
		c.left2 = qoff + KMER - 1; c.right2 = qlen - qoff;
		c.num1 = std::min(c.left1, c.left2); c.num2 = std::min(c.right1, c.right2);
		if (c.num1 + c.num2 < min_span) continue;
		c.loc1 = gpos - sstart; c.loc2 = qoff;
		int extra = 0;
		int nl = (c.num1 + SEGW - 1) / SEGW;
		for (int u = seg - 1; u >= 0 && nl > 0; --u, --nl) {
			Bucket& nb = S.db[u];
			if (nb.score <= 0) continue;
			int cnt = std::min<int>(nb.score, SLOTS), ok = 0;
			for (int j = 0; j < cnt; ++j) if (ddf_close_f64(gpos - u * SEGW - nb.loc[j], anchor_seed - nb.seed[j])) ++ok;
			extra += ok;
			if (ok * 1.0 / cnt > 0.4) nb.score = 0;
		}
		int nr = (c.num2 + SEGW - 1) / SEGW;
		for (int u = seg + 1; nr; ++u, --nr) {
			if ((size_t)u >= S.db.size()) continue;   // reference reads its +5 slack, always empty there
			Bucket& nb = S.db[u];
			if (nb.score <= 0) continue;
			int cnt = std::min<int>(nb.score, SLOTS), ok = 0;
			for (int j = 0; j < cnt; ++j) if (ddf_close_f64(u * SEGW + nb.loc[j] - gpos, nb.seed[j] - anchor_seed)) ++ok;
			extra += ok;
			if (ok * 1.0 / cnt > 0.4) nb.score = 0;
		}
		c.score += extra;
		// stable insertion into the score-descending list capped at maxc (:442-455)
		int at = have;
		while (at > 0 && list[at - 1].score < c.score) --at;
		if (at < maxc) {
			int last = std::min(have, maxc - 1);
			for (int u = last; u > at; --u) list[u] = list[u - 1];
			list[at] = c;
		}
		if (have < maxc) ++have;
	}
	for (int seg : S.touched) { S.db[seg].score = 0; S.db[seg].order = -1; }
	return have;
}

int candidates_of_read(const Index& idx, const orc_volume* ref, const orc_volume* reads, int rid,
                       const orc_pw_params* P, Seeder& S, std::vector<Cand>& list,
                       std::vector<char>& fwd, std::vector<char>& rev)
{
	if ((int)list.size() < P->num_candidates + 1) list.resize(P->num_candidates + 1);
	int len = reads->offset_size[2 * rid + 1];
	unpack_read(reads, rid, 0, fwd);
	unpack_read(reads, rid, 1, rev);
	int n = 0;
	for (int s = 0; s < 2; ++s) {
		seed_strand(idx, s ? rev.data() : fwd.data(), len, S);
		n = collect_candidates(ref, S, rid + reads->start_read_id, len, s, P, list, n);
	}
	return n;
}

// ------------------------------------------------------------------ A8-A11 O(nd) extension
using orc::BlockAln;
using orc::DiffScratch;

// Align, diff_gapalign.cpp:107-219: budget 0.3 * (q + t), best-cell fall-back
void align_block(const char* q, int qlen, const char* t, int tlen, int tol, int fwd, DiffScratch& W, BlockAln& A)
{
	orc::align_block(q, qlen, t, tlen, tol, (int)(.3 * (qlen + tlen)), true, fwd, W, A);
}

// trim_mismatch_end, gapalign.cpp:48-67
bool trim_tail(const BlockAln& A, int need, int& qcnt, int& tcnt, int& acnt)
{
	int m = 0, k;
	qcnt = tcnt = acnt = 0;
	for (k = A.size - 1; k >= 0 && m < need; --k) {
		++acnt;
		if (A.q[k] != 4) ++qcnt;
		if (A.t[k] != 4) ++tcnt;
		m = (A.q[k] == A.t[k]) ? m + 1 : 0;
	}
	return m == need && k > 0;
}

// dw_in_one_direction, diff_gapalign.cpp:221-292 (+ retrieve_next_aln_block, gapalign.cpp:10-45)
void extend_one_way(const char* q, int qsize, const char* t, int tsize, int fwd, DiffScratch& W,
                    std::vector<char>& oq, std::vector<char>& ot)
{
	int qi = 0, ti = 0;
	BlockAln A;
	for (;;) {
		int qleft = qsize - qi, tleft = tsize - ti, qblk, tblk;
		bool last;
		if (qleft < 600 || tleft < 600) {
			qblk = std::min(qleft, static_cast<int>(tleft + tleft * 0.2));
			tblk = std::min(tleft, static_cast<int>(qleft + qleft * 0.2));
			last = true;
		} else { qblk = tblk = 500; last = false; }
		const char* Q = fwd ? q + qi : q - qi;
		const char* T = fwd ? t + ti : t - ti;
		align_block(Q, qblk, T, tblk, (int)(0.3 * std::max(qblk, tblk)), fwd, W, A);
		int qc, tc, ac;
		if (!trim_tail(A, 4, qc, tc, ac)) break;
		bool full = (qblk - A.q_e <= 20) || (tblk - A.t_e <= 20);
		bool stop = last || !full;
		if (stop) { qc -= 4; tc -= 4; ac -= 4; }
		int keep = A.size - ac;
		oq.insert(oq.end(), A.q.begin(), A.q.begin() + keep);
		ot.insert(ot.end(), A.t.begin(), A.t.begin() + keep);
		if (stop) break;
		qi += A.q_e - qc;
		ti += A.t_e - tc;
	}
}

struct GoResult
{
	bool ok = false;
	int qs = 0, qe = 0, ts = 0, te = 0, size = 0, matches = 0;
	double ident = 0;
	std::string qstr, tstr;
};

// DiffAligner::go, diff_gapalign.cpp:295-349 ; calc_ident diff_gapalign.h:90-97
void diff_go(const char* q, int qstart, int qsize, const char* t, int tstart, int tsize, int min_aln,
             DiffScratch& W, GoResult& R, bool want_str)
{
	std::vector<char> lq, lt, rq, rt;
	extend_one_way(q + qstart - 1, qstart, t + tstart - 1, tstart, 0, W, lq, lt);
	extend_one_way(q + qstart, qsize - qstart, t + tstart, tsize - tstart, 1, W, rq, rt);
	int li = 0, lj = 0, ri = 0, rj = 0, m = 0;
	for (size_t i = 0; i < lq.size(); ++i) { li += lq[i] != 4; lj += lt[i] != 4; m += lq[i] == lt[i]; }
	for (size_t i = 0; i < rq.size(); ++i) { ri += rq[i] != 4; rj += rt[i] != 4; m += rq[i] == rt[i]; }
	R.qs = qstart - li; R.ts = tstart - lj; R.qe = qstart + ri; R.te = tstart + rj;
	R.size = (int)(lq.size() + rq.size());
	R.matches = m;
	R.ident = R.size ? 100.0 * m / R.size : 0.0;
	R.ok = R.size >= min_aln;
	if (want_str) {
		R.qstr.clear(); R.tstr.clear();
		for (size_t i = lq.size(); i-- > 0;) { R.qstr.push_back("ACGT-"[(int)lq[i]]); R.tstr.push_back("ACGT-"[(int)lt[i]]); }
		for (size_t i = 0; i < rq.size(); ++i) { R.qstr.push_back("ACGT-"[(int)rq[i]]); R.tstr.push_back("ACGT-"[(int)rt[i]]); }
	}
}

// ------------------------------------------------------------------ records
struct EC { int32_t qdir, qid, qext, qsize, qoff, qend, sdir, sid, sext, ssize, soff, send, score; };  // alignment.h:8-13
struct M4                                                                                            // alignment.h:21-37
{
	int64_t qid, sid; double ident; int32_t vscore, qdir; int64_t qoff, qend, qsize; int32_t sdir, pad;
	int64_t soff, send, ssize, qext, sext;
};
static_assert(sizeof(EC) == 52, "ExtensionCandidate layout");
static_assert(sizeof(M4) == 104, "M4Record layout");

struct M4Order   // CmpM4RecordByQidAndOvlpSize, pw_impl.cpp:539-548
{
	bool operator()(const M4& a, const M4& b) const
	{
		if (a.qid != b.qid) return a.qid < b.qid;
		int64_t oa = std::min(a.qend - a.qoff, a.send - a.soff), ob = std::min(b.qend - b.qoff, b.send - b.soff);
		return oa > ob;
	}
};

// append_m4v + check_records_containment, pw_impl.cpp:550-610
void filter_contained(std::vector<M4>& v, std::vector<M4>& out)
{
	std::sort(v.begin(), v.end(), M4Order());
	std::vector<int> ok(v.size(), 1);
	for (size_t i = 0; i < v.size();) {
		size_t j = i + 1;
		while (j < v.size() && v[j].qid == v[i].qid) ++j;
		for (size_t a = i; a < j; ++a) {
			if (!ok[a]) continue;
			for (size_t b = a + 1; b < j; ++b) {
				if (!ok[b] || v[a].sdir != v[b].sdir) continue;
				if (v[b].qoff + 100 >= v[a].qoff && v[b].qend - 100 <= v[a].qend &&
				    v[b].soff + 100 >= v[a].soff && v[b].send - 100 <= v[a].send) ok[b] = 0;
			}
		}
		i = j;
	}
	for (size_t i = 0; i < v.size(); ++i) if (ok[i]) out.push_back(v[i]);
	v.clear();
}

}  // namespace

extern "C" {

int orc_pack_reads(const char* const* seqs, const int32_t* lens, int n, int32_t** offset_size, uint8_t** pac,
                   int32_t* num_bases)
{
	int64_t total = 0;
	for (int i = 0; i < n; ++i) total += lens[i] + 1;
	int32_t* os = (int32_t*)malloc(sizeof(int32_t) * 2 * (n ? n : 1));
	uint8_t* p = (uint8_t*)calloc((total + 3) / 4 + 16, 1);
	int64_t cur = 0;
	for (int i = 0; i < n; ++i) {
		os[2 * i] = (int32_t)cur; os[2 * i + 1] = lens[i];
		for (int j = 0; j < lens[i]; ++j, ++cur) {
			int c;
			switch (seqs[i][j]) {   // defs.cpp:3-42 (only ACGT are meaningful for volumes)
			case 'A': case 'a': c = 0; break;
			case 'C': case 'c': c = 1; break;
			case 'G': case 'g': c = 2; break;
			case 'T': case 't': c = 3; break;
			default: c = 0; break;
			}
			p[cur >> 2] |= (uint8_t)(c << (((~cur) & 3) << 1));
		}
		++cur;   // one pad base per read, split_database.cpp:251
	}
	*offset_size = os; *pac = p; *num_bases = (int32_t)cur;
	return 0;
}

// create_ref_index, lookup_table.cpp:64-160
void* orc_index_build(const orc_volume* v)
{
	Index* I = new Index;
	std::vector<int32_t> cnt(NCODES, 0);
	const uint32_t mask = NCODES - 1;
	for (int r = 0; r < v->num_reads; ++r) {
		int off = v->offset_size[2 * r], sz = v->offset_size[2 * r + 1];
		uint32_t code = 0;
		for (int j = 0; j < sz; ++j) {
			code = (code << 2) | (uint32_t)base_at(v->pac, off + j);
			if (j >= KMER - 1) { code &= mask; ++cnt[code]; }
		}
	}
	I->begin.resize((size_t)NCODES + 1);
	uint32_t run = 0;
	for (uint32_t c = 0; c < NCODES; ++c) {
		I->begin[c] = run;
		if (cnt[c] > MAX_OCC) cnt[c] = 0;
		run += (uint32_t)cnt[c];
	}
	I->begin[NCODES] = run;
	I->pos.resize(run);
	std::vector<uint32_t> cursor(I->begin.begin(), I->begin.end() - 1);
	for (int r = 0; r < v->num_reads; ++r) {
		int off = v->offset_size[2 * r], sz = v->offset_size[2 * r + 1];
		uint32_t code = 0;
		for (int j = 0; j < sz; ++j) {
			code = (code << 2) | (uint32_t)base_at(v->pac, off + j);
			if (j >= KMER - 1) {
				code &= mask;
				if (cnt[code]) I->pos[cursor[code]++] = off + j + 1 - KMER;
			}
		}
	}
	return I;
}

void orc_index_free(void* idx) { delete (Index*)idx; }

int orc_index_lookup(const void* idx, uint32_t code, const int32_t** list)
{
	const Index* I = (const Index*)idx;
	if (list) *list = I->pos.data() + I->begin[code];
	return (int)(I->begin[code + 1] - I->begin[code]);
}

int64_t orc_index_num_kmers(const void* idx) { return (int64_t)((const Index*)idx)->pos.size(); }

int orc_seeding(const void* idx, const orc_volume* ref, const orc_volume* reads, int rid, int strand,
                int32_t* seg, int16_t* index_score, int16_t* rows, int cap)
{
	Seeder S(ref->num_bases);
	std::vector<char> rd;
	unpack_read(reads, rid, strand, rd);
	int n = seed_strand(*(const Index*)idx, rd.data(), (int)rd.size(), S);
	for (int i = 0; i < n && i < cap; ++i) {
		const Bucket& b = S.db[S.touched[i]];
		seg[i] = S.touched[i];
		index_score[i] = S.snap[i];
		int16_t* r = rows + 82 * i;
		r[0] = b.score;
		memcpy(r + 1, b.loc, sizeof b.loc);
		memcpy(r + 41, b.seed, sizeof b.seed);
		r[81] = 0;
	}
	return n;
}

void orc_insert_loc(int16_t* score, int16_t* loczhi, int16_t* seedno, int loc, int seedn)
{
	Bucket b;
	b.score = *score;
	memcpy(b.loc, loczhi, sizeof b.loc);
	memcpy(b.seed, seedno, sizeof b.seed);
	insert_loc_impl(b, loc, seedn);
	*score = b.score;
	memcpy(loczhi, b.loc, sizeof b.loc);
	memcpy(seedno, b.seed, sizeof b.seed);
}

int orc_find_location(const int* t_loc, const int* t_seedn, int* t_score, int* loc, int k, int* rep_loc, int read_len)
{
	return find_location_impl(t_loc, t_seedn, t_score, loc, k, rep_loc, read_len);
}

int orc_pw_candidates(const void* idx, const orc_volume* ref, const orc_volume* reads, int rid,
                      const orc_pw_params* p, int32_t* out)
{
	Seeder S(ref->num_bases);
	std::vector<Cand> list;
	std::vector<char> f, r;
	int n = candidates_of_read(*(const Index*)idx, ref, reads, rid, p, S, list, f, r);
	for (int i = 0; i < n; ++i) {
		const Cand& c = list[i];
		int32_t* o = out + 12 * i;
		o[0] = c.loc1; o[1] = c.loc2; o[2] = c.left1; o[3] = c.left2; o[4] = c.right1; o[5] = c.right2;
		o[6] = c.score; o[7] = c.num1; o[8] = c.num2; o[9] = c.readno; o[10] = c.readstart; o[11] = c.chain;
	}
	return n;
}

int orc_diff_go(const char* q, int qstart, int qsize, const char* t, int tstart, int tsize, int min_aln,
                int32_t* out, double* ident, char* qstr, char* tstr, int cap)
{
	DiffScratch W;
	GoResult R;
	diff_go(q, qstart, qsize, t, tstart, tsize, min_aln, W, R, qstr != NULL);
	out[0] = R.ok; out[1] = R.qs; out[2] = R.qe; out[3] = R.ts; out[4] = R.te; out[5] = R.size; out[6] = R.matches;
	if (ident) *ident = R.ident;
	if (qstr && (int)R.qstr.size() < cap) { memcpy(qstr, R.qstr.c_str(), R.qstr.size() + 1); memcpy(tstr, R.tstr.c_str(), R.tstr.size() + 1); }
	return R.ok;
}

void orc_diff_align_block(const char* q, int qlen, const char* t, int tlen, int right_extend, int32_t* out)
{
	DiffScratch W;
	BlockAln A;
	align_block(q, qlen, t, tlen, (int)(0.3 * std::max(qlen, tlen)), right_extend, W, A);
	int qc, tc, ac;
	bool ok = trim_tail(A, 4, qc, tc, ac);
	out[0] = A.q_e; out[1] = A.t_e; out[2] = A.dist; out[3] = A.size; out[4] = ok; out[5] = qc; out[6] = tc; out[7] = ac;
}

// process_one_volume for one (index volume, query volume) pair: pw_impl.cpp:623-818, 835-882
int orc_pw_tile(const orc_volume* ref, const orc_volume* reads, const orc_pw_params* P, int threads,
                void** records, size_t* n)
{
	Index* I = (Index*)orc_index_build(ref);
	const int N = reads->num_reads;
	std::vector<std::vector<EC>> ec_out(P->task == 0 ? N : 0);
	std::vector<std::vector<M4>> m4_out(P->task == 1 ? N : 0);
	if (threads < 1) threads = 1;
	std::vector<std::thread> pool;
	std::mutex mu;
	int next = 0;
	for (int t = 0; t < threads; ++t)
		pool.emplace_back([&]() {
			Seeder S(ref->num_bases);
			std::vector<Cand> list;
			std::vector<char> fwd, rev, subj;
			DiffScratch W;
			GoResult R;
			std::vector<M4> local;
			for (;;) {
				int lo, hi;
				{ std::lock_guard<std::mutex> g(mu); lo = next; next += 64; }
				if (lo >= N) break;
				hi = std::min(N, lo + 64);
				for (int rid = lo; rid < hi; ++rid) {
					int nc = candidates_of_read(*I, ref, reads, rid, P, S, list, fwd, rev);
					const int qsize = reads->offset_size[2 * rid + 1], qid = rid + reads->start_read_id;
					for (int c = 0; c < nc; ++c) {
						const Cand& cd = list[c];
						int sstart = cd.loc1, qstart = cd.loc2;
						if (qstart && sstart) { qstart += KMER / 2; sstart += KMER / 2; }   // pw_impl.cpp:681-685,771-775
						const int sidx = cd.readno - ref->start_read_id;
						const int ssize = ref->offset_size[2 * sidx + 1];
						if (P->task == 0) {
							EC e; memset(&e, 0, sizeof e);
							e.qid = qid; e.qdir = cd.chain; e.qext = qstart; e.sid = cd.readno; e.sdir = 0; e.sext = sstart;
							e.score = cd.score; e.qsize = qsize; e.ssize = ssize;
							if (e.qdir == 1) e.qext = e.qsize - 1 - e.qext;   // pw_impl.cpp:791
							ec_out[rid].push_back(e);
						} else {
							unpack_read(ref, sidx, 0, subj);
							const char* q = cd.chain ? rev.data() : fwd.data();
							if (P->tech == 1) {       // XdropAligner for nanopore reads, pw_impl.cpp:638-642
								int32_t o[8]; double id = 0;
								R.ok = orc_xdrop_go(q, qstart, qsize, subj.data(), sstart, ssize, P->min_align_size, o, &id, NULL, NULL, 0) != 0;
								R.qs = o[1]; R.qe = o[2]; R.ts = o[3]; R.te = o[4]; R.size = o[5]; R.matches = o[6]; R.ident = id;
							} else
								diff_go(q, qstart, qsize, subj.data(), sstart, ssize, P->min_align_size, W, R, false);
							if (!R.ok) continue;
							M4 m; memset(&m, 0, sizeof m);   // fill_m4record, pw_impl.cpp:467-506
							m.qid = cd.readno; m.sid = qid; m.ident = R.ident; m.vscore = cd.score; m.qdir = 0;
							m.qoff = R.ts; m.qend = R.te; m.qsize = ssize; m.ssize = qsize; m.qext = sstart;
							if (!cd.chain) { m.sdir = 0; m.soff = R.qs; m.send = R.qe; m.sext = qstart; }
							else { m.sdir = 1; m.soff = qsize - R.qe; m.send = qsize - R.qs; m.sext = qsize - 1 - qstart; }
							local.push_back(m);
						}
					}
					if (P->task == 1) filter_contained(local, m4_out[rid]);
				}
			}
		});
	for (auto& th : pool) th.join();
	delete I;
	size_t total = 0;
	if (P->task == 0) {
		for (auto& v : ec_out) total += v.size();
		EC* out = (EC*)malloc(sizeof(EC) * (total ? total : 1));
		size_t k = 0;
		for (auto& v : ec_out) for (auto& e : v) out[k++] = e;
		*records = out;
	} else {
		for (auto& v : m4_out) total += v.size();
		M4* out = (M4*)malloc(sizeof(M4) * (total ? total : 1));
		size_t k = 0;
		for (auto& v : m4_out) for (auto& e : v) out[k++] = e;
		*records = out;
	}
	*n = total;
	return 0;
}

void orc_free(void* p) { free(p); }

}  // extern "C"
