// oracle/oracle_ref.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of mecat2ref (reads against a reference), the next row of the scope table.
// No CUDA path exists for its seeding yet; this file and the golden output of the UNMODIFIED binary
// (tests/golden/refmap.*) are the checker that path will be built against.  Follows
//   creat_ref_index, reference_mapping, insert_loc, transnum_buchang      src/mecat2ref/mecat2ref_impl_large.cpp:44-131,133-271,274-891
//   find_location, extract_sequences, extend_candidate, rescue_clipped_align, output_results
//                                                                         src/mecat2ref/mecat2ref_aux.cpp:6-473, mecat2ref_aux.h:9-43
//   chang_fastqfile, output_query_results, get_chr_id                     src/mecat2ref/mecat2ref.cpp:192-248,280-356
//   print_ref_result, print_m4_result                                     src/mecat2ref/output.cpp:8-88
// The gapped extension itself is the pw / ref flavour already restated in oracle_pw.cpp (orc_diff_go).
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

static int g_ref_tech = 0;      // -x of the run in progress (orc_ref_map_x)

namespace {

// mecat2ref_defs.h:16-28
const int ZV = 1000, ZVS = 2000, SM = 20, SI = 21, CLIPPED = 2000;
const int SEED_LEN = 13;                                  // meap_ref_impl_large, mecat2ref_impl_large.cpp:952
const double DDFS_CUTOFF = 0.25;                          // pacbio, :13

struct BackList { short score, score2, loczhi[SM], seedno[SM], seednum; int index; };           // Back_List
struct Candidate { long loc1, loc2, left1, left2, right1, right2; int score, num1, num2; char chain; };
struct AlignInfo
{
	int qid, qoff, qend, parent_id, id, prev_id, next_id;
	char valid, qdir;
	long soff, send;
	bool operator<(const AlignInfo& r) const { return (qend - qoff) > (r.qend - r.qoff); }
};
struct TempResult { int read_id; char read_dir; int vscore, qb, qe, qs; long sb, se; std::string qmap, smap; };
struct Chr { long start, size; std::string name; };

struct RefIndex
{
	std::string seq;                 // REFSEQ: every chromosome, upper case, concatenated
	std::vector<Chr> chr;
	std::vector<int> countin;        // occurrences per 13-mer (0 when above 128)
	std::vector<long> begin;         // start of the k-mer's list in pos
	std::vector<long> pos;           // 1-based start positions, ascending
};

int atct(char c) { return c == 'A' ? 0 : c == 'T' ? 1 : c == 'C' ? 2 : c == 'G' ? 3 : 4; }    // atcttrans: A0 T1 C2 G3

// creat_ref_index, mecat2ref_impl_large.cpp:133-271
bool build_index(const char* path, RefIndex& I)
{
	FILE* f = fopen(path, "r");
	if (!f) return false;
	std::string name;
	long rsize = 0;
	bool have = false;
	for (int ch = getc(f); ch != EOF; ch = getc(f)) {
		if (ch == '>') {
			char line[4096];
			line[0] = 0;
			if (fscanf(f, "%4095[^\n]", line) != 1) line[0] = 0;
			if (have) I.chr.back().size = rsize;
			rsize = 0;
			size_t k = 0;
			while (line[k] && line[k] != ' ' && line[k] != '\t') ++k;
			line[k] = 0;
			Chr c; c.start = (long)I.seq.size(); c.size = 0; c.name = line;
			I.chr.push_back(c);
			have = true;
		} else if (ch != '\n' && ch != '\r') {
			if (ch > 'Z') ch = toupper(ch);
			I.seq.push_back((char)ch);
			++rsize;
		}
	}
	fclose(f);
	if (have) I.chr.back().size = rsize;
	const long n = (long)I.seq.size();
	const int ncodes = 1 << (2 * SEED_LEN);
	const unsigned mask = (1u << (2 * (SEED_LEN - 1))) - 1;      // the reference keeps the last 12 bases (eit << 8 >> 8) and shifts the next one in
	I.countin.assign(ncodes, 0);
	for (int pass = 0; pass < 2; ++pass) {
		unsigned eit = 0;
		long start = 0;
		for (long i = 0; i < n; ++i) {
			const int t = atct(I.seq[i]);
			if (I.seq[i] == 'N' || t == 4) { eit = 0; start = 0; continue; }
			eit = (eit << 2) + (unsigned)t;
			++start;
			if (start >= SEED_LEN) {
				if (pass == 0) ++I.countin[eit];
				else if (I.begin[eit] >= 0) { I.pos[I.begin[eit] + I.countin[eit]] = i + 2 - SEED_LEN; ++I.countin[eit]; }
				eit &= mask;
			}
		}
		if (pass == 0) {
			// sumvalue_x: k-mers seen more than 128 times are dropped
			long total = 0;
			I.begin.assign(ncodes, -1);
			for (int c = 0; c < ncodes; ++c) {
				if (I.countin[c] > 128) I.countin[c] = 0;
				if (I.countin[c] > 0) { I.begin[c] = total; total += I.countin[c]; I.countin[c] = 0; }
			}
			I.pos.assign((size_t)total, 0);
		}
	}
	return true;
}

// transnum_buchang, :64-90: every BC-th 13-mer of the read, -1 when it holds a non-ACGT letter
int sample_kmers(const std::string& s, std::vector<int>& value, int BC)
{
	const int len = (int)s.size();
	if (len < SEED_LEN) { value.clear(); return 0; }
	const int num = (len - SEED_LEN) / BC + 1;
	value.assign((size_t)num, 0);
	for (int i = 0; i < num; ++i) {
		int eit = 0;
		for (int j = 0; j < SEED_LEN; ++j) {
			const int t = atct(s[i * BC + j]);
			if (t == 4) { eit = -1; break; }
			eit = (eit << 2) + t;
		}
		value[i] = eit;
	}
	return num;
}

// insert_loc, :92-130 (len is the float the reference passes BC as)
void insert_loc(BackList* spr, int loc, int seedn, float len)
{
	int list_loc[SI], list_score[SI], list_seed[SI];
	for (int i = 0; i < SM; ++i) { list_loc[i] = spr->loczhi[i]; list_seed[i] = spr->seedno[i]; list_score[i] = 0; }
	list_loc[SM] = loc; list_seed[SM] = seedn; list_score[SM] = 0;
	for (int i = 0; i < SM; ++i)
		for (int j = i + 1; j < SI; ++j)
			if (list_seed[j] - list_seed[i] > 0 && list_loc[j] - list_loc[i] > 0 &&
			    fabs((list_loc[j] - list_loc[i]) / ((list_seed[j] - list_seed[i]) * len) - 1.0) < DDFS_CUTOFF) {
				++list_score[i]; ++list_score[j];
			}
	int mini = -1, minval = 10000;
	for (int i = 0; i < SI; ++i) if (minval > list_score[i]) { minval = list_score[i]; mini = i; }
	if (minval == SM) { spr->loczhi[SM - 1] = (short)loc; spr->seedno[SM - 1] = (short)seedn; }
	else if (minval < SM && mini < SM) {
		for (int i = mini; i < SM; ++i) { spr->loczhi[i] = (short)list_loc[i + 1]; spr->seedno[i] = (short)list_seed[i + 1]; }
		--spr->score;
	}
}

// find_location, mecat2ref_aux.cpp:6-84 (float arithmetic as written there: int / (int * float) - int, compared with a double)
int find_location(const int* t_loc, const int* t_seedn, int* t_score, long* loc, int k, int* rep_loc, float len, int read_len1)
{
	int maxval = 0, maxi = 0, rep = 0, lasti = 0;
	for (int i = 0; i < k; ++i) t_score[i] = 0;
	for (int i = 0; i < k - 1; ++i)
		for (int j = i + 1; j < k; ++j)
			if (t_seedn[j] - t_seedn[i] > 0 && t_loc[j] - t_loc[i] > 0 && t_loc[j] - t_loc[i] < read_len1 &&
			    fabs((t_loc[j] - t_loc[i]) / ((t_seedn[j] - t_seedn[i]) * len) - 1) < DDFS_CUTOFF) {
				++t_score[i]; ++t_score[j];
			}
	for (int i = 0; i < k; ++i) {
		if (maxval < t_score[i]) { maxval = t_score[i]; maxi = i; rep = 0; }
		else if (maxval == t_score[i]) { ++rep; lasti = i; }
	}
	for (int i = 0; i < 4; ++i) loc[i] = 0;
	if (maxval >= 5 && rep == maxval) {
		loc[0] = t_loc[maxi]; loc[1] = t_seedn[maxi];
		*rep_loc = maxi;
		loc[2] = t_loc[lasti]; loc[3] = t_seedn[lasti];
		return 1;
	}
	if (maxval >= 5 && rep != maxval) {
		auto take = [&](int j) {
			if (loc[0] == 0) { loc[0] = t_loc[j]; loc[1] = t_seedn[j]; *rep_loc = j; }
			else { loc[2] = t_loc[j]; loc[3] = t_seedn[j]; }
		};
		for (int j = 0; j < maxi; ++j)
			if (t_seedn[maxi] - t_seedn[j] > 0 && t_loc[maxi] - t_loc[j] > 0 && t_loc[maxi] - t_loc[j] < read_len1 &&
			    fabs((t_loc[maxi] - t_loc[j]) / ((t_seedn[maxi] - t_seedn[j]) * len) - 1) < DDFS_CUTOFF) take(j);
		take(maxi);
		for (int j = maxi + 1; j < k; ++j)
			if (t_seedn[j] - t_seedn[maxi] > 0 && t_loc[j] - t_loc[maxi] > 0 && t_loc[j] - t_loc[maxi] <= read_len1 &&
			    fabs((t_loc[j] - t_loc[maxi]) / ((t_seedn[j] - t_seedn[maxi]) * len) - 1) < DDFS_CUTOFF) take(j);
		return 1;
	}
	return 0;
}

std::string revcomp(const std::string& s)      // reference_mapping: reverse, then complement ACGT only
{
	std::string r(s.rbegin(), s.rend());
	for (char& c : r) c = c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c;
	return r;
}

struct Strand
{
	std::vector<BackList> database;
	std::vector<int> index_list;
	std::vector<short> index_score;
	int nblk = 0;
	void init(long seqcount)
	{
		const size_t n = (size_t)(seqcount / ZV + 5);
		database.assign(n, BackList());
		for (auto& b : database) { b.score = 0; b.score2 = 0; b.index = -1; b.seednum = 0; }
		index_list.assign(n, 0); index_score.assign(n, 0);
	}
	void reset()      // mecat2ref_impl_large.cpp:629-640
	{
		for (int t = 0; t < nblk; ++t) { BackList& b = database[index_list[t]]; b.score = 0; b.score2 = 0; b.index = -1; }
	}
};

const unsigned char* encode_table()            // get_dna_encode_table: A0 C1 G2 T3 (either case), everything else > 3
{
	static unsigned char t[256];
	static bool done = false;
	if (!done) {
		memset(t, 16, sizeof t);
		const char* lo = "-acmgrsvtwyhkdbn";
		const unsigned char val[] = {15, 0, 1, 6, 2, 4, 9, 13, 3, 8, 5, 12, 7, 11, 10, 14};
		for (int i = 0; lo[i]; ++i) { t[(unsigned char)lo[i]] = val[i]; if (lo[i] != '-') t[(unsigned char)(lo[i] - 'a' + 'A')] = val[i]; }
		done = true;
	}
	return t;
}

struct Mapper
{
	const RefIndex& I;
	int maxc;
	Strand fwd, rev;
	std::vector<AlignInfo> alns;
	std::vector<TempResult> results;
	explicit Mapper(const RefIndex& idx, int maxc_) : I(idx), maxc(maxc_) { fwd.init((long)I.seq.size()); rev.init((long)I.seq.size()); }

	// extract_sequences + extend_candidate, mecat2ref_aux.cpp:86-183
	bool extend(const Candidate& can, const std::string& fwd_read, const std::string& rev_read, int read_name, bool record_aln)
	{
		const std::string& raw = can.chain == 'F' ? fwd_read : rev_read;
		const int read_len = (int)raw.size();
		const int read_start = (int)can.loc2;
		const long ref_start = can.loc1 - 1, ref_size = (long)I.seq.size();
		const long L1 = read_start, R1 = read_len - read_start, L2 = ref_start, R2 = ref_size - ref_start;
		const long L = std::min(L1, L2), R = std::min(R1, R2);
		const long left_ref = std::min(L2, (long)(L * 1.2)), right_ref = std::min(R2, (long)(R * 1.2));
		const unsigned char* et = encode_table();
		std::vector<char> q((size_t)read_len), t((size_t)(left_ref + right_ref));
		for (long i = 0; i < left_ref + right_ref; ++i) { unsigned char c = et[(unsigned char)I.seq[(size_t)(ref_start - left_ref + i)]]; t[(size_t)i] = (char)(c > 3 ? 0 : c); }
		for (int i = 0; i < read_len; ++i) { unsigned char c = et[(unsigned char)raw[(size_t)i]]; q[(size_t)i] = (char)(c > 3 ? 0 : c); }
		int32_t o[8];
		double ident;
		const int cap = 2 * (int)(q.size() + t.size()) + 64;
		std::vector<char> qs((size_t)cap), ts((size_t)cap);
		// the aligner follows the technology: DiffAligner, or XdropAligner for -x 1 (mecat2ref_impl_large.cpp:329-332)
		if (!(g_ref_tech == 1 ? orc_xdrop_go : orc_diff_go)(q.data(), read_start, read_len, t.data(), (int)left_ref, (int)t.size(), 1000, o, &ident, qs.data(), ts.data(), cap)) return false;
		TempResult r;
		r.read_id = read_name; r.read_dir = can.chain; r.vscore = can.score;
		r.qb = o[1]; r.qe = o[2]; r.qs = read_len;
		r.sb = ref_start - left_ref + o[3]; r.se = ref_start - left_ref + o[4];
		r.qmap = qs.data(); r.smap = ts.data();
		results.push_back(r);
		if (record_aln) {
			AlignInfo a;
			a.qid = 0; a.qoff = r.qb; a.qend = r.qe; a.qdir = r.read_dir; a.soff = r.sb; a.send = r.se; a.valid = 1;
			a.id = (int)results.size() - 1; a.prev_id = -1; a.next_id = -1; a.parent_id = -1;
			alns.push_back(a);
		}
		return true;
	}

	// seeding + candidate selection of one strand (the body of the ii loop, mecat2ref_impl_large.cpp:407-614)
	void strand_candidates(const std::string& read, Strand& S, int BC, int zv, int gate, char chain, std::vector<Candidate>& cand)
	{
		const int read_len = (int)read.size();
		const long seqcount = (long)I.seq.size();
		std::vector<int> mvalue;
		const int cleave = sample_kmers(read, mvalue, BC);
		BackList* database = S.database.data();
		int j = 0;
		for (int k = 0; k < cleave; ++k) {
			if (mvalue[k] < 0) continue;
			const int count1 = I.countin[mvalue[k]];
			const long* lead = count1 ? &I.pos[(size_t)I.begin[mvalue[k]]] : NULL;
			for (int i = 0; i < count1; ++i) {
				const int templong = (int)(lead[i] / zv);
				const long u_k = lead[i] % zv;
				if (templong < 0) continue;
				BackList* spr = database + templong;
				if (spr->score == 0 || spr->seednum < k + 1) {
					const long loc = ++spr->score;
					if (loc <= SM) { spr->loczhi[loc - 1] = (short)u_k; spr->seedno[loc - 1] = (short)(k + 1); }
					else insert_loc(spr, (int)u_k, k + 1, (float)BC);
					const long s_k = templong > 0 ? spr->score + (spr - 1)->score : spr->score;
					if (spr->index == -1) { S.index_list[j] = templong; S.index_score[j] = (short)s_k; spr->index = j; ++j; }
					else S.index_score[spr->index] = (short)s_k;
					spr->score2 = spr->score;
				}
				spr->seednum = (short)(k + 1);
			}
		}
		S.nblk = j;
		const int cc1 = j;
		int temp_list[200], temp_seedn[200], temp_score[200];
		for (int i = 0; i < cc1; ++i) {
			if (!(S.index_score[i] > gate)) continue;
			const int blk = S.index_list[i];
			BackList* spr = database + blk;
			if (spr->score == 0) continue;
			const long s_k = spr->score;
			long loc = 0, start_loc = (long)blk * zv;
			if (blk > 0) { loc = (spr - 1)->score; if (loc > 0) start_loc = (long)(blk - 1) * zv; }
			int u = 0;
			if (loc == 0) {
				for (int q = 0; q < s_k && q < SM; ++q) { temp_list[u] = spr->loczhi[q]; temp_seedn[u] = spr->seedno[q]; ++u; }
			} else {
				const BackList* prev = spr - 1;
				for (int q = 0; q < loc && q < SM; ++q) { temp_list[u] = prev->loczhi[q]; temp_seedn[u] = prev->seedno[q]; ++u; }
				for (int q = 0; q < s_k && q < SM; ++q) { temp_list[u] = spr->loczhi[q] + zv; temp_seedn[u] = spr->seedno[q]; ++u; }
			}
			long location_loc[4];
			int repeat_loc = 0;
			if (!find_location(temp_list, temp_seedn, temp_score, location_loc, u, &repeat_loc, (float)BC, read_len)) continue;
			if (temp_score[repeat_loc] < 6) continue;
			Candidate c;
			c.score = temp_score[repeat_loc];
			const int loc_seed = temp_seedn[repeat_loc];
			location_loc[0] = start_loc + location_loc[0];
			location_loc[1] = (location_loc[1] - 1) * BC;
			const long loc_list = location_loc[0];
			const long left1 = location_loc[0] + SEED_LEN - 1, right1 = seqcount - location_loc[0];
			const long left2 = location_loc[1] + SEED_LEN - 1, right2 = read_len - location_loc[1];
			const int num1 = (int)(left1 >= left2 ? left2 : left1), num2 = (int)(right1 >= right2 ? right2 : right1);
			int seedcount = 0;
			c.loc1 = location_loc[0]; c.num1 = num1; c.loc2 = location_loc[1]; c.num2 = num2;
			c.left1 = left1; c.left2 = left2; c.right1 = right1; c.right2 = right2;
			// votes of the blocks further left / right; a block that mostly agrees is consumed
			{
				long bk = blk - 2;
				int k = num1 / zv;
				for (BackList* p = spr - 2; bk >= 0 && k >= 0; --p, --k, --bk) {
					if (!(p->score > 0)) continue;
					const long sl = bk * (long)zv;
					const int scnt = std::min((int)p->score, SM);
					long s = 0;
					for (int q = 0; q < scnt; ++q)
						if (fabs((loc_list - sl - p->loczhi[q]) / ((loc_seed - p->seedno[q]) * BC * 1.0) - 1.0) < DDFS_CUTOFF) { ++seedcount; ++s; }
					if (s * 1.0 / scnt > 0.4) p->score = 0;
				}
			}
			{
				long bk = blk + 1;
				int k = num2 / zv;
				for (BackList* p = spr + 1; k > 0; ++p, --k, ++bk) {
					if (!(p->score > 0)) continue;
					const long sl = bk * (long)zv;
					const int scnt = std::min((int)p->score, SM);
					long s = 0;
					for (int q = 0; q < scnt; ++q)
						if (fabs((sl + p->loczhi[q] - loc_list) / ((p->seedno[q] - loc_seed) * BC * 1.0) - 1.0) < DDFS_CUTOFF) { ++seedcount; ++s; }
					if (s * 1.0 / scnt > 0.4) p->score = 0;
				}
			}
			c.score += seedcount;
			c.chain = chain;
			// sorted insert into the list of at most maxc candidates (binary search on the score, equal scores behind)
			int n = (int)cand.size(), low = 0, high = n - 1;
			while (low <= high) {
				const int mid = (low + high) / 2;
				if (mid >= n || cand[(size_t)mid].score < c.score) high = mid - 1; else low = mid + 1;
			}
			if (n < maxc) { cand.insert(cand.begin() + (high + 1), c); }
			else if (high + 1 < maxc) { cand.insert(cand.begin() + (high + 1), c); cand.pop_back(); }
		}
	}

	// fill_clipped_candidate, mecat2ref_aux.cpp:186-213
	bool fill_clipped(const BackList* block, long bid, Candidate& can, char chain, int read_size, int BC, int block_size)
	{
		int seedn[SM], boff[SM], score[SM], rep_loc = 0;
		long locations[4];
		const int n = std::min((int)block->score2, SM);
		for (int i = 0; i < n; ++i) { seedn[i] = block->seedno[i]; boff[i] = block->loczhi[i]; score[i] = 0; }
		if (!find_location(boff, seedn, score, locations, n, &rep_loc, (float)BC, read_size)) return false;
		can.score = score[rep_loc]; can.chain = chain;
		can.loc1 = bid * block_size + locations[0];
		can.loc2 = (locations[1] - 1) * BC;
		return true;
	}
	bool left_clipped(const AlignInfo& aln, Candidate& can, const BackList* database, int block_size, int read_size, int BC)
	{
		if (aln.qoff <= CLIPPED || aln.soff <= CLIPPED) return false;
		int n = std::min(aln.qoff / block_size, (int)(aln.soff / block_size));
		int n2 = (int)(aln.soff / block_size);
		int max_score = 0;
		const BackList* block = NULL;
		long bid = -1;
		for (--n2; n >= 0 && n2 >= 0; --n, --n2)
			if (database[n2].score2 > max_score) { max_score = database[n2].score2; block = database + n2; bid = n2; }
		return block && block->score2 > 4 && fill_clipped(block, bid, can, aln.qdir, read_size, BC, block_size);
	}
	bool right_clipped(const AlignInfo& aln, Candidate& can, const BackList* database, int block_size, int read_size, long ref_size, int BC)
	{
		if (read_size - aln.qend <= CLIPPED || ref_size - aln.send <= CLIPPED) return false;
		int n = std::min((read_size - aln.qend) / block_size, (int)((ref_size - aln.send) / block_size));
		int max_score = 0;
		long bid = -1;
		const BackList* block = NULL;
		int k = (int)(aln.send / block_size) + 1;
		for (; n >= 0; --n, ++k)
			if (database[k].score2 > max_score) { max_score = database[k].score2; block = database + k; bid = k; }
		return block && block->score2 > 4 && fill_clipped(block, bid, can, aln.qdir, read_size, BC, block_size);
	}
	static bool is_full(const AlignInfo& a, int qsize) { return a.qend - a.qoff >= qsize * 0.9; }
	static bool contained(const AlignInfo& a, const AlignInfo& b)
	{
		const int extra = 100;
		return a.qdir == b.qdir && b.qoff + extra >= a.qoff && b.qend <= a.qend + extra && b.soff + extra >= a.soff && b.send <= a.send + extra;
	}
	static bool left_of(const AlignInfo& a, const AlignInfo& b)      // is_left_clipped_align
	{
		if (a.qdir != b.qdir) return false;
		if (abs(b.qend - a.qoff) <= 200 && a.soff - b.send > -200 && a.soff - b.send < 10000) return true;
		if (labs(b.send - a.soff) <= 200 && a.qoff - b.qend > -200 && a.qoff - b.qend < 10000) return true;
		return false;
	}
	static bool right_of(const AlignInfo& a, const AlignInfo& b)     // is_right_clipped_align
	{
		if (a.qdir != b.qdir) return false;
		if (abs(a.qend - b.qoff) <= 200 && b.soff - a.send > -200 && b.soff - a.send < 10000) return true;
		if (labs(a.send - b.soff) <= 200 && b.qoff - a.qend > -200 && b.qoff - a.qend < 10000) return true;
		return false;
	}

	// rescue_clipped_align, mecat2ref_aux.cpp:305-452
	void rescue(const std::string& fwd_read, const std::string& rev_read, int read_name, int read_len, int block_size, int BC)
	{
		int naln = (int)alns.size();
		alns.resize((size_t)naln + 8);
		AlignInfo* alnv = alns.data();
		std::sort(alnv, alnv + naln);
		for (int i = 0; i < naln - 1; ++i) {
			if (!alnv[i].valid) continue;
			for (int j = i + 1; j < naln; ++j) if (alnv[j].valid && contained(alnv[i], alnv[j])) alnv[j].valid = 0;
		}
		int k = 0;
		for (int i = 0; i < naln; ++i) if (alnv[i].valid) alnv[k++] = alnv[i];
		naln = k;
		auto finish = [&]() { alns.resize((size_t)naln); };
		if (is_full(alnv[0], read_len)) { finish(); return; }     // (alnv[0] is read even when nothing aligned, like the reference)
		for (int i = 0; i < naln - 1; ++i) {
			if (alnv[i].parent_id != -1) continue;
			for (int j = i + 1; j < naln; ++j) {
				if (alnv[j].parent_id != -1) continue;
				if (alnv[i].prev_id != -1 && left_of(alnv[i], alnv[j])) { alnv[i].prev_id = alnv[j].id; alnv[j].parent_id = alnv[i].id; }
				if (alnv[i].next_id != -1 && right_of(alnv[i], alnv[j])) { alnv[i].next_id = alnv[j].id; alnv[j].parent_id = alnv[i].id; }
			}
		}
		const int n = std::min(naln, 3);
		k = 0;
		Candidate can;
		memset(&can, 0, sizeof can);
		for (int i = 0; i < n; ++i) {
			if (alnv[i].parent_id != -1) continue;
			const BackList* database = alnv[i].qdir == 'F' ? fwd.database.data() : rev.database.data();
			for (int side = 0; side < 2; ++side) {
				const bool want = side == 0 ? (alnv[i].prev_id == -1 && left_clipped(alnv[i], can, database, block_size, read_len, BC))
				                            : (alnv[i].next_id == -1 && right_clipped(alnv[i], can, database, block_size, read_len, (long)I.seq.size(), BC));
				if (!want) continue;
				if (!extend(can, fwd_read, rev_read, read_name, false)) continue;
				alnv = alns.data();
				const TempResult& r = results.back();
				AlignInfo& ai = alnv[naln + k];
				ai.qid = 0; ai.qoff = r.qb; ai.qend = r.qe; ai.qdir = r.read_dir; ai.soff = r.sb; ai.send = r.se; ai.valid = 1;
				ai.id = (int)results.size() - 1; ai.prev_id = -1; ai.next_id = -1; ai.parent_id = -1;
				if (side == 0 ? left_of(alnv[i], ai) : right_of(alnv[i], ai)) {
					ai.parent_id = alnv[i].id;
					if (side == 0) alnv[i].prev_id = ai.id; else alnv[i].next_id = ai.id;
					++k;
				}
			}
		}
		if (!k) { finish(); return; }
		naln += k;
		std::sort(alnv, alnv + naln);
		k = 0;
		for (int i = 0; i < naln; ++i)
			if (is_full(alnv[i], read_len)) { alnv[i].parent_id = -1; alnv[i].prev_id = -1; alnv[i].next_id = -1; ++k; }
		if (k) naln = k;
		finish();
	}

	// one pass over one read (first with the adaptive stride and 1 000-base blocks, then, when nothing aligned, stride 5 and
	// 2 000-base blocks); appends the results in output order (output_results, mecat2ref_aux.cpp:454-473)
	void map_read(int read_name, const std::string& read, int num_output, std::vector<TempResult>& out)
	{
		const std::string rc = revcomp(read);
		const int read_len = (int)read.size();
		for (int pass = 0; pass < 2; ++pass) {
			int BC = pass == 0 ? 5 + read_len / 1000 : 5;
			if (BC > 20) BC = 20;
			const int zv = pass == 0 ? ZV : ZVS;
			std::vector<Candidate> cand;
			strand_candidates(read, fwd, BC, zv, pass == 0 ? 6 : 4, 'F', cand);
			strand_candidates(rc, rev, BC, zv, pass == 0 ? 6 : 4, 'R', cand);
			alns.clear(); results.clear();
			for (const Candidate& c : cand) extend(c, read, rc, read_name, true);
			const int naln0 = (int)alns.size();
			if (naln0 == 0) {
				// the reference still runs rescue_clipped_align on an empty list (it reads alnv[0] of its scratch array there);
				// nothing can come of it, the first pass just ends without output
			} else {
				rescue(read, rc, read_name, read_len, zv, BC);
			}
			int n = 0;
			for (size_t i = 0; i < alns.size() && n < num_output; ++i) {
				if (alns[i].parent_id != -1) continue;
				out.push_back(results[(size_t)alns[i].id]);
				if (alns[i].prev_id != -1) out.push_back(results[(size_t)alns[i].prev_id]);
				if (alns[i].next_id != -1) out.push_back(results[(size_t)alns[i].next_id]);
				++n;
			}
			fwd.reset(); rev.reset();
			if (naln0 != 0) break;
		}
	}
};

// chang_fastqfile, mecat2ref.cpp:192-248: FASTA reads are numbered from 0, FASTQ reads from 1
bool load_reads(const char* path, std::vector<std::pair<int, std::string>>& reads)
{
	FILE* f = fopen(path, "r");
	if (!f) return false;
	std::string all;
	char buf[1 << 16];
	size_t r;
	while ((r = fread(buf, 1, sizeof buf, f)) > 0) all.append(buf, r);
	fclose(f);
	if (all.empty()) return true;
	if (all[0] == '>') {
		size_t i = 0;
		int kk = 0;
		while (i < all.size()) {
			if (all[i] == '>') {
				while (i < all.size() && all[i] != '\n') ++i;
				reads.push_back(std::make_pair(kk++, std::string()));
			} else {
				if (all[i] != '\n' && all[i] != '\r') reads.back().second.push_back(all[i]);
				++i;
			}
		}
	} else {
		size_t i = 0;
		int kk = 0;
		std::vector<std::string> lines;
		while (i < all.size()) {
			const size_t e = all.find('\n', i);
			std::string l = all.substr(i, e == std::string::npos ? std::string::npos : e - i);
			if (!l.empty() && l[l.size() - 1] == '\r') l.erase(l.size() - 1);
			lines.push_back(l);
			if (e == std::string::npos) break;
			i = e + 1;
		}
		for (size_t k = 0; k + 3 < lines.size(); k += 4) reads.push_back(std::make_pair(++kk, lines[k + 1]));
	}
	return true;
}

int chr_of(const std::vector<Chr>& chr, long offset)      // get_chr_id, mecat2ref.cpp:280-298
{
	const int n = (int)chr.size();
	int left = 0, right = n, mid = 0;
	while (left < right) {
		mid = (left + right) >> 1;
		if (offset >= chr[(size_t)mid].start) {
			if (mid == n - 1) break;
			if (offset < chr[(size_t)mid + 1].start) break;
			left = mid + 1;
		} else right = mid;
	}
	return mid;
}

}  // namespace

extern "C" {

// mecat2ref -d reads -r reference -n num_candidates -b num_output -m format (0 = ref, 1 = m4), pacbio.  The text the
// reference writes (records of a read together, reads in input order); malloc'ed, free with orc_free.
int orc_ref_map(const char* reference_path, const char* reads_path, int num_candidates, int num_output, int format, char** text, size_t* bytes)
{
	RefIndex I;
	if (!build_index(reference_path, I)) return 1;
	std::vector<std::pair<int, std::string>> reads;
	if (!load_reads(reads_path, reads)) return 1;
	Mapper M(I, num_candidates);
	std::string out;
	char line[512];
	for (auto& rd : reads) {
		std::vector<TempResult> res;
		M.map_read(rd.first, rd.second, num_output, res);
		int cnt = 0;
		for (const TempResult& r : res) {       // output_query_results: at most num_output records per read
			const int sid = chr_of(I.chr, r.sb);
			const Chr& c = I.chr[(size_t)sid];
			int qb = r.qb, qe = r.qe;
			if (r.read_dir == 'R') { qb = r.qs - r.qe; qe = r.qs - r.qb; }
			if (format == 0) {
				snprintf(line, sizeof line, "%d\t%s\t%c\t%d\t%d\t%d\t%d\t%ld\t%ld\t%ld\n", r.read_id, c.name.c_str(), r.read_dir == 'R' ? 'R' : 'F',
				         r.vscore, qb, qe, r.qs, r.sb - c.start, r.se - c.start, c.size);
				out += line; out += r.qmap; out += '\n'; out += r.smap; out += '\n';
			} else {
				double ident = 0.0;
				const size_t n = r.qmap.size();
				for (size_t i = 0; i < n; ++i) if (r.qmap[i] == r.smap[i]) ident += 1.0;
				ident = ident / (double)n;
				ident *= 100.0;
				snprintf(line, sizeof line, "%d\t%s\t%.4f\t%d\t%d\t%d\t%d\t%d\t0\t%ld\t%ld\t%ld\n", r.read_id, c.name.c_str(), ident, r.vscore,
				         r.read_dir == 'F' ? 0 : 1, qb, qe, r.qs, r.sb - c.start, r.se - c.start, c.size);
				out += line;
			}
			if (++cnt == num_output) break;
		}
	}
	char* p = (char*)malloc(out.size() + 1);
	if (!p) return 1;
	memcpy(p, out.data(), out.size());
	p[out.size()] = 0;
	*text = p; *bytes = out.size();
	return 0;
}

// the same with -x: 0 = pacbio (DiffAligner), 1 = nanopore (XdropAligner); nothing else of mecat2ref depends on it
int orc_ref_map_x(const char* reference_path, const char* reads_path, int num_candidates, int num_output, int format, int tech,
                  char** text, size_t* bytes)
{
	g_ref_tech = tech;
	const int rc = orc_ref_map(reference_path, reads_path, num_candidates, num_output, format, text, bytes);
	g_ref_tech = 0;
	return rc;
}

}  // extern "C"
