// oracle/oracle_xdrop.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the gapped extension the reference uses for nanopore reads (`-x 1`):
//   XdropAligner::go            src/common/xdrop_gapalign.cpp:351-439
//   align_ex (block chain)      src/common/xdrop_gapalign.cpp:249-349
//   xdrop_align (one block)     src/common/xdrop_gapalign.cpp:10-204   (BLAST-style X-drop DP with trace-back)
//   script_to_aligned_string    src/common/xdrop_gapalign.cpp:206-247
//   retrieve_next_aln_block / trim_mismatch_end   src/common/gapalign.cpp:10-67
// Parameters: XdropAlignParameters::init(0), xdrop_gapalign.h:85-103 (reward 1, penalty -1, gap open 0, gap extend 1,
// X = 30, 500-base blocks).  Pinned against the unmodified class through oracle/ref_shim.cpp (ref_xdrop_go) in
// tests/test_oracle.py.
#include "oracle.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

namespace {

const int REWARD = 1, PENALTY = -1, GAP_OPEN = 0, GAP_EXTEND = 1, X_DROPOFF = 30, BLOCK = 500;
const int NEG = -100000000;                      // MIN_SCORE, xdrop_gapalign.cpp:8
enum { OP_SUB = 3, OP_GAP_A = 0, OP_GAP_B = 6, OP_MASK = 7, EXT_A = 0x10, EXT_B = 0x40 };   // xdrop_gapalign.h:27-36

inline int at(const char* s, int i, bool fwd) { return (unsigned char)(fwd ? s[i] : s[-i]); }

struct Cell { int best, best_gap; };

// One block.  Returns the ops of the optimal path from the END backwards (the order the trace-back visits them).
int block_align(const char* A, int M, const char* B, int N, bool fwd, int& ae, int& be, std::vector<unsigned char>& ops)
{
	ae = be = 0;
	ops.clear();
	if (M <= 0 || N <= 0) return 0;
	const int oe = GAP_OPEN + GAP_EXTEND;
	int xd = X_DROPOFF;
	if (xd < oe) xd = oe;
	std::vector<Cell> sc((size_t)N + 2);
	std::vector<std::vector<unsigned char>> rows((size_t)M + 1);   // rows[a][b - start[a]]
	std::vector<int> start((size_t)M + 1, 0);
	rows[0].assign((size_t)N + 2, 0);
	int score = -oe, i;
	sc[0].best = 0; sc[0].best_gap = -oe;
	for (i = 1; i <= N; ++i) {
		if (score < -xd) break;
		sc[i].best = score; sc[i].best_gap = score - oe;
		score -= GAP_EXTEND;
		rows[0][i] = OP_GAP_A;
	}
	int b_size = i, best_score = 0, first_b = 0;
	for (int a = 1; a <= M; ++a) {
		const int ac = at(A, a - 1, fwd);
		start[a] = first_b;
		std::vector<unsigned char>& row = rows[a];
		row.assign((size_t)(N + 2 - first_b), 0);
		const int orig = first_b;
		score = NEG;
		int gap_row = NEG, last_b = first_b, b;
		for (b = first_b; b < b_size; ++b) {
			const int bc = at(B, b, fwd);
			int gap_col = sc[b].best_gap;
			const int next = sc[b].best + (ac == bc ? REWARD : PENALTY);
			unsigned char script = OP_SUB;
			if (score < gap_col) { script = OP_GAP_B; score = gap_col; }
			if (score < gap_row) { script = OP_GAP_A; score = gap_row; }
			if (best_score - score > xd) {
				if (first_b == b) ++first_b;
				else sc[b].best = NEG;
			} else {
				last_b = b;
				if (score > best_score) { best_score = score; ae = a; be = b; }
				gap_col -= GAP_EXTEND;
				if (gap_col < score - oe) sc[b].best_gap = score - oe;
				else { sc[b].best_gap = gap_col; script += EXT_A; }          // the reference attaches the flags this way round
				gap_row -= GAP_EXTEND;
				if (gap_row < score - oe) gap_row = score - oe;
				else script += EXT_B;
				sc[b].best = score;
			}
			score = next;
			row[b - orig] = script;
		}
		if (first_b == b_size) break;
		if (last_b < b_size - 1) b_size = last_b + 1;
		else {
			while (gap_row >= best_score - xd && b_size < N) {
				sc[b_size].best = gap_row; sc[b_size].best_gap = gap_row - oe;
				gap_row -= GAP_EXTEND;
				row[b_size - orig] = OP_GAP_A;
				++b_size;
			}
		}
		if (b_size < N) { sc[b_size].best = NEG; sc[b_size].best_gap = NEG; ++b_size; }
	}
	int a = ae, b = be;
	unsigned char script = OP_SUB;
	while (a > 0 || b > 0) {
		const unsigned char next = rows[a][b - start[a]];
		switch (script) {
		case OP_GAP_A: script = next & OP_MASK; if (next & EXT_A) script = OP_GAP_A; break;
		case OP_GAP_B: script = next & OP_MASK; if (next & EXT_B) script = OP_GAP_B; break;
		default: script = next & OP_MASK; break;
		}
		if (script == OP_GAP_A) --b;
		else if (script == OP_GAP_B) --a;
		else { --a; --b; }
		ops.push_back(script);
	}
	return best_score;
}

// align_ex: the chain of blocks in one direction; strings hold codes 0-3 and 4 = gap, in extension order
void chain(const char* q, int qsize, const char* t, int tsize, bool fwd, std::string& qa, std::string& ta)
{
	qa.clear(); ta.clear();
	int qi = 0, ti = 0;
	std::vector<unsigned char> ops;
	std::string bq, bt;
	for (;;) {
		const int qleft = qsize - qi, tleft = tsize - ti;
		int qblk, tblk;
		bool last;
		if (qleft < BLOCK + 100 || tleft < BLOCK + 100) {
			qblk = std::min(qleft, (int)(tleft + tleft * 0.2));
			tblk = std::min(tleft, (int)(qleft + qleft * 0.2));
			last = true;
		} else { qblk = tblk = BLOCK; last = false; }
		const char* Q = fwd ? q + qi : q - qi;
		const char* T = fwd ? t + ti : t - ti;
		int ae, be;
		block_align(Q, qblk, T, tblk, fwd, ae, be, ops);
		bq.clear(); bt.clear();
		int x = 0, y = 0;
		for (size_t k = ops.size(); k-- > 0;) {
			if (ops[k] == OP_SUB) { bq += (char)at(Q, x++, fwd); bt += (char)at(T, y++, fwd); }
			else if (ops[k] == OP_GAP_A) { bq += (char)4; bt += (char)at(T, y++, fwd); }
			else { bq += (char)at(Q, x++, fwd); bt += (char)4; }
		}
		const bool full = (qblk - ae <= 20 || tblk - be <= 20);
		if (!full || last) { qa += bq; ta += bt; break; }
		int m = 0, k, qc = 0, tc = 0, ac = 0;
		for (k = (int)bq.size() - 1; k >= 0 && m < 4; --k) {
			++ac;
			if (bq[k] != 4) ++qc;
			if (bt[k] != 4) ++tc;
			if (bq[k] == bt[k]) ++m; else m = 0;
		}
		if (!(m == 4 && k > 0)) break;
		qa.append(bq, 0, bq.size() - ac); ta.append(bt, 0, bt.size() - ac);
		qi += ae - qc; ti += be - tc;
	}
}

}  // namespace

extern "C" int orc_xdrop_go(const char* q, int qstart, int qsize, const char* t, int tstart, int tsize, int min_aln,
                            int32_t* out, double* ident, char* qstr, char* tstr, int cap)
{
	std::string lq, lt, rq, rt;
	chain(q + qstart - 1, qstart, t + tstart - 1, tstart, false, lq, lt);
	chain(q + qstart, qsize - qstart, t + tstart, tsize - tstart, true, rq, rt);
	static const char dt[] = "ACGT-";
	std::string oq, ot;
	int i = 0, j = 0;
	// the left part is emitted from its last column but one (xdrop_gapalign.cpp:392-393)
	for (int k = (int)lq.size() - 2; k >= 0; --k) {
		if (lq[k] != 4) ++i;
		if (lt[k] != 4) ++j;
		oq += dt[(int)lq[k]]; ot += dt[(int)lt[k]];
	}
	const int qoff = qstart - i, toff = tstart - j;
	i = j = 0;
	for (size_t k = 0; k < rq.size(); ++k) {
		if (rq[k] != 4) ++i;
		if (rt[k] != 4) ++j;
		oq += dt[(int)rq[k]]; ot += dt[(int)rt[k]];
	}
	const int qend = qstart + i, tend = tstart + j;
	int mat = 0;
	for (size_t k = 0; k < oq.size(); ++k) mat += oq[k] == ot[k];
	const int ok = qend - qoff >= min_aln;
	out[0] = ok; out[1] = qoff; out[2] = qend; out[3] = toff; out[4] = tend; out[5] = (int)oq.size(); out[6] = mat;
	if (ident) *ident = oq.empty() ? 0.0 : 100.0 * mat / (int)oq.size();
	if (qstr && (int)oq.size() < cap) { memcpy(qstr, oq.c_str(), oq.size() + 1); memcpy(tstr, ot.c_str(), ot.size() + 1); }
	return ok;
}
