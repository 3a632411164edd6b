// oracle/oracle_align.h -- TEST INFRASTRUCTURE ONLY (see oracle.h).
// The O(nd) furthest-reaching block aligner shared by the two flavours of the reference:
//   pw / ref : Align + GetAlignString, src/common/diff_gapalign.cpp:40-219
//   cns      : ns_banded_sw::Align,    src/mecat2cns/dw.cpp:146-305
// They differ only in the edit-distance budget (max_d) and in what happens when no block end
// is reached (pw falls back to the best x+y cell, cns reports nothing).
#ifndef MECAT_ORACLE_ALIGN_H
#define MECAT_ORACLE_ALIGN_H

#include <algorithm>
#include <utility>
#include <vector>

namespace orc {

struct Cell { int pre_k, x1, y1, x2, y2; };

struct BlockAln
{
	int q_e = 0, t_e = 0, dist = 0, size = 0;
	bool aligned = false;
	std::vector<char> q, t;   // codes 0-3, 4 = gap
};

struct DiffScratch
{
	std::vector<int> V, U;
	std::vector<std::vector<Cell>> rows;   // rows[d][ (k - min_k[d]) / 2 ]
	std::vector<int> row_min;
};

inline char seq_at(const char* s, int i, int fwd) { return fwd ? s[i] : s[-i]; }

// GetAlignString, diff_gapalign.cpp:40-104 (inlined in dw.cpp:226-299)
inline void trace_path(const char* q, const char* t, DiffScratch& W, int d, int k, int fwd, BlockAln& A)
{
	std::vector<std::pair<int, int>> pts;
	for (int cd = d, ck = k; cd >= 0; --cd) {
		const Cell& c = W.rows[cd][(ck - W.row_min[cd]) / 2];
		pts.emplace_back(c.x2, c.y2);
		pts.emplace_back(c.x1, c.y1);
		ck = c.pre_k;
	}
	A.q.clear(); A.t.clear();
	int cx = pts.back().first, cy = pts.back().second;
	for (int i = (int)pts.size() - 2; i >= 0; --i) {
		int nx = pts[i].first, ny = pts[i].second;
		if (cx == nx && cy == ny) continue;
		if (cx == nx) {
			for (int j = 0; j < ny - cy; ++j) { A.q.push_back(4); A.t.push_back(seq_at(t, cy + j, fwd)); }
		} else if (cy == ny) {
			for (int j = 0; j < nx - cx; ++j) { A.q.push_back(seq_at(q, cx + j, fwd)); A.t.push_back(4); }
		} else {
			for (int j = 0; j < nx - cx; ++j) A.q.push_back(seq_at(q, cx + j, fwd));
			for (int j = 0; j < ny - cy; ++j) A.t.push_back(seq_at(t, cy + j, fwd));
		}
		cx = nx; cy = ny;
	}
	A.size = (int)A.q.size();
}

// One block.  max_d: edit budget; fallback: take the best x+y cell when no end is reached.
inline void align_block(const char* q, int qlen, const char* t, int tlen, int tol, int max_d, bool fallback, int fwd,
                        DiffScratch& W, BlockAln& A)
{
	const int koff = max_d;
	W.V.assign(4096, 0); W.U.assign(4096, 0);   // diff_gapalign.cpp:232-233, dw.cpp:329-330
	W.rows.clear(); W.row_min.clear();
	int best_m = -1, best_x = -1, best_y = -1, best_d = 0, best_k = 0;
	int min_k = 0, max_k = 0, x = -1, y = -1, k = 0, d = 0;
	bool aligned = false;
	A = BlockAln();
	for (d = 0; d < max_d; ++d) {
		if (max_k - min_k > 2 * tol) break;
		W.rows.emplace_back(); W.row_min.push_back(min_k);
		std::vector<Cell>& row = W.rows.back();
		for (k = min_k; k <= max_k; k += 2) {
			Cell c;
			if (k == min_k || (k != max_k && W.V[k - 1 + koff] < W.V[k + 1 + koff])) { c.pre_k = k + 1; x = W.V[k + 1 + koff]; }
			else { c.pre_k = k - 1; x = W.V[k - 1 + koff] + 1; }
			y = x - k;
			c.x1 = x; c.y1 = y;
			while (x < qlen && y < tlen && seq_at(q, x, fwd) == seq_at(t, y, fwd)) { ++x; ++y; }
			c.x2 = x; c.y2 = y;
			row.push_back(c);
			W.V[k + koff] = x; W.U[k + koff] = x + y;
			if (x + y > best_m) { best_m = x + y; best_x = x; best_y = y; best_d = d; best_k = k; }
			if (x >= qlen || y >= tlen) { aligned = true; break; }
		}
		int lo = max_k, hi = min_k;
		for (int k2 = min_k; k2 <= max_k; k2 += 2)
			if (W.U[k2 + koff] >= best_m - tol) { lo = std::min(lo, k2); hi = std::max(hi, k2); }
		max_k = hi + 1; min_k = lo - 1;
		if (aligned) {
			A.q_e = x; A.t_e = y; A.dist = d; A.aligned = true;
			trace_path(q, t, W, d, k, fwd, A);
			return;
		}
	}
	if (fallback && best_x > 0) {
		A.q_e = best_x; A.t_e = best_y; A.dist = best_d;
		trace_path(q, t, W, best_d, best_k, fwd, A);
	}
}

}  // namespace orc
#endif
