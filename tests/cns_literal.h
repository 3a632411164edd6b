// tests/cns_literal.h -- TEST INFRASTRUCTURE ONLY.
//
// Sequential, line-by-line restatements of the reference functions that the product's kernels compute in a different
// form (one streaming pass, 32 positions per step).  The unit tests in tests/cns_host_harness.cpp run both forms on
// random inputs and require identical results.  Nothing under mecat_b200/ includes this file.
//   normalize_gaps            src/mecat2cns/reads_correction_aux.cpp:3-79
//   meap_add_one_aln          src/mecat2cns/mecat_correction.cpp:37-60
//   CnsAln cursor             src/mecat2cns/reads_correction_aux.h:47-68
//   consensus_worker (runs)   src/mecat2cns/mecat_correction.cpp:203-239
//   meap_consensus_one_segment (anchor walk)   src/mecat2cns/mecat_correction.cpp:81-108
#pragma once
#include "../mecat_b200/csrc/cns_core.cuh"

namespace mbcns {

// ------------------------------------------------------------------------------------------ C4
// normalize_gaps: mismatch columns become a (gap, base)/(base, gap) pair, then every gap is pushed right
// while the base behind its run equals the base opposite.  nq/nt need 2n + 1 bytes; the terminator the
// reference's std::string supplies at index len is written explicitly.  Returns the normalised length.
CNS_HD inline int normalize_gaps(const char* q, const char* t, int n, char* nq, char* nt)
{
	int len = 0;
	for (int i = 0; i < n; ++i) {
		const char a = q[i], b = t[i];
		if (a != b && a != '-' && b != '-') { nq[len] = '-'; nt[len] = b; ++len; nq[len] = a; nt[len] = '-'; ++len; }
		else { nq[len] = a; nt[len] = b; ++len; }
	}
	nq[len] = 0; nt[len] = 0;
	for (int i = 0; i < len - 1; ++i) {
		if (nt[i] == '-') {
			int j = i;
			for (;;) {
				const char c = nt[++j];
				if (c != '-' || j > len - 1) { if (c == nq[i]) { nt[i] = c; nt[j] = '-'; } break; }
			}
		}
		if (nq[i] == '-') {
			int j = i;
			for (;;) {
				const char c = nq[++j];
				if (c != '-' || j > len - 1) { if (c == nt[i]) { nq[i] = c; nq[j] = '-'; } break; }
			}
		}
	}
	return len;
}

// ------------------------------------------------------------------------------------------ C5
// meap_add_one_aln on a normalised alignment.  votes/base are the read's arrays (index = template position).
CNS_HD inline void add_votes(const char* q, const char* s, int n, int soff, uint32_t* votes, char* base)
{
	int i = 0;
	while (i < n) {
		const char a = q[i], b = s[i];
		if (a == '-' && b == '-') { ++i; continue; }
		if (a == b) { vote_add(votes + soff, 1u); base[soff] = b; ++soff; ++i; }
		else if (a == '-') { vote_add(votes + soff, 1u << 8); ++soff; ++i; }
		else {
			int j = i + 1;
			while (j < n && s[j] == '-') ++j;
			vote_add(votes + soff - 1, 1u << 16);
			i = j;
		}
	}
}

// Column of every template position of a normalised alignment, the way CnsAln's cursor counts them
// (reads_correction_aux.h:47-68): the cursor starts on column 0 at template position soff and a column
// idx >= 1 advances the position iff its template character is a base.  colidx[p - soff] = first column
// at position p.  Returns the last position reached (the entries [0, ret - soff] are written).
CNS_HD inline int column_index(const char* s, int n, int soff, int32_t* colidx)
{
	int p = soff;
	colidx[0] = 0;
	for (int idx = 1; idx < n; ++idx)
		if (s[idx] != '-') { ++p; colidx[p - soff] = idx; }
	return p;
}


// consensus_worker's run search inside the effective ranges (mecat_correction.cpp:203-239): maximal runs
// with coverage >= min_cov that are at least 0.95 * min_size long.  segs receives up to cap {beg, end}
// pairs; the return value is the number found (callers size cap so that it always fits).
CNS_HD inline int find_segments(const Range* e, int ne, const uint32_t* votes, int min_cov, double size95, int32_t* segs, int cap)
{
	int ns = 0;
	for (int r = 0; r < ne; ++r) {
		const int R = e[r].end;
		int beg = e[r].start;
		while (beg < R) {
			while (beg < R && vote_mat(votes[beg]) + vote_ins(votes[beg]) < min_cov) ++beg;
			int end = beg + 1;
			while (end < R && vote_mat(votes[end]) + vote_ins(votes[end]) >= min_cov) ++end;
			if ((double)(end - beg) >= size95) {
				if (ns < cap) { segs[2 * ns] = beg; segs[2 * ns + 1] = end; }
				++ns;
			}
			beg = end;
		}
	}
	return ns;
}


// meap_consensus_one_segment's walk over the anchors (positions whose flag has FMAT).  flags[0..n) must hold
// classify() of the segment.  Calls on_anchor(i, j, refine) for each anchor i with next anchor j.
template <class F>
CNS_HD inline void walk_anchors(const uint8_t* flags, int n, F&& on_anchor)
{
	int i = 0;
	while (i < n && !(flags[i] & FMAT)) ++i;
	while (i < n) {
		int j = i + 1;
		while (j < n && !(flags[j] & FMAT)) ++j;
		bool refine = false;
		for (int k = i; k < j; ++k) if (flags[k] & (UNDS | FDEL)) { refine = true; break; }
		on_anchor(i, j, refine);
		i = j;
	}
}


}  // namespace mbcns
