// oracle/oracle_cns.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the consensus-flavour gapped extension of mecat2cns:
//   ns_banded_sw::Align / dw_in_one_direction / dw / GetAlignment   src/mecat2cns/dw.cpp:146-553
//   normalize_gaps                                                  src/mecat2cns/reads_correction_aux.cpp:3-79
// Pinned against the unmodified reference through oracle/_ref/libmecatref.so
// (tests/test_oracle.py::test_cns_alignment_against_reference).
#include "oracle.h"
#include "oracle_align.h"

#include <cstring>
#include <string>
#include <vector>

namespace {

using orc::BlockAln;
using orc::DiffScratch;

// dw_in_one_direction, dw.cpp:307-376: square blocks of 500 (last: whatever is left when <= 600),
// a failed block ends the chain, non-last blocks are cut in front of their last 4-match run.
void cns_extend_one_way(const char* q, int qsize, const char* t, int tsize, int fwd, double err, DiffScratch& W,
                        std::vector<char>& oq, std::vector<char>& ot)
{
	int e1 = 0, e2 = 0;
	int extend_size = std::min(qsize, tsize);
	bool more = true;
	BlockAln A;
	while (more) {
		int seg;
		if (extend_size > 600) seg = 500;
		else { seg = extend_size; more = false; }
		const char* Q = fwd ? q + e1 : q - e1;
		const char* T = fwd ? t + e2 : t - e2;
		orc::align_block(Q, seg, T, seg, (int)(0.3 * seg), (int)(2.0 * err * (seg + seg)), false, fwd, W, A);
		bool ok = A.aligned && (A.q_e == seg || A.t_e == seg);
		if (!ok) break;
		int k, i = 0, j = 0, m = 0;
		for (k = A.size - 1; k > -1 && m < 4; --k) {
			if (A.q[k] != 4) ++i;
			if (A.t[k] != 4) ++j;
			m = (A.q[k] == A.t[k]) ? m + 1 : 0;
		}
		if (more) {
			i = 500 - A.q_e + i; j = 500 - A.t_e + j;
			if (i == 500) ok = false;
			e1 += 500 - i; e2 += 500 - j;
		} else {
			i = extend_size - A.q_e; j = extend_size - A.t_e;
			if (i == extend_size) ok = false;
			e1 += extend_size - i; e2 += extend_size - j;
			k = A.size - 1;
		}
		if (!ok) break;
		oq.insert(oq.end(), A.q.begin(), A.q.begin() + k + 1);
		ot.insert(ot.end(), A.t.begin(), A.t.begin() + k + 1);
		extend_size = std::min(qsize - e1, tsize - e2);
	}
}

}  // namespace

extern "C" {

// dw + GetAlignment, dw.cpp:378-553
int orc_cns_get_alignment(const char* q, int qstart, int qsize, const char* t, int tstart, int tsize, double err,
                          int min_aln, int32_t* out, char* qaln, char* saln, int cap)
{
	DiffScratch W;
	std::vector<char> lq, lt, rq, rt;
	cns_extend_one_way(q + qstart - 1, qstart, t + tstart - 1, tstart, 0, err, W, lq, lt);
	cns_extend_one_way(q + qstart, qsize - qstart, t + tstart, tsize - tstart, 1, err, W, rq, rt);
	std::string a, b;
	int li = 0, lj = 0, ri = 0, rj = 0;
	for (size_t i = lq.size(); i-- > 0;) { a.push_back("ACGT-"[(int)lq[i]]); b.push_back("ACGT-"[(int)lt[i]]); li += lq[i] != 4; lj += lt[i] != 4; }
	for (size_t i = 0; i < rq.size(); ++i) { a.push_back("ACGT-"[(int)rq[i]]); b.push_back("ACGT-"[(int)rt[i]]); ri += rq[i] != 4; rj += rt[i] != 4; }
	const int qs = qstart - li, ts = tstart - lj, qe = qstart + ri, te = tstart + rj;
	const int n = (int)a.size();
	out[0] = 0;
	if (n < min_aln) return 0;
	int qrb = 0, trb = 0, run = 0, k = 0;
	for (k = 0; k < n && run < 4; ++k) {
		if (a[k] != '-') ++qrb;
		if (b[k] != '-') ++trb;
		run = (a[k] == b[k]) ? run + 1 : 0;
	}
	if (run < 4) return 0;
	k -= 4; qrb -= 4; trb -= 4;
	const int start_id = k;
	int qre = 0, tre = 0;
	for (k = n - 1, run = 0; k >= 0 && run < 4; --k) {
		if (a[k] != '-') ++qre;
		if (b[k] != '-') ++tre;
		run = (a[k] == b[k]) ? run + 1 : 0;
	}
	if (run < 4) return 0;
	k += 4; qre -= 4; tre -= 4;
	const int end_id = k + 1, size = end_id - start_id;
	out[0] = 1; out[1] = qs + qrb; out[2] = qe - qre; out[3] = ts + trb; out[4] = te - tre;
	if (qaln && size + 1 <= cap) {
		memcpy(qaln, a.data() + start_id, (size_t)size); qaln[size] = 0;
		memcpy(saln, b.data() + start_id, (size_t)size); saln[size] = 0;
	}
	return 1;
}

// normalize_gaps, reads_correction_aux.cpp:3-79
int orc_normalize_gaps(const char* qstr, const char* tstr, int n, int push, char* qout, char* tout, int cap)
{
	std::string qn, tn;
	for (int i = 0; i < n; ++i) {
		const char qc = qstr[i], tc = tstr[i];
		if (qc != tc && qc != '-' && tc != '-') { qn += '-'; qn += qc; tn += tc; tn += '-'; }
		else { qn += qc; tn += tc; }
	}
	if (push) {
		// the scan may read the terminating NUL at index size (std::string guarantees it)
		const int len = (int)qn.size();
		for (int i = 0; i < len - 1; ++i) {
			if (tn[i] == '-') {
				int j = i;
				while (true) {
					const char c = tn[++j];
					if (c != '-' || j > len - 1) {
						if (c == qn[i]) { tn[i] = c; tn[j] = '-'; }
						break;
					}
				}
			}
			if (qn[i] == '-') {
				int j = i;
				while (true) {
					const char c = qn[++j];
					if (c != '-' || j > len - 1) {
						if (c == tn[i]) { qn[i] = c; qn[j] = '-'; }
						break;
					}
				}
			}
		}
	}
	if ((int)qn.size() + 1 > cap) return -1;
	memcpy(qout, qn.c_str(), qn.size() + 1);
	memcpy(tout, tn.c_str(), tn.size() + 1);
	return (int)qn.size();
}

}  // extern "C"
