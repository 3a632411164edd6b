// tests/cns_host_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Runs the product's consensus stage sequence (mecat_b200/csrc/cns_pipeline.h) and kernel bodies
// (mecat_b200/csrc/cns_core.cuh) on the host: "device memory" is malloc, a kernel launch is a loop over the
// units, a scan is a loop.  This lets the CPU test-suite check the statements the GPU executes against the
// reference's golden output and the oracle without a GPU.  It is compiled by tests/util.py into
// tests/_build/libcns_harness.so and is never part of the product library.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../include/mecat_b200.h"
#include "../mecat_b200/csrc/cns_pipeline.h"
#include "cns_literal.h"

namespace {

struct HostBackend
{
	std::vector<void*> owned;
	std::string err;
	template <class T> T* alloc(size_t n)
	{
		void* p = calloc(n ? n : 1, sizeof(T));
		if (!p) { err = "out of memory"; return nullptr; }
		memset(p, 0xAB, (n ? n : 1) * sizeof(T));      // like device memory: never zero by luck
		owned.push_back(p);
		return (T*)p;
	}
	template <class T> bool upload(T* d, const T* h, size_t n) { if (n) memcpy(d, h, n * sizeof(T)); return true; }
	template <class T> bool download(T* h, const T* d, size_t n) { if (n) memcpy(h, d, n * sizeof(T)); return true; }
	const char* download_staged(const char* d, size_t) { return d; }
	bool fill(void* d, int byte, size_t bytes) { memset(d, byte, bytes); return true; }
	template <class F> bool launch(int64_t n, const F& f, int) { for (int64_t i = 0; i < n; ++i) f(i); return true; }
	template <class F> bool launch_warp(int64_t n, const F& f, int) { mbcns::EmuLanes one; for (int64_t i = 0; i < n; ++i) f(i, one); return true; }
	// graphs: narrow (int16_t) indices where the product uses them, or wide everywhere when a test asks for it
	bool launch_graphs(int64_t n, const mbcns::PoaFn& f, int)
	{
		for (int64_t k = 0; k < n; ++k) f(k, min_width);
		return true;
	}
	int min_width = 1;
	bool scan(const int32_t* in, int64_t* out, int64_t n, int64_t* total)
	{
		int64_t s = 0;
		for (int64_t i = 0; i < n; ++i) { out[i] = s; s += in[i]; }
		out[n] = s; *total = s;
		return true;
	}
	bool release(void* p)
	{
		for (size_t i = 0; i < owned.size(); ++i) if (owned[i] == p) { free(p); owned[i] = owned.back(); owned.pop_back(); return true; }
		err = "release of an unknown block";
		return false;
	}
	int64_t poa_budget_bytes() const { return budget; }
	int64_t budget = 1 << 20;          // small on purpose: the golden tests run the region graphs in many waves
	void fail(const char* m) { err = m; }
	void end_batch() { for (void* p : owned) free(p); owned.clear(); }
};

}  // namespace

extern "C" {

// R reads; read r owns the candidates / results [first[r], first[r+1]) in trial order.  Strings of result t start at
// res[t].str_offset in qstr / sstr (NUL terminated).
int harness_cns_batch(int R, const int32_t* first, const mecat_candidate* cand, const mecat_align_result* res, const char* qstr,
                      const char* sstr, const mecat_cns_params* p, mecat_cns_piece** pieces, size_t* npieces, char** seqs,
                      size_t* seq_bytes, char* errbuf, int errcap)
{
	const int64_t T = first[R];
	std::vector<int32_t> rsize((size_t)R), tqid((size_t)T), tqsize((size_t)T), info((size_t)T * 8);
	std::vector<int64_t> rid((size_t)R);
	std::vector<unsigned long long> outoff((size_t)T + 1, 0);
	for (int r = 0; r < R; ++r) { rsize[r] = cand[first[r]].ssize; rid[r] = cand[first[r]].sid; }
	for (int64_t t = 0; t < T; ++t) {
		tqid[t] = cand[t].qid; tqsize[t] = cand[t].qsize;
		int32_t* o = &info[8 * t];
		o[0] = res[t].ok; o[1] = res[t].qstart; o[2] = res[t].qend; o[3] = res[t].sstart; o[4] = res[t].send;
		o[5] = res[t].columns; o[6] = res[t].matches; o[7] = 0;
		outoff[t] = res[t].ok ? (unsigned long long)res[t].str_offset : 0;
	}
	// the kernels read the gapped strings through aligned 16-byte windows: give the blobs the slack device arenas have
	size_t blob = 0;
	for (int64_t t = 0; t < T; ++t) if (res[t].ok) blob = std::max(blob, (size_t)res[t].str_offset + (size_t)res[t].columns + 1);
	std::vector<char> qpad(blob + 64, 0), spad(blob + 64, 0);
	char* qa = qpad.data() + ((16 - ((uintptr_t)qpad.data() & 15)) & 15) + 16;
	char* sa = spad.data() + ((16 - ((uintptr_t)spad.data() & 15)) & 15) + 16;
	if (blob) { memcpy(qa, qstr, blob); memcpy(sa, sstr, blob); }
	qstr = qa; sstr = sa;
	mbcns::BatchIn in;
	in.R = R; in.T = T; in.h_first = first; in.h_read_size = rsize.data(); in.h_read_id = rid.data();
	in.h_tqid = tqid.data(); in.h_tqsize = tqsize.data();
	in.d_info = info.data(); in.d_q = qstr; in.d_s = sstr; in.d_outoff = outoff.data();
	mbcns::Params P;
	P.min_mapping_ratio = p->min_mapping_ratio; P.min_align_size = p->min_align_size; P.min_cov = p->min_cov; P.min_size = p->min_size;
	P.tech = p->tech; P.input_type = p->input_type;
	HostBackend be;
	if (const char* e = getenv("MECAT_HARNESS_GRAPH_INDEX_BYTES")) be.min_width = atoi(e);      // 2 or 4: wider indices than needed
	mbcns::PieceVector sink;
	std::vector<mbcns::Piece>& out = sink.pieces;
	if (mbcns::consensus_batch(be, in, P, sink)) {
		if (errbuf && errcap > 0) snprintf(errbuf, (size_t)errcap, "%s", be.err.c_str());
		return 1;
	}
	size_t bytes = 0;
	for (auto& pc : out) bytes += pc.seq.size();
	mecat_cns_piece* o = (mecat_cns_piece*)malloc(sizeof(mecat_cns_piece) * (out.size() ? out.size() : 1));
	char* sq = (char*)malloc(bytes + 1);
	if (!o || !sq) { free(o); free(sq); return 1; }
	size_t at = 0;
	for (size_t i = 0; i < out.size(); ++i) {
		o[i].id = out[i].id; o[i].beg = out[i].beg; o[i].end = out[i].end; o[i].seq_offset = (int64_t)at; o[i].seq_len = (int64_t)out[i].seq.size();
		memcpy(sq + at, out[i].seq.data(), out[i].seq.size());
		at += out[i].seq.size();
	}
	sq[bytes] = 0;
	*pieces = o; *npieces = out.size(); *seqs = sq; *seq_bytes = bytes;
	return 0;
}

void harness_free(void* p) { free(p); }

// The kernels' get_effective_ranges on n {start, end} pairs; returns the number of ranges written to out.
int harness_effective_ranges(const int* in, int n, int read_size, long long min_size, int* out)
{
	mbcns::Range m[mbcns::MAX_ACCEPT], e[mbcns::MAX_ACCEPT];
	for (int i = 0; i < n; ++i) { m[i].start = in[2 * i]; m[i].end = in[2 * i + 1]; }
	const int ne = mbcns::effective_ranges(m, n, e, read_size, 0.95 * (double)min_size);
	for (int i = 0; i < ne; ++i) { out[2 * i] = e[i].start; out[2 * i + 1] = e[i].end; }
	return ne;
}

// The kernels' one-pass normalise + vote on one gapped alignment: normalised strings to nq / nt, votes as
// positions x {base, mat, ins, del}.  Returns the normalised length.
int harness_normalize_and_vote(const char* q, const char* t, int n, int soff, int positions, unsigned char* out, char* nq_out, char* nt_out)
{
	std::vector<char> qp((size_t)n + 64, 0), tp((size_t)n + 64, 0);
	char* qa = qp.data() + ((16 - ((uintptr_t)qp.data() & 15)) & 15) + 16 + 1;
	char* ta = tp.data() + ((16 - ((uintptr_t)tp.data() & 15)) & 15) + 16 + 7;
	memcpy(qa, q, (size_t)n); memcpy(ta, t, (size_t)n);
	std::vector<unsigned long long> a((size_t)(2 * n + 64) / 8 + 1, 0), b((size_t)(2 * n + 64) / 8 + 1, 0);
	std::vector<uint32_t> votes((size_t)positions + 2, 0);
	std::vector<char> base((size_t)positions + 2, 'N');
	std::vector<int32_t> colidx((size_t)positions + 4, 0);
	int tend = 0;
	const int len = mbcns::normalize_vote_index(qa, ta, n, soff, (char*)a.data(), (char*)b.data(), votes.data(), base.data(), colidx.data(), &tend);
	for (int i = 0; i < positions; ++i) {
		out[4 * i] = (unsigned char)base[i]; out[4 * i + 1] = (unsigned char)mbcns::vote_mat(votes[i]);
		out[4 * i + 2] = (unsigned char)mbcns::vote_ins(votes[i]); out[4 * i + 3] = (unsigned char)mbcns::vote_del(votes[i]);
	}
	memcpy(nq_out, a.data(), (size_t)len + 1); memcpy(nt_out, b.data(), (size_t)len + 1);
	return len;
}

// One region graph with the product's flat-array implementation (PoaT<I>, I = 1 / 2 / 4 byte indices) in an arena of
// exactly the size the pipeline would give it.  Returns the consensus length (bytes in out), or -(100 + POA_ERR_*).
int harness_poa_consensus(int blen, int naln, const char* const* q, const char* const* t, const int* start, int min_weight,
                          int index_bytes, char* out, int cap)
{
	int nodes = blen + 2, e0 = blen + 1;
	for (int i = 0; i < naln; ++i) {
		for (const char *a = q[i], *b = t[i]; *a; ++a, ++b) {
			if (*a == '-') continue;
			++e0;
			if (*a != *b) ++nodes;
		}
		++e0;
	}
	const int ecap = (int)mbcns::poa_edge_cap(nodes, e0);
	std::vector<char> arena((size_t)mbcns::poa_arena_bytes<int32_t>(nodes, e0) + 16, (char)0xAB);
	std::vector<char> path((size_t)nodes + 1);
	int off = 0, len = 0, err = 0;
	auto run = [&](auto tag) {
		typedef decltype(tag) I;
		mbcns::PoaT<I> g;
		g.init(arena.data(), nodes, ecap, blen);
		for (int i = 0; i < naln; ++i) g.add_alignment(q[i], t[i], 0, (int)strlen(q[i]) - 1, start[i]);
		g.merge_nodes();
		g.consensus(min_weight, path.data(), off, len);
		err = g.err;
	};
	if (index_bytes == 1) run((int8_t)0); else if (index_bytes == 2) run((int16_t)0); else run((int32_t)0);
	if (err) return -(100 + err);
	if (len > cap) return -1;
	memcpy(out, path.data() + off, (size_t)len);
	return len;
}

// Unit check of the 32-positions-per-step anchor walk (anchor_chunk, what the GPU warps run) against the literal
// sequential walk (walk_anchors, written after meap_consensus_one_segment) on one flag array.  Returns 0 when the
// anchors, their ranks, the refine intervals, their ordinals and their chaining (prev_se) all agree.
int harness_anchor_compare(const uint8_t* flags, int n, int beg)
{
	struct Ival { int i, j; bool refine; };
	std::vector<Ival> want;
	mbcns::walk_anchors(flags, n, [&](int i, int j, bool refine) { want.push_back(Ival{i, j, refine}); });
	struct Seen { int pos, rank, prevpos, ordinal, prev_se; bool closes; };
	std::vector<Seen> got;
	mbcns::EmuLanes lanes;
	mbcns::AnchorCarry c;
	for (int base = 0; base < n; base += 32)
		mbcns::anchor_chunk(lanes, base, n, beg, [&](int p) { return (int)flags[p]; }, c,
		                    [&](int, int pos, int rank, int prevpos, bool closes, int ordinal, int prev_se) {
			got.push_back(Seen{pos, rank, prevpos, ordinal, prev_se, closes});
		});
	if (got.size() != want.size()) return 1;
	if ((int)want.size() != c.nanchors) return 2;
	int regions = 0, last_se = -1;
	for (size_t a = 0; a < want.size(); ++a) {
		const Seen& g = got[a];
		if (g.pos != want[a].i || g.rank != (int)a) return 3;
		if (g.prevpos != (a ? want[a - 1].i : -1)) return 4;
		const bool closes = a > 0 && want[a - 1].refine;          // the interval that ends at this anchor
		if (g.closes != closes) return 5;
		if (g.ordinal != regions) return 6;
		if (g.prev_se != last_se) return 7;
		if (closes) { ++regions; last_se = beg + g.pos; }
	}
	const bool tail = !want.empty() && want.back().refine;        // the interval from the last anchor to the end
	if ((c.last_anchor >= 0 && c.pending) != tail) return 8;
	if (c.nregions != regions || c.last_se_abs != last_se) return 9;
	if (!want.empty() && c.last_anchor != want.back().i) return 10;
	return 0;
}

// The 32-positions-per-step run search (find_segments with lanes) against the literal one.
int harness_segments_compare(const uint32_t* votes, const int32_t* ranges, int nranges, int min_cov, double size95)
{
	std::vector<mbcns::Range> e((size_t)nranges);
	for (int k = 0; k < nranges; ++k) { e[k].start = ranges[2 * k]; e[k].end = ranges[2 * k + 1]; }
	const int cap = 4096;
	std::vector<int32_t> a(2 * cap, -1), b(2 * cap, -1);
	const int na = mbcns::find_segments(e.data(), nranges, votes, min_cov, size95, a.data(), cap);
	const int nb = mbcns::find_segments(mbcns::EmuLanes(), e.data(), nranges, votes, min_cov, size95, b.data(), cap);
	if (na != nb) return 1;
	return a == b ? 0 : 2;
}

// Unit check of the fused single-pass kernel body (normalize_vote_index) against the three literal restatements
// (normalize_gaps, add_votes, column_index) on one gapped alignment.  Returns 0 when every output agrees.
int harness_normalize_compare(const char* q, const char* t, int n, int soff, int positions)
{
	std::vector<char> qp((size_t)n + 64, 0), tp((size_t)n + 64, 0);
	char* qa = qp.data() + ((16 - ((uintptr_t)qp.data() & 15)) & 15) + 16 + 3;      // deliberately unaligned start
	char* ta = tp.data() + ((16 - ((uintptr_t)tp.data() & 15)) & 15) + 16 + 5;
	memcpy(qa, q, (size_t)n); memcpy(ta, t, (size_t)n);
	const size_t cap = 2 * (size_t)n + 64;
	std::vector<char> nq1(cap, 1), nt1(cap, 1);
	std::vector<unsigned long long> nq2s(cap / 8 + 1, ~0ull), nt2s(cap / 8 + 1, ~0ull);
	char* nq2 = (char*)nq2s.data(); char* nt2 = (char*)nt2s.data();
	std::vector<uint32_t> v1((size_t)positions + 2, 0), v2((size_t)positions + 2, 0);
	std::vector<char> b1((size_t)positions + 2, 'N'), b2((size_t)positions + 2, 'N');
	std::vector<int32_t> c1((size_t)positions + 2, -7), c2((size_t)positions + 2, -7);
	const int len1 = mbcns::normalize_gaps(qa, ta, n, nq1.data(), nt1.data());
	mbcns::add_votes(nq1.data(), nt1.data(), len1, soff, v1.data() + 1, b1.data() + 1);
	const int tend1 = mbcns::column_index(nt1.data(), len1, soff, c1.data());
	int tend2 = -1;
	const int len2 = mbcns::normalize_vote_index(qa, ta, n, soff, nq2, nt2, v2.data() + 1, b2.data() + 1, c2.data(), &tend2);
	if (len1 != len2) return 1;
	if (memcmp(nq1.data(), nq2, (size_t)len1 + 1) || memcmp(nt1.data(), nt2, (size_t)len1 + 1)) return 2;
	if (v1 != v2) return 3;
	if (b1 != b2) return 4;
	if (tend1 != tend2) return 5;
	if (c1 != c2) return 6;
	return 0;
}

}  // extern "C"
