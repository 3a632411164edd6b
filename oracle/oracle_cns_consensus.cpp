// oracle/oracle_cns_consensus.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// CPU restatement of the per-read consensus of mecat2cns (rows C3-C7), given the GetAlignment results of
// the read's candidates.  Pinned against corrected FASTA written by the UNMODIFIED reference binary
// (tests/golden/*.cns_*.fa.gz, tests/test_cns_host.py).  The product computes the same on the GPU
// (mecat_b200/csrc/cns.cu); this file is what the GPU kernels are checked against.  Semantics of
//   consensus_one_read_can_pacbio, check_ovlp_mapping_range, check_cov_stats   src/mecat2cns/mecat_correction.cpp:191-200,362-450
//   normalize_gaps                                                             src/mecat2cns/reads_correction_aux.cpp:3-79
//   meap_add_one_aln, identify_one_consensus_item, meap_consensus_one_segment  mecat_correction.cpp:15-108
//   get_effective_ranges, consensus_worker, output_cns_result                  mecat_correction.cpp:119-239
//   CnsAln::retrieve_aln_subseqs                                               src/mecat2cns/reads_correction_aux.h:47-68
//   AlnGraphBoost (mini partial-order graph of one ambiguous region)           src/mecat2cns/MECAT_AlnGraphBoost.C:76-592
// The graph is a small adjacency structure with the iteration orders of the reference's
// boost::adjacency_list<vecS, vecS, bidirectionalS> (edge lists in insertion order, order-preserving
// removal), because tie-breaks in the best-path search depend on them.
#include "oracle_cns_consensus.h"

#include <stdlib.h>

#include <algorithm>
#include <cfloat>
#include <cstring>
#include <map>
#include <queue>
#include <set>

namespace orccns {

namespace {

// ------------------------------------------------------------------ C4
void normalize_gaps(const char* q, const char* t, int n, std::string& qn, std::string& tn)
{
	qn.clear(); tn.clear();
	for (int i = 0; i < n; ++i) {
		const char a = q[i], b = t[i];
		if (a != b && a != '-' && b != '-') { qn += '-'; qn += a; tn += b; tn += '-'; }
		else { qn += a; tn += b; }
	}
	// push gaps to the right, never past the end; every step sees the swaps made before it
	const int len = (int)qn.size();
	for (int i = 0; i < len - 1; ++i) {
		if (tn[i] == '-') {
			int j = i;
			for (;;) {
				const char c = tn.c_str()[++j];            // index len reads the terminator, like the reference
				if (c != '-' || j > len - 1) { if (c == qn[i]) { tn[i] = c; tn[j] = '-'; } break; }
			}
		}
		if (qn[i] == '-') {
			int j = i;
			for (;;) {
				const char c = qn.c_str()[++j];
				if (c != '-' || j > len - 1) { if (c == tn[i]) { qn[i] = c; qn[j] = '-'; } break; }
			}
		}
	}
}

// ------------------------------------------------------------------ C5
struct Vote { char base; uint8_t mat, ins, del; };

void add_votes(const std::string& q, const std::string& s, int soff, Vote* table)
{
	const int n = (int)q.size();
	int i = 0;
	while (i < n) {
		const char a = q[i], b = s[i];
		if (a == '-' && b == '-') { ++i; continue; }
		if (a == b) { ++table[soff].mat; table[soff].base = b; ++soff; ++i; }
		else if (a == '-') { ++table[soff].ins; ++soff; ++i; }
		else {
			int j = i + 1;
			while (j < n && s[j] == '-') ++j;
			++table[soff - 1].del;
			i = j;
		}
	}
}

enum { FMAT = 1, FDEL = 2, FINS = 4, UNDS = 8 };

inline uint8_t classify(const Vote& v)
{
	uint8_t f = 0;
	const int cov = v.mat + v.ins;
	if (v.mat >= cov * 0.8) f |= FMAT;
	if (v.ins >= cov * 0.8) f |= FINS;
	if (!f) f |= UNDS;
	if (v.del >= cov * 0.4) f |= FDEL;
	return f;
}

// ------------------------------------------------------------------ accepted alignments with their read cursor
struct Kept
{
	int soff, send, idx, size;
	std::string q, s;
	// CnsAln::retrieve_aln_subseqs: the cursor only moves forward across calls
	bool slice(int sb, int se, std::string& qs, std::string& ts, int& sb_out)
	{
		if (se <= soff || sb >= send || idx >= size - 1) return false;
		sb_out = std::max(soff, sb);
		qs.clear(); ts.clear();
		while (soff < sb && idx < size - 1) { ++idx; if (s[idx] != '-') ++soff; }
		qs += q[idx]; ts += s[idx];
		while (soff < se && idx < size - 1) { ++idx; if (s[idx] != '-') ++soff; qs += q[idx]; ts += s[idx]; }
		return true;
	}
};

// ------------------------------------------------------------------ C7 mini partial-order graph
class Poa
{
	struct Node { char base = 'N'; int coverage = 0, weight = 0; bool backbone = false; std::vector<int> in, out; };
	struct Edge { int u, v, count = 0; bool visited = false; };
	std::vector<Node> nd;
	std::vector<Edge> ed;
	std::map<int, int> bb;     // _bbMap: node -> backbone node (operator[] semantics: missing -> 0)
	int enter, exit_;

	int add_edge(int u, int v)
	{
		Edge e; e.u = u; e.v = v;
		ed.push_back(e);
		const int id = (int)ed.size() - 1;
		nd[u].out.push_back(id);
		nd[v].in.push_back(id);
		return id;
	}
	int find_edge(int u, int v) const { for (int e : nd[u].out) if (ed[e].v == v) return e; return -1; }
	static void erase_id(std::vector<int>& l, int id) { l.erase(std::find(l.begin(), l.end(), id)); }
	void clear_vertex(int n)
	{
		for (int e : nd[n].out) erase_id(nd[ed[e].v].in, e);
		for (int e : nd[n].in) erase_id(nd[ed[e].u].out, e);
		nd[n].out.clear(); nd[n].in.clear();
	}
	void link(int u, int v)    // addEdge: bump an existing edge or create one
	{
		bool have = false;
		for (int e : nd[v].in) if (ed[e].u == u) { ++ed[e].count; have = true; }
		if (!have) ++ed[add_edge(u, v)].count;
	}
	void merge_in(int n)
	{
		std::map<char, std::vector<int> > groups;
		for (int e : nd[n].in) { const int u = ed[e].u; if (nd[u].out.size() == 1) groups[nd[u].base].push_back(u); }
		for (auto& kv : groups) {
			const std::vector<int> nodes = kv.second;
			if (nodes.size() <= 1) continue;
			const int an = nodes[0];
			const int an_out = nd[an].out.front();
			for (size_t i = 1; i < nodes.size(); ++i) {
				ed[an_out].count += ed[nd[nodes[i]].out.front()].count;
				nd[an].weight += nd[nodes[i]].weight;
			}
			for (size_t i = 1; i < nodes.size(); ++i) {
				const int x = nodes[i];
				for (size_t k = 0; k < nd[x].in.size(); ++k) {
					const int ie = nd[x].in[k];
					const int n1 = ed[ie].u;
					const int e = find_edge(n1, an);
					if (e >= 0) ed[e].count += ed[ie].count;
					else { const int ne = add_edge(n1, an); ed[ne].count = ed[ie].count; ed[ne].visited = ed[ie].visited; }
				}
				clear_vertex(x);
			}
			merge_in(an);
		}
	}
	void merge_out(int n)
	{
		std::map<char, std::vector<int> > groups;
		for (int e : nd[n].out) { const int v = ed[e].v; if (nd[v].in.size() == 1) groups[nd[v].base].push_back(v); }
		for (auto& kv : groups) {
			const std::vector<int> nodes = kv.second;
			if (nodes.size() <= 1) continue;
			const int an = nodes[0];
			const int an_in = nd[an].in.front();
			for (size_t i = 1; i < nodes.size(); ++i) {
				ed[an_in].count += ed[nd[nodes[i]].in.front()].count;
				nd[an].weight += nd[nodes[i]].weight;
			}
			for (size_t i = 1; i < nodes.size(); ++i) {
				const int x = nodes[i];
				for (size_t k = 0; k < nd[x].out.size(); ++k) {
					const int oe = nd[x].out[k];
					const int n2 = ed[oe].v;
					const int e = find_edge(an, n2);
					if (e >= 0) ed[e].count += ed[oe].count;
					else { const int ne = add_edge(an, n2); ed[ne].count = ed[oe].count; ed[ne].visited = ed[oe].visited; }
				}
				clear_vertex(x);
			}
		}
	}

public:
	explicit Poa(int blen) : nd((size_t)blen + 2)
	{
		for (int i = 0; i < blen + 1; ++i) add_edge(i, i + 1);
		enter = 0; exit_ = blen + 1;
		nd[enter].base = '^'; nd[enter].backbone = true;
		for (int i = 1; i <= blen; ++i) { nd[i].backbone = true; nd[i].weight = 1; nd[i].base = 'N'; bb[i] = i; }
		nd[exit_].base = '$'; nd[exit_].backbone = true;
	}
	void add_alignment(const std::string& q, const std::string& t, int start)
	{
		int pos = start, prev = enter;
		for (size_t i = 0; i < q.size(); ++i) {
			const char a = q[i], b = t[i];
			if (a == '-' && b == '-') continue;
			const int cur = pos;
			if (a == b) {
				Node& B = nd[bb[cur]];
				++B.coverage; B.base = b;
				++nd[cur].weight;
				link(prev, cur);
				++pos; prev = cur;
			} else if (a == '-') {
				Node& B = nd[bb[cur]];
				++B.coverage; B.base = b;
				++pos;
			} else {
				nd.emplace_back();
				const int nv = (int)nd.size() - 1;
				nd[nv].base = a; ++nd[nv].weight;
				bb[nv] = pos;
				link(prev, nv);
				prev = nv;
			}
		}
		link(prev, exit_);
	}
	void merge_nodes()
	{
		std::queue<int> seeds;
		seeds.push(enter);
		while (!seeds.empty()) {
			const int u = seeds.front(); seeds.pop();
			merge_in(u);
			merge_out(u);
			for (size_t k = 0; k < nd[u].out.size(); ++k) {
				const int e = nd[u].out[k];
				ed[e].visited = true;
				const int v = ed[e].v;
				int open = 0;
				for (int ie : nd[v].in) if (!ed[ie].visited) ++open;
				if (open == 0) seeds.push(v);
			}
		}
	}
	void consensus(int min_weight, std::string& out)
	{
		for (auto& e : ed) e.visited = false;
		std::map<int, int> best_edge;
		std::map<int, float> score;
		std::queue<int> seeds;
		seeds.push(exit_);
		score[exit_] = 0.0f;
		while (!seeds.empty()) {
			const int n = seeds.front(); seeds.pop();
			bool found = false;
			float best = -FLT_MAX;
			int best_e = -1;
			for (int e : nd[n].out) {
				const int v = ed[e].v;
				const Node& V = nd[v];
				const float s = score[v];
				float ns;
				if (V.backbone && V.weight == 1) ns = s - 10.0f;
				else ns = ed[e].count - nd[bb[v]].coverage * 0.5f + s;
				if (ns > best) { best = ns; best_e = e; found = true; }
			}
			if (found) { score[n] = best; best_edge[n] = best_e; }
			for (int ie : nd[n].in) {
				ed[ie].visited = true;
				const int u = ed[ie].u;
				int open = 0;
				for (int oe : nd[u].out) if (!ed[oe].visited) ++open;
				if (open == 0) seeds.push(u);
			}
		}
		std::string cns;
		int offs = 0, best_offs = 0, length = 0, idx = 0;
		bool met = false;
		for (int p = enter;;) {
			const Node& N = nd[p];
			if (!(N.base == '^' || N.base == '$')) {
				cns += N.base;
				if (!met && N.weight >= min_weight) { offs = idx; met = true; }
				else if (met && N.weight < min_weight) {
					if (idx - offs > length) { best_offs = offs; length = idx - offs; }
					met = false;
				}
				++idx;
			}
			auto it = best_edge.find(p);
			if (it == best_edge.end()) break;
			p = ed[it->second].v;
		}
		if (met && idx - offs > length) { best_offs = offs; length = idx - offs; }
		out.assign(cns.data() + best_offs, (size_t)length);
	}
};

struct Range { int start, end; };

// get_effective_ranges, mecat_correction.cpp:119-153
void effective_ranges(std::vector<Range>& m, std::vector<Range>& e, int read_size, int64_t min_size)
{
	e.clear();
	if (m.empty()) return;
	for (const Range& r : m)
		if (r.start <= 500 && read_size - r.end <= 500) { e.push_back(Range{0, read_size}); return; }
	std::sort(m.begin(), m.end(), [](const Range& a, const Range& b) { return (a.start == b.start) ? (a.end > b.end) : (a.start < b.start); });
	const int nr = (int)m.size();
	int i = 0, left = m[0].start, right;
	while (i < nr) {
		int j = i + 1;
		while (j < nr && m[j].end <= m[i].end) ++j;
		if (j == nr) {
			right = m[i].end;
			if (right - left >= min_size * 0.95) e.push_back(Range{left, right});
			break;
		}
		if (m[i].end - m[j].start < 1000) {
			right = std::min(m[i].end, m[j].start);
			if (right - left >= min_size * 0.95) e.push_back(Range{left, right});
			left = std::max(m[i].end, m[j].start);
		}
		i = j;
	}
}

// output_cns_result, mecat_correction.cpp:156-188
void emit(std::vector<Piece>& out, int64_t id, int64_t beg, int64_t end, const std::string& seq)
{
	const size_t MaxSeq = 60000, Ovlp = 10000, Blk = MaxSeq - Ovlp - 1000;
	const size_t size = seq.size();
	if (size <= MaxSeq) { out.push_back(Piece{id, beg, end, seq}); return; }
	const size_t cutoff = size - Ovlp - 1000;
	size_t L = 0, R;
	do {
		R = L + Blk;
		if (R >= cutoff) R = size;
		Piece p;
		p.id = id; p.beg = (int64_t)L + beg;
		p.end = (R < size && (int64_t)R + beg < end) ? (int64_t)R + beg : end;
		p.seq = seq.substr(L, R - L);
		out.push_back(p);
		L = R - Ovlp;
	} while (R < size);
}

}  // namespace

struct Scratch::Impl
{
	std::vector<Vote> table;
	std::vector<uint8_t> flags;     // cov_stats during the accept loop, then the per-position flags
	std::vector<Kept> kept;
	std::string nq, nt, aux_q, aux_t, cns, seg;
};

Scratch::Scratch() : p(new Impl) {}
Scratch::~Scratch() { delete p; }

void sort_candidates(mecat_candidate* c, int n)   // CmpExtensionCandidateByScore, mecat_correction.cpp:362-370
{
	std::sort(c, c + n, [](const mecat_candidate& a, const mecat_candidate& b) {
		if (a.score != b.score) return a.score > b.score;
		if (a.qid != b.qid) return a.qid < b.qid;
		return a.qext < b.qext;
	});
}

void consensus_one_read(int64_t read_id, int read_size, const mecat_candidate* cand, int ncand, const mecat_align_result* res,
                        const char* qstr, const char* sstr, const Params& P, Scratch& scratch, std::vector<Piece>& out)
{
	Scratch::Impl& S = *scratch.p;
	S.table.assign((size_t)read_size + 1, Vote{'N', 0, 0, 0});
	S.flags.assign((size_t)read_size + 1, 0);
	S.kept.clear();
	uint8_t* cov = S.flags.data();
	const double ratio = P.min_mapping_ratio - 0.02;
	std::set<int> used;
	int added = 0, tried = 0;
	const int max_added = P.tech == 1 ? 100 : 60;      // mecat_correction.cpp:407 / MAX_CNS_OVLPS at :482
	const bool m4 = P.input_type == 1;      // consensus_one_read_m4_*: every overlap handed in, no trial limit, no coverage gate
	for (int i = 0; i < ncand && added < max_added && (m4 || tried < 200); ++i) {
		++tried;
		const mecat_candidate& ec = cand[i];
		if (!m4 && used.count(ec.qid)) continue;
		const mecat_align_result& r = res[i];
		if (!r.ok) continue;
		// check_ovlp_mapping_range (can input; for m4 input only the nanopore variant applies it, mecat_correction.cpp:347)
		const int oq = r.qend - r.qstart, qqs = (int)(ec.qsize * ratio), os = r.send - r.sstart, qss = (int)(ec.ssize * ratio);
		if ((!m4 || P.tech == 1) && !(oq >= qqs || os >= qss)) continue;
		if (!m4) {
			// check_cov_stats: at least 200 positions not yet covered 20 times
			int full = 0;
			for (int k = r.sstart; k < r.send; ++k) if (cov[k] >= 20) ++full;
			if (!(r.send - r.sstart >= full + 200)) continue;
			for (int k = r.sstart; k < r.send; ++k) ++cov[k];
		}
		++added;
		used.insert(ec.qid);
		normalize_gaps(qstr + r.str_offset, sstr + r.str_offset, r.columns, S.nq, S.nt);
		add_votes(S.nq, S.nt, r.sstart, S.table.data());
		S.kept.emplace_back();
		Kept& k = S.kept.back();
		k.soff = r.sstart; k.send = r.send; k.idx = 0; k.size = (int)S.nq.size(); k.q = S.nq; k.s = S.nt;
	}
	std::vector<Range> mr, er;
	for (const Kept& k : S.kept) mr.push_back(Range{k.soff, k.send});
	if (P.tech == 1) er.push_back(Range{0, read_size});        // consensus_one_read_can_nanopore, mecat_correction.cpp:508-509
	else effective_ranges(mr, er, read_size, P.min_size);

	// consensus_worker
	Vote* table = S.table.data();
	uint8_t* flags = S.flags.data();
	for (const Range& rg : er) {
		const int R = rg.end;
		int beg = rg.start;
		while (beg < R) {
			while (beg < R && table[beg].mat + table[beg].ins < P.min_cov) ++beg;
			int end = beg + 1;
			while (end < R && table[end].mat + table[end].ins >= P.min_cov) ++end;
			if (end - beg >= 0.95 * P.min_size) {
				// meap_consensus_one_segment
				const Vote* list = table + beg;
				const int n = end - beg;
				for (int i = 0; i < n; ++i) flags[i] = classify(list[i]);
				std::string& target = S.seg;
				target.clear();
				int i = 0;
				while (i < n && !(flags[i] & FMAT)) ++i;
				while (i < n) {
					target.push_back(list[i].base);
					int j = i + 1;
					while (j < n && !(flags[j] & FMAT)) ++j;
					bool refine = false;
					for (int k = i; k < j; ++k) if ((flags[k] & UNDS) || (flags[k] & FDEL)) { refine = true; break; }
					if (refine) {
						// meap_cns_one_indel
						const int sb = i + beg, se = j + beg;
						Poa g(se - sb + 1);
						int sb_out;
						for (Kept& k : S.kept)
							if (k.slice(sb, se, S.aux_q, S.aux_t, sb_out)) g.add_alignment(S.aux_q, S.aux_t, sb_out - sb + 1);
						g.merge_nodes();
						g.consensus((int)((list[i].mat + list[i].ins) * 0.4), S.cns);
						if (S.cns.size() > 2) target.append(S.cns.data() + 1, S.cns.size() - 2);
					}
					i = j;
				}
				if ((int64_t)target.size() >= P.min_size) emit(out, read_id, beg, end, target);
			}
			beg = end;
		}
	}
}

}  // namespace orccns

// ------------------------------------------------------------------ C entry points (ctypes)
extern "C" {

void orc_cns_sort_candidates(mecat_candidate* c, int n) { orccns::sort_candidates(c, n); }

// M4 input (-i 1): one partition's records in the order partition_m4records wrote them -> the order the reference works in:
// std::sort by sid (build_cns_thrd_data_can, reads_correction_aux.cpp:102), then, for a read with more than `cap` overlaps,
// std::sort by overlap size (CompareOverlapByOverlapSize, mecat_correction.cpp:26-34,261-272; the first `cap` are used).
// This file is compiled without the parallel mode: std::sort is the sequential introsort the reference's parallel-mode
// sort falls back to with one OpenMP thread (the fixtures of tests/golden/*.i1.* were made that way).
void orc_cns_m4_order(mecat_candidate* c, int n, int cap)
{
	std::sort(c, c + n, [](const mecat_candidate& a, const mecat_candidate& b) { return a.sid < b.sid; });
	for (int i = 0; i < n;) {
		int j = i + 1;
		while (j < n && c[j].sid == c[i].sid) ++j;
		if (j - i > cap)
			std::sort(c + i, c + j, [](const mecat_candidate& a, const mecat_candidate& b) {
				return std::max(a.qend - a.qoff, a.send - a.soff) > std::max(b.qend - b.qoff, b.send - b.soff);
			});
		i = j;
	}
}

// One region graph (AlnGraphBoost restatement) on its own: backbone of blen positions, naln gapped alignments
// (q[i], t[i] of equal length, first column at backbone position start[i]).  Returns the length of the consensus
// written to out, or -1 when it does not fit cap.
int orc_poa_consensus(int blen, int naln, const char* const* q, const char* const* t, const int* start, int min_weight, char* out, int cap)
{
	orccns::Poa g(blen);
	for (int i = 0; i < naln; ++i) g.add_alignment(std::string(q[i]), std::string(t[i]), start[i]);
	g.merge_nodes();
	std::string cns;
	g.consensus(min_weight, cns);
	if ((int)cns.size() > cap) return -1;
	memcpy(out, cns.data(), cns.size());
	return (int)cns.size();
}

int orc_cns_consensus(const mecat_candidate* cand, int ncand, const mecat_align_result* res, const char* qstr,
                      const char* sstr, const mecat_cns_params* p, mecat_cns_piece** pieces, size_t* npieces,
                      char** seqs, size_t* seq_bytes)
{
	if (!cand || ncand <= 0 || !res || !p || !pieces || !npieces || !seqs || !seq_bytes) return 1;
	orccns::Params P;
	P.min_mapping_ratio = p->min_mapping_ratio; P.min_align_size = p->min_align_size; P.min_cov = p->min_cov; P.min_size = p->min_size;
	P.tech = p->tech; P.input_type = p->input_type;
	orccns::Scratch scratch;
	std::vector<orccns::Piece> out;
	orccns::consensus_one_read(cand[0].sid, cand[0].ssize, cand, ncand, res, qstr, sstr, P, scratch, out);
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

}  // extern "C"
