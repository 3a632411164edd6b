// mecat_b200/csrc/cns_pipeline.h -- the consensus stage of mecat2cns (rows C3-C7) as a sequence of kernels.
//
// consensus_batch() takes the GetAlignment results of a batch of reads where the extension kernels left
// them (device memory) and produces the corrected pieces.  Every stage is a launch of one functor over
// n independent units (reads, accepted alignments, ambiguous regions, segments) with cns_core.cuh as the
// per-unit body; exclusive scans between the stages size the next stage's arenas exactly, so no stage
// needs a capacity guess or a retry.  The code is written against a small backend interface:
//   * cns.cu provides the CUDA backend (functor -> k_cns<F><<<...>>>, device scans, device memory) and is
//     the only backend in the product library;
//   * tests/cns_host_harness.cpp provides a host backend so that the CPU test-suite can run the same stage
//     sequence and bodies against the reference's golden output without a GPU.
//
//   stage            unit                reference
//   AcceptFn         read (warp)         consensus_one_read_can_pacbio accept loop, mecat_correction.cpp:389-450
//   FlattenFn        read                (layout only)
//   NormVoteFn       accepted alignment  normalize_gaps + meap_add_one_aln + CnsAln cursor index
//   SegmentFn        read (warp)         get_effective_ranges + consensus_worker run search
//   RegionFn<0/1>    read (warp)         identify_one_consensus_item + meap_consensus_one_segment's anchor walk: count, then write, the regions
//   DemandFn         region              node / edge demand of the region's graph
//   PoaFn            region              meap_cns_one_indel: AlnGraphBoost build, merge, best path
//   InteriorLenFn, TargetLenFn               (layout only: exact corrected length of every segment)
//   AssembleFn       segment             meap_consensus_one_segment: every anchor's base to its final place
//   InteriorFn       region              the refined interior behind its anchor
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <string>
#include <vector>

#include "cns_core.cuh"

namespace mbcns {

struct Params { double min_mapping_ratio; int min_align_size; int min_cov; int64_t min_size; int tech = 0; int input_type = 0; };
struct Piece { int64_t id, beg, end; std::string seq; };   // CnsResult, src/common/alignment.h

// kernel slots for the per-stage timers of the backend
enum { ST_ACCEPT = 0, ST_NORMVOTE = 1, ST_SEGMENT = 2, ST_REGION = 3, ST_POA = 4, ST_ASSEMBLE = 5, ST_NUM = 6 };

struct BatchIn
{
	int R = 0;                     // reads of the batch (each with its candidate range, in trial order)
	int64_t T = 0;                 // extension tasks of the batch
	const int32_t* h_first = nullptr;      // [R + 1] task range of read r
	const int32_t* h_read_size = nullptr;  // [R]
	const int64_t* h_read_id = nullptr;    // [R]
	const int32_t* h_tqid = nullptr;       // [T] partner read of task t
	const int32_t* h_tqsize = nullptr;     // [T] its length
	// device side, as left by the extension kernels (align.cu)
	const int32_t* d_info = nullptr;                // [8 T] {ok, qstart, qend, sstart, send, columns, matches, -}
	const char* d_q = nullptr;                      // gapped strings, task t at d_outoff[t]
	const char* d_s = nullptr;
	const unsigned long long* d_outoff = nullptr;   // [T + 1]
};

// ---------------------------------------------------------------------------------------------- functors
struct AcceptFn
{
	const int32_t* first; const int32_t* info; const int32_t* tqid; const int32_t* tqsize; const int32_t* read_size;
	const int64_t* pos_off; uint8_t* cov; double ratio; int32_t* acc; int32_t* nacc; int max_accept; int mode;
	template <class L>
	CNS_HD void operator()(int64_t r, const L& lanes) const      // a warp per read
	{
		const int n = accept_read(lanes, first[r], first[r + 1], info, tqid, tqsize, read_size[r], ratio, cov + pos_off[r], acc + r * MAX_ACCEPT, max_accept, mode);
		if (lanes.leader()) nacc[r] = n;
	}
};

struct FlattenFn      // accepted alignment A = aln_first[r] + k; capacities of its normalised strings and column index
{
	const int32_t* info; const int32_t* acc; const int32_t* nacc; const int64_t* aln_first;
	int32_t* aln_task; int32_t* aln_read; int32_t* cap_norm; int32_t* cap_col;
	CNS_HD void operator()(int64_t r) const
	{
		for (int k = 0; k < nacc[r]; ++k) {
			const int64_t A = aln_first[r] + k;
			const int t = acc[r * MAX_ACCEPT + k];
			const int32_t* o = info + 8 * (int64_t)t;
			aln_task[A] = t; aln_read[A] = (int32_t)r;
			cap_norm[A] = ((2 * o[5] + 2 + 15) & ~15) + 16;     // 16-byte aligned slots, room for the writer's last 8-byte store
			cap_col[A] = o[4] - o[3] + 2;
		}
	}
};

struct NormVoteFn
{
	const int32_t* info; const char* q; const char* s; const unsigned long long* outoff;
	const int32_t* aln_task; const int32_t* aln_read; const int64_t* norm_off; const int64_t* col_off; const int64_t* pos_off;
	char* nq; char* nt; int32_t* colidx; uint32_t* votes; char* base; KeptAln* kept;
	CNS_HD void operator()(int64_t A) const
	{
		const int t = aln_task[A];
		const int32_t* o = info + 8 * (int64_t)t;
		char* a = nq + norm_off[A];
		char* b = nt + norm_off[A];
		const int64_t po = pos_off[aln_read[A]];
		int tend;
		const int len = normalize_vote_index(q + outoff[t], s + outoff[t], o[5], o[3], a, b, votes + po, base + po, colidx + col_off[A], &tend);
		KeptAln K;
		K.q = a; K.s = b; K.colidx = colidx + col_off[A];
		K.size = len; K.soff = o[3]; K.send = o[4]; K.tend = tend;
		kept[A] = K;
	}
};

struct SegmentFn
{
	const int32_t* nacc; const int64_t* aln_first; const KeptAln* kept; const int32_t* read_size; const int64_t* pos_off;
	const uint32_t* votes; int min_cov; double size95; const int64_t* seg_slot; int32_t* segs; int32_t* nseg; int whole_read;
	template <class L>
	CNS_HD void operator()(int64_t r, const L& lanes) const      // a warp per read
	{
		// the ranges live in the warp's scratch (shared memory on the device), not in 32 private copies: 1.6 KB of stack per
		// thread made the driver re-size its local-memory pool at the first launch of this kernel (0.5-0.7 s, now and then)
		Range* m = (Range*)lanes.scratch();
		Range* e = m + MAX_ACCEPT;
		int* pne = (int*)(e + MAX_ACCEPT);
		const int n = nacc[r];
		if (lanes.leader()) {
			for (int k = 0; k < n; ++k) { m[k].start = kept[aln_first[r] + k].soff; m[k].end = kept[aln_first[r] + k].send; }
			if (whole_read) { e[0].start = 0; e[0].end = read_size[r]; *pne = 1; }      // nanopore: mecat_correction.cpp:508-509
			else *pne = effective_ranges(m, n, e, read_size[r], size95);
		}
		lanes.sync();
		const int ne = *pne;
		const int cap = (int)(seg_slot[r + 1] - seg_slot[r]);
		const int ns = find_segments(lanes, e, ne, votes + pos_off[r], min_cov, size95, segs + 2 * seg_slot[r], cap);
		if (lanes.leader()) nseg[r] = ns <= cap ? ns : -1;       // -1: capacity formula violated (reported by the host as an error)
	}
};

// Both passes over a read's segments, a warp per read.  FILL = false: classify the positions (flags) and count the
// regions; FILL = true: write the regions, each segment's first region and its number of anchors.
template <bool FILL>
struct RegionFn
{
	const int32_t* nseg; const int64_t* seg_slot; const int32_t* segs; const int64_t* pos_off; const uint32_t* votes;
	uint8_t* flags; int32_t* nreg; const int64_t* reg_first; const int64_t* seg_first; Region* regions; int64_t* seg_reg;
	int32_t* seg_anchors;
	template <class L>
	CNS_HD void operator()(int64_t r, const L& lanes) const
	{
		const int64_t po = pos_off[r];
		int64_t g = FILL ? reg_first[r] : 0;
		int last_se_abs = -1;
		for (int k = 0; k < nseg[r]; ++k) {
			const int beg = segs[2 * (seg_slot[r] + k)], end = segs[2 * (seg_slot[r] + k) + 1];
			const int n = end - beg;
			uint8_t* f = flags + po + beg;
			const uint32_t* v = votes + po + beg;
			const int64_t S = FILL ? seg_first[r] + k : 0;
			if (!FILL) lanes.each([&](int l) { for (int i = l; i < n; i += L::count) f[i] = classify(v[i]); });
			AnchorCarry c;
			c.last_se_abs = last_se_abs;
			for (int base = 0; base < n; base += L::count)
				anchor_chunk(lanes, base, n, beg, [&](int p) { return (int)f[p]; }, c,
				             [&](int, int pos, int rank, int prevpos, bool closes, int ordinal, int prev_se) {
					if (!FILL || !closes) return;
					Region G;
					G.read = (int32_t)r; G.sb = prevpos + beg; G.se = pos + beg; G.prev_se = prev_se;
					G.min_weight = (int)((double)(vote_mat(v[prevpos]) + vote_ins(v[prevpos])) * 0.4);
					G.seg = (int32_t)S; G.rank = rank - 1;
					regions[g + ordinal] = G;
				});
			if (c.last_anchor >= 0 && c.pending) {           // the interval from the last anchor to the segment end
				if (FILL && lanes.leader()) {
					Region G;
					G.read = (int32_t)r; G.sb = c.last_anchor + beg; G.se = end; G.prev_se = c.last_se_abs;
					G.min_weight = (int)((double)(vote_mat(v[c.last_anchor]) + vote_ins(v[c.last_anchor])) * 0.4);
					G.seg = (int32_t)S; G.rank = c.nanchors - 1;
					regions[g + c.nregions] = G;
				}
				c.last_se_abs = end;
				++c.nregions;
			}
			if (FILL && lanes.leader()) { seg_reg[S] = g; seg_anchors[S] = c.nanchors; }
			g += c.nregions;
			last_se_abs = c.last_se_abs;
		}
		if (!FILL && lanes.leader()) nreg[r] = (int32_t)g;
	}
};

struct DemandFn
{
	const Region* regions; const int32_t* nacc; const int64_t* aln_first; const KeptAln* kept; int32_t* dn; int32_t* de;
	CNS_HD void operator()(int64_t g) const
	{
		const Region G = regions[g];
		int n, e;
		region_demand(kept + aln_first[G.read], nacc[G.read], G.sb, G.se, G.prev_se, n, e);
		dn[g] = n; de[g] = e;
	}
};

struct PoaFn
{
	const Region* regions; const int32_t* nacc; const int64_t* aln_first; const KeptAln* kept;
	const int64_t* node_off; const int64_t* edge0_off;
	int64_t g_base, n_base, e_base;      // first region of this wave and its offsets: the arena holds one wave
	char* arena; char* gout; int32_t* goff; int32_t* glen; int32_t* gerr;
	// capacities of the wave's k-th graph and the place of its int32-layout arena
	CNS_HD void shape(int64_t k, int& ncap, int& e0) const
	{
		const int64_t g = g_base + k;
		ncap = (int)(node_off[g + 1] - node_off[g]);
		e0 = (int)(edge0_off[g + 1] - edge0_off[g]);
	}
	CNS_HD char* wide_arena(int64_t k) const
	{
		const int64_t g = g_base + k;
		return arena + 112 * (node_off[g] - n_base) + 32 * (edge0_off[g] - e_base) + 320 * k;
	}
	// index width a graph of this shape needs: 1 (int8_t), 2 (int16_t) or 4 (int32_t) bytes
	CNS_HD static int width(int ncap, int e0)
	{
		const int64_t m = ncap > poa_edge_cap(ncap, e0) ? ncap : poa_edge_cap(ncap, e0);
		return m < POA_TINY_LIMIT ? 1 : m < POA_SMALL_LIMIT ? 2 : 4;
	}
	CNS_HD static int64_t bytes_for(int w, int ncap, int e0)
	{
		return w == 1 ? poa_arena_bytes<int8_t>(ncap, e0) : w == 2 ? poa_arena_bytes<int16_t>(ncap, e0) : poa_arena_bytes<int32_t>(ncap, e0);
	}
	CNS_HD void solve_width(int w, int64_t k, char* scratch) const
	{
		if (w == 1) solve<int8_t>(k, scratch); else if (w == 2) solve<int16_t>(k, scratch); else solve<int32_t>(k, scratch);
	}
	// the graph of the wave's k-th region with index type I in the given scratch
	template <class I>
	CNS_HD void solve(int64_t k, char* scratch) const
	{
		const int64_t g = g_base + k;
		const Region G = regions[g];
		int ncap, e0, off, len;
		shape(k, ncap, e0);
		gerr[g] = region_consensus<I>(kept + aln_first[G.read], nacc[G.read], G.sb, G.se, G.prev_se, G.min_weight, scratch, ncap,
		                              (int)poa_edge_cap(ncap, e0), gout + node_off[g], off, len);
		goff[g] = off; glen[g] = len;
	}
	// a thread per region entirely in its global arena (the GPU kernel prefers shared memory, cns.cu: k_cns_poa);
	// min_width lets tests force wider indices than the graph needs
	CNS_HD void operator()(int64_t k, int min_width = 1) const
	{
		int ncap, e0;
		shape(k, ncap, e0);
		const int w = width(ncap, e0);
		solve_width(w > min_width ? w : min_width, k, wide_arena(k));
	}
};

struct InteriorLenFn   // bases a region contributes between its two anchors: the best path without its first and last node
{
	const int32_t* glen; int32_t* ilen;
	CNS_HD void operator()(int64_t g) const { ilen[g] = glen[g] > 2 ? glen[g] - 2 : 0; }
};

struct TargetLenFn     // exact corrected length of a segment: its anchors plus the interiors of its regions
{
	const int32_t* seg_read; const int64_t* seg_reg; const int64_t* reg_first; const int64_t* seg_first; const int32_t* nseg;
	const int32_t* seg_anchors; const int64_t* isum; int32_t* tlen;
	CNS_HD int64_t reg_end(int64_t S) const      // one past the last region of segment S
	{
		const int r = seg_read[S];
		return (S + 1 < seg_first[r] + nseg[r]) ? seg_reg[S + 1] : reg_first[r + 1];
	}
	CNS_HD void operator()(int64_t S) const { tlen[S] = (int32_t)(seg_anchors[S] + (isum[reg_end(S)] - isum[seg_reg[S]])); }
};

struct AssembleFn      // a warp per segment: every anchor's base goes to its final place (rank + interiors before it)
{
	const int32_t* seg_read; const int32_t* seg_beg; const int32_t* seg_end; const int64_t* seg_reg; const int64_t* pos_off;
	const uint8_t* flags; const char* base; const int64_t* isum; const int64_t* tgt_off; char* target;
	template <class L>
	CNS_HD void operator()(int64_t S, const L& lanes) const
	{
		const int beg = seg_beg[S], n = seg_end[S] - beg;
		const int64_t po = pos_off[seg_read[S]] + beg;
		const uint8_t* f = flags + po;
		char* out = target + tgt_off[S];
		const int64_t g0 = seg_reg[S];
		AnchorCarry c;
		for (int b0 = 0; b0 < n; b0 += L::count)
			anchor_chunk(lanes, b0, n, beg, [&](int p) { return (int)f[p]; }, c,
			             [&](int, int pos, int rank, int, bool closes, int ordinal, int) {
				const int64_t g = g0 + ordinal + (closes ? 1 : 0);       // regions whose interior precedes this anchor
				out[rank + (isum[g] - isum[g0])] = base[po + pos];
			});
	}
};

struct InteriorFn      // a thread per region: its interior goes right behind its anchor
{
	const Region* regions; const int64_t* seg_reg; const int64_t* isum; const int64_t* tgt_off; const int64_t* node_off;
	const char* gout; const int32_t* goff; const int32_t* ilen; char* target;
	CNS_HD void operator()(int64_t g) const
	{
		const int l = ilen[g];
		if (l <= 0) return;
		const Region G = regions[g];
		char* dst = target + tgt_off[G.seg] + G.rank + 1 + (isum[g] - isum[seg_reg[G.seg]]);
		const char* src = gout + node_off[g] + goff[g] + 1;
		for (int k = 0; k < l; ++k) dst[k] = src[k];
	}
};

struct ErrCountFn { const int32_t* err; uint32_t* n; CNS_HD void operator()(int64_t g) const { if (err[g]) fetch_add(n, 1u); } };

struct SegFlattenFn
{
	const int32_t* nseg; const int64_t* seg_slot; const int32_t* segs; const int64_t* seg_first;
	int32_t* seg_read; int32_t* seg_beg; int32_t* seg_end;
	CNS_HD void operator()(int64_t r) const
	{
		for (int k = 0; k < nseg[r]; ++k) {
			const int64_t S = seg_first[r] + k;
			seg_read[S] = (int32_t)r; seg_beg[S] = segs[2 * (seg_slot[r] + k)]; seg_end[S] = segs[2 * (seg_slot[r] + k) + 1];
		}
	}
};

// output_cns_result, mecat_correction.cpp:156-188 (host: splits pieces longer than 60 000 bases).  Sink: add(id, beg, end, seq, len).
template <class Sink>
inline void emit_piece(Sink& out, int64_t id, int64_t beg, int64_t end, const char* seq, size_t size)
{
	const size_t MaxSeq = 60000, Ovlp = 10000, Blk = MaxSeq - Ovlp - 1000;
	if (size <= MaxSeq) { out.add(id, beg, end, seq, size); return; }
	const size_t cutoff = size - Ovlp - 1000;
	size_t L = 0, R;
	do {
		R = L + Blk;
		if (R >= cutoff) R = size;
		const int64_t pb = (int64_t)L + beg;
		const int64_t pe = (R < size && (int64_t)R + beg < end) ? (int64_t)R + beg : end;
		out.add(id, pb, pe, seq + L, R - L);
		L = R - Ovlp;
	} while (R < size);
}

struct PieceVector      // sink that keeps every piece as its own string (tests)
{
	std::vector<Piece> pieces;
	void add(int64_t id, int64_t beg, int64_t end, const char* seq, size_t len) { pieces.push_back(Piece{id, beg, end, std::string(seq, len)}); }
};

// ---------------------------------------------------------------------------------------------- pipeline
// Backend B:
//   template <class T> T* alloc(size_t n)            device array, freed by end_batch(); nullptr + error on failure
//   bool upload(T* d, const T* h, size_t n), bool download(T* h, const T* d, size_t n), bool fill(void* d, int byte, size_t bytes)
//   template <class F> bool launch(int64_t n, const F& f, int stage)        f(i), one thread per unit
//   template <class F> bool launch_warp(int64_t n, const F& f, int stage)   f(i, lanes), one warp per unit
//   bool launch_graphs(int64_t n, const PoaFn& f, int stage)               the region graphs of one wave
//   bool scan(const int32_t* d_in, int64_t* d_out, int64_t n, int64_t* total)   d_out[0..n] exclusive prefix, total on the host
//   const char* download_staged(const char* d, size_t n)   bulk result bytes in a host buffer owned by the backend
//   bool release(void* d)                            early free of an alloc() block (stream ordered)
//   int64_t poa_budget_bytes()                       scratch budget of one wave of region graphs
//   void fail(const char* msg), void end_batch()
template <class B, class Sink>
int consensus_batch(B& be, const BatchIn& in, const Params& P, Sink& out)
{
	struct Guard { B& b; ~Guard() { b.end_batch(); } } guard{be};
	const int R = in.R;
	const int64_t T = in.T;
	if (R == 0) return 0;
	const double ratio = P.min_mapping_ratio - 0.02;
	const double size95 = 0.95 * (double)P.min_size;
	const bool debug = getenv("MECAT_CNS_DEBUG") != nullptr;
	struct timespec ts0;
	clock_gettime(CLOCK_MONOTONIC, &ts0);
	auto lap = [&](const char* what) {       // debug only: wall time since the previous lap (stages end in a scan or download, i.e. a sync)
		if (!debug) return;
		struct timespec t1;
		clock_gettime(CLOCK_MONOTONIC, &t1);
		fprintf(stderr, "[cns batch] %-28s %8.1f ms\n", what, (t1.tv_sec - ts0.tv_sec) * 1e3 + (t1.tv_nsec - ts0.tv_nsec) * 1e-6);
		ts0 = t1;
	};

	// per-read position arenas and segment slots
	std::vector<int64_t> h_pos((size_t)R + 1, 0), h_slot((size_t)R + 1, 0);
	for (int r = 0; r < R; ++r) {
		h_pos[r + 1] = h_pos[r] + in.h_read_size[r] + 1;
		const double unit = size95 > 1.0 ? size95 : 1.0;
		h_slot[r + 1] = h_slot[r] + (int64_t)((double)in.h_read_size[r] / unit) + MAX_ACCEPT + 1;
	}
	const int64_t POS = h_pos[R];
#define CNS_TRY(x) do { if (!(x)) return 1; } while (0)
#define CNS_ALLOC(var, type, count) type* var = be.template alloc<type>((size_t)(count)); if (!var) return 1
	CNS_ALLOC(d_first, int32_t, R + 1);
	CNS_ALLOC(d_rsize, int32_t, R);
	CNS_ALLOC(d_tqid, int32_t, T);
	CNS_ALLOC(d_tqsize, int32_t, T);
	CNS_ALLOC(d_pos, int64_t, R + 1);
	CNS_ALLOC(d_slot, int64_t, R + 1);
	CNS_ALLOC(d_cov, uint8_t, POS);
	CNS_ALLOC(d_votes, uint32_t, POS);
	CNS_ALLOC(d_base, char, POS);
	CNS_ALLOC(d_acc, int32_t, (int64_t)R * MAX_ACCEPT);
	CNS_ALLOC(d_nacc, int32_t, R);
	CNS_ALLOC(d_alnfirst, int64_t, R + 1);
	CNS_TRY(be.upload(d_first, in.h_first, (size_t)R + 1));
	CNS_TRY(be.upload(d_rsize, in.h_read_size, (size_t)R));
	CNS_TRY(be.upload(d_tqid, in.h_tqid, (size_t)T));
	CNS_TRY(be.upload(d_tqsize, in.h_tqsize, (size_t)T));
	CNS_TRY(be.upload(d_pos, h_pos.data(), (size_t)R + 1));
	CNS_TRY(be.upload(d_slot, h_slot.data(), (size_t)R + 1));
	CNS_TRY(be.fill(d_cov, 0, (size_t)POS));
	CNS_TRY(be.fill(d_votes, 0, (size_t)POS * 4));
	CNS_TRY(be.fill(d_base, 'N', (size_t)POS));

	// C3: which alignments vote
	CNS_TRY(be.launch_warp(R, AcceptFn{d_first, in.d_info, d_tqid, d_tqsize, d_rsize, d_pos, d_cov, ratio, d_acc, d_nacc, P.tech == 1 ? MAX_ACCEPT : MAX_ACCEPT_PACBIO, P.input_type == 1 ? (P.tech == 1 ? 2 : 1) : 0}, ST_ACCEPT));
	int64_t NA = 0;
	CNS_TRY(be.scan(d_nacc, d_alnfirst, R, &NA));
	lap("setup + accept");
	if (NA == 0) return 0;
	CNS_ALLOC(d_alntask, int32_t, NA);
	CNS_ALLOC(d_alnread, int32_t, NA);
	CNS_ALLOC(d_capnorm, int32_t, NA);
	CNS_ALLOC(d_capcol, int32_t, NA);
	CNS_ALLOC(d_normoff, int64_t, NA + 1);
	CNS_ALLOC(d_coloff, int64_t, NA + 1);
	CNS_ALLOC(d_kept, KeptAln, NA);
	CNS_TRY(be.launch(R, FlattenFn{in.d_info, d_acc, d_nacc, d_alnfirst, d_alntask, d_alnread, d_capnorm, d_capcol}, ST_ACCEPT));
	int64_t NORM = 0, COL = 0;
	CNS_TRY(be.scan(d_capnorm, d_normoff, NA, &NORM));
	CNS_TRY(be.scan(d_capcol, d_coloff, NA, &COL));
	CNS_ALLOC(d_nq, char, NORM);
	CNS_ALLOC(d_nt, char, NORM);
	CNS_ALLOC(d_colidx, int32_t, COL);

	// C4 + C5: normalise, vote, index the columns.  (In read order: the 32 alignments of a warp then belong to one or two
	// templates and share their vote / base lines.  Launching them in order of length instead -- so that a warp's threads
	// finish together -- was measured at 593 ms against 221 ms.)
	CNS_TRY(be.launch(NA, NormVoteFn{in.d_info, in.d_q, in.d_s, in.d_outoff, d_alntask, d_alnread, d_normoff, d_coloff, d_pos,
	                                 d_nq, d_nt, d_colidx, d_votes, d_base, d_kept}, ST_NORMVOTE));

	lap("flatten + norm/vote launch");
	// C6: covered runs of each read
	CNS_ALLOC(d_segs, int32_t, 2 * h_slot[R]);
	CNS_ALLOC(d_nseg, int32_t, R);
	CNS_ALLOC(d_segfirst, int64_t, R + 1);
	CNS_TRY(be.launch_warp(R, SegmentFn{d_nacc, d_alnfirst, d_kept, d_rsize, d_pos, d_votes, P.min_cov, size95, d_slot, d_segs, d_nseg, P.tech == 1 ? 1 : 0}, ST_SEGMENT));
	std::vector<int32_t> h_nseg((size_t)R);
	CNS_TRY(be.download(h_nseg.data(), d_nseg, (size_t)R));
	for (int r = 0; r < R; ++r) if (h_nseg[r] < 0) { be.fail("cns: segment slots of a read overflowed"); return 1; }
	int64_t NS = 0;
	CNS_TRY(be.scan(d_nseg, d_segfirst, R, &NS));
	lap("norm/vote + segments");
	if (NS == 0) return 0;

	// C6: anchors and ambiguous regions of every segment
	uint8_t* d_flags = d_cov;      // the accept loop is done with its coverage bytes
	CNS_ALLOC(d_nreg, int32_t, R);
	CNS_ALLOC(d_regfirst, int64_t, R + 1);
	CNS_ALLOC(d_segreg, int64_t, NS + 1);
	CNS_ALLOC(d_segread, int32_t, NS);
	CNS_ALLOC(d_segbeg, int32_t, NS);
	CNS_ALLOC(d_segend, int32_t, NS);
	CNS_TRY(be.launch(R, SegFlattenFn{d_nseg, d_slot, d_segs, d_segfirst, d_segread, d_segbeg, d_segend}, ST_SEGMENT));
	CNS_ALLOC(d_seganchors, int32_t, NS);
	CNS_TRY(be.launch_warp(R, RegionFn<false>{d_nseg, d_slot, d_segs, d_pos, d_votes, d_flags, d_nreg, nullptr, nullptr, nullptr, nullptr, nullptr}, ST_REGION));
	int64_t NG = 0;
	CNS_TRY(be.scan(d_nreg, d_regfirst, R, &NG));
	CNS_ALLOC(d_regions, Region, NG + 1);
	CNS_TRY(be.launch_warp(R, RegionFn<true>{d_nseg, d_slot, d_segs, d_pos, d_votes, d_flags, d_nreg, d_regfirst, d_segfirst, d_regions, d_segreg,
	                                         d_seganchors}, ST_REGION));

	// C7: one graph per region, each in an arena of exactly its demand
	CNS_ALLOC(d_dn, int32_t, NG + 1);
	CNS_ALLOC(d_de, int32_t, NG + 1);
	CNS_ALLOC(d_nodeoff, int64_t, NG + 1);
	CNS_ALLOC(d_edgeoff, int64_t, NG + 1);
	CNS_ALLOC(d_goff, int32_t, NG + 1);
	CNS_ALLOC(d_glen, int32_t, NG + 1);
	CNS_ALLOC(d_gerr, int32_t, NG + 1);
	int64_t NODES = 0, EDGES0 = 0;
	if (NG) CNS_TRY(be.launch(NG, DemandFn{d_regions, d_nacc, d_alnfirst, d_kept, d_dn, d_de}, ST_POA));
	CNS_TRY(be.scan(d_dn, d_nodeoff, NG, &NODES));
	CNS_TRY(be.scan(d_de, d_edgeoff, NG, &EDGES0));
	if (getenv("MECAT_CNS_DEBUG"))
		fprintf(stderr, "[cns batch] reads %d tasks %lld accepted %lld segments %lld regions %lld nodes %lld edges0 %lld norm bytes %lld\n", R,
		        (long long)T, (long long)NA, (long long)NS, (long long)NG, (long long)NODES, (long long)EDGES0, (long long)NORM);
	lap("regions + demand");
	CNS_ALLOC(d_gout, char, NODES);
	{
		// The graphs run in waves whose global scratch fits the backend's budget.  Scratch bytes of regions [a, b) =
		// cost(b) - cost(a) with cost(g) = 112 nodes_before(g) + 32 edges0_before(g) + 320 g (poa_arena_bytes<int32_t>);
		// wave ends are found by bisection on the device-resident prefix sums.
		const int64_t budget = be.poa_budget_bytes();
		auto cost_at = [&](int64_t g, int64_t& n, int64_t& e) -> bool {
			if (g == NG) { n = NODES; e = EDGES0; return true; }
			return be.download(&n, d_nodeoff + g, 1) && be.download(&e, d_edgeoff + g, 1);
		};
		int64_t a = 0, na = 0, ea = 0;
		while (a < NG) {
			int64_t b = NG, nb = NODES, eb = EDGES0;
			if (112 * (nb - na) + 32 * (eb - ea) + 320 * (b - a) > budget) {
				int64_t lo = a + 1, hi = NG;                 // largest b in [a + 1, NG] whose wave fits; a + 1 always accepted
				while (lo < hi) {
					const int64_t mid = lo + (hi - lo + 1) / 2;
					int64_t nm, em;
					CNS_TRY(cost_at(mid, nm, em));
					if (112 * (nm - na) + 32 * (em - ea) + 320 * (mid - a) <= budget) lo = mid; else hi = mid - 1;
				}
				b = lo;
				CNS_TRY(cost_at(b, nb, eb));
			}
			char* d_arena = be.template alloc<char>((size_t)(112 * (nb - na) + 32 * (eb - ea) + 320 * (b - a)));
			if (!d_arena) return 1;
			CNS_TRY(be.launch_graphs(b - a, PoaFn{d_regions, d_nacc, d_alnfirst, d_kept, d_nodeoff, d_edgeoff, a, na, ea, d_arena, d_gout,
			                                      d_goff, d_glen, d_gerr}, ST_POA));
			CNS_TRY(be.release(d_arena));
			a = b; na = nb; ea = eb;
		}
	}

	lap("graph waves (launch)");
	// corrected bases of every segment: exact lengths first, then anchors and interiors straight to their final places
	CNS_ALLOC(d_ilen, int32_t, NG + 1);
	CNS_ALLOC(d_isum, int64_t, NG + 1);
	CNS_ALLOC(d_tgtoff, int64_t, NS + 1);
	CNS_ALLOC(d_tlen, int32_t, NS);
	int64_t ISUM = 0, TGT = 0;
	if (NG) CNS_TRY(be.launch(NG, InteriorLenFn{d_glen, d_ilen}, ST_ASSEMBLE));
	CNS_TRY(be.scan(d_ilen, d_isum, NG, &ISUM));
	CNS_TRY(be.launch(NS, TargetLenFn{d_segread, d_segreg, d_regfirst, d_segfirst, d_nseg, d_seganchors, d_isum, d_tlen}, ST_ASSEMBLE));
	CNS_TRY(be.scan(d_tlen, d_tgtoff, NS, &TGT));
	CNS_ALLOC(d_target, char, TGT);
	CNS_TRY(be.launch_warp(NS, AssembleFn{d_segread, d_segbeg, d_segend, d_segreg, d_pos, d_flags, d_base, d_isum, d_tgtoff, d_target}, ST_ASSEMBLE));
	if (NG) CNS_TRY(be.launch(NG, InteriorFn{d_regions, d_segreg, d_isum, d_tgtoff, d_nodeoff, d_gout, d_goff, d_ilen, d_target}, ST_ASSEMBLE));

	lap("graphs + assemble");
	// results to the host
	std::vector<int32_t> h_segread((size_t)NS), h_segbeg((size_t)NS), h_segend((size_t)NS), h_tlen((size_t)NS);
	std::vector<int64_t> h_tgtoff((size_t)NS + 1);
	CNS_TRY(be.download(h_segread.data(), d_segread, (size_t)NS));
	CNS_TRY(be.download(h_segbeg.data(), d_segbeg, (size_t)NS));
	CNS_TRY(be.download(h_segend.data(), d_segend, (size_t)NS));
	CNS_TRY(be.download(h_tlen.data(), d_tlen, (size_t)NS));
	CNS_TRY(be.download(h_tgtoff.data(), d_tgtoff, (size_t)NS + 1));
	const char* h_target = be.download_staged(d_target, (size_t)TGT);      // valid until the batch ends
	if (!h_target) return 1;
	// error codes of the region graphs: counted on the device (the codes of 11 M regions were a 45 MB copy per batch); the
	// array itself only comes to the host when something failed
	uint32_t nerr = 0;
	if (NG) {
		CNS_ALLOC(d_nerr, uint32_t, 1);
		CNS_TRY(be.fill(d_nerr, 0, sizeof(uint32_t)));
		CNS_TRY(be.launch(NG, ErrCountFn{d_gerr, d_nerr}, ST_ASSEMBLE));
		CNS_TRY(be.download(&nerr, d_nerr, 1));
	}
	if (nerr) {
		std::vector<int32_t> h_gerr((size_t)NG);
		CNS_TRY(be.download(h_gerr.data(), d_gerr, (size_t)NG));
		for (int64_t g = 0; g < NG; ++g)
			if (h_gerr[g]) {
				char msg[128];
				snprintf(msg, sizeof msg, "cns: region graph %lld ran out of scratch (code %d)", (long long)g, h_gerr[g]);
				be.fail(msg);
				return 1;
			}
	}
	lap("downloads");
	for (int64_t S = 0; S < NS; ++S)
		if ((int64_t)h_tlen[S] >= P.min_size)
			emit_piece(out, in.h_read_id[h_segread[S]], h_segbeg[S], h_segend[S], h_target + h_tgtoff[S], (size_t)h_tlen[S]);
	lap("emit pieces");
#undef CNS_TRY
#undef CNS_ALLOC
	return 0;
}

}  // namespace mbcns
