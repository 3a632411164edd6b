// mecat_b200/csrc/cns_core.cuh -- bodies of the consensus kernels of mecat2cns (rows C3-C7).
//
// Every function here is the work of ONE GPU thread on one unit (a read, an accepted alignment, a
// template segment, an ambiguous region); cns.cu launches them as kernels.  The bodies are plain
// integer / byte code over raw arrays and carry CNS_HD (= __host__ __device__ under nvcc), so the CPU
// test-suite can also compile this header with g++ and execute the very same statements without a GPU
// (tests/cns_host_harness.cpp); the product library instantiates the device side only.
//
// Semantics follow the reference (src/mecat2cns):
//   consensus_one_read_can_pacbio, check_ovlp_mapping_range, check_cov_stats   mecat_correction.cpp:191-200,373-450
//   normalize_gaps                                                             reads_correction_aux.cpp:3-79
//   meap_add_one_aln, identify_one_consensus_item, meap_consensus_one_segment  mecat_correction.cpp:15-108
//   get_effective_ranges, consensus_worker                                     mecat_correction.cpp:119-239
//   CnsAln::retrieve_aln_subseqs                                               reads_correction_aux.h:47-68
//   AlnGraphBoost (mini partial-order graph of one ambiguous region)           MECAT_AlnGraphBoost.C:76-592
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CNS_HD __host__ __device__
#else
#define CNS_HD
#endif

namespace mbcns {

constexpr int MAX_ACCEPT = 100;     // capacity: MAX_CNS_OVLPS, reads_correction_aux.h:32 (the nanopore cap, mecat_correction.cpp:482)
constexpr int MAX_ACCEPT_PACBIO = 60;   // max_added of consensus_one_read_can_pacbio, mecat_correction.cpp:407
constexpr int MAX_TRIED = 200;      // mecat_correction.cpp:410 (MAX_EXAMINED_OVLPS)
constexpr int COV_FULL = 20;        // check_cov_stats, mecat_correction.cpp:373-386
constexpr int COV_NEED = 200;

enum { FMAT = 1, FDEL = 2, FINS = 4, UNDS = 8 };

// vote word of one template position: mat | ins << 8 | del << 16 (CnsTableItem's three uint8 counters; at most
// 60 alignments vote, so no byte ever carries into its neighbour)
CNS_HD inline int vote_mat(uint32_t w) { return (int)(w & 255u); }
CNS_HD inline int vote_ins(uint32_t w) { return (int)((w >> 8) & 255u); }
CNS_HD inline int vote_del(uint32_t w) { return (int)((w >> 16) & 255u); }

CNS_HD inline void vote_add(uint32_t* p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
	atomicAdd(p, v);
#else
	*p += v;
#endif
}

CNS_HD inline uint32_t fetch_add(uint32_t* p, uint32_t v)
{
#if defined(__CUDA_ARCH__)
	return atomicAdd(p, v);
#else
	const uint32_t old = *p;
	*p += v;
	return old;
#endif
}

// identify_one_consensus_item, mecat_correction.cpp:15-24 (int compared with a double product)
CNS_HD inline uint8_t classify(uint32_t w)
{
	const int mat = vote_mat(w), ins = vote_ins(w), del = vote_del(w);
	const int cov = mat + ins;
	uint8_t f = 0;
	if ((double)mat >= (double)cov * 0.8) f |= FMAT;
	if ((double)ins >= (double)cov * 0.8) f |= FINS;
	if (!f) f |= UNDS;
	if ((double)del >= (double)cov * 0.4) f |= FDEL;
	return f;
}

// ------------------------------------------------------------------------------------------ C3
CNS_HD inline int popcount64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
	return __popcll(x);
#else
	return __builtin_popcountll(x);
#endif
}

// Some units are worked on by a whole warp.  Their bodies are written once against a "lanes" object with 32 lanes:
//   each(f)      f(lane) on every lane; memory effects are visible to all lanes afterwards
//   sum(f)       sum of f(lane) over the lanes, the same value on every lane
//   ballot(f)    bit l = f(l); ballot2(f, m0, m1) for two predicates at once (f returns bit 0 | bit 1 << 1)
//   leader()     true on exactly one lane (single writes)
// Everything outside the callbacks is uniform (every lane computes the same scalars).  cns.cu's WarpLanes maps this
// to warp intrinsics; EmuLanes below runs the 32 lanes in a loop and is what the host harness uses, so the mask
// arithmetic the GPU executes is the arithmetic the CPU tests check.
struct EmuLanes
{
	static constexpr int count = 32;
	template <class F> void each(F&& f) const { for (int l = 0; l < count; ++l) f(l); }
	template <class F> int sum(F&& f) const { int s = 0; for (int l = 0; l < count; ++l) s += f(l); return s; }
	template <class F> uint32_t ballot(F&& f) const { uint32_t m = 0; for (int l = 0; l < count; ++l) if (f(l)) m |= 1u << l; return m; }
	template <class F> void ballot2(F&& f, uint32_t& m0, uint32_t& m1) const
	{
		m0 = m1 = 0;
		for (int l = 0; l < count; ++l) { const int v = f(l); if (v & 1) m0 |= 1u << l; if (v & 2) m1 |= 1u << l; }
	}
	bool leader() const { return true; }
	void sync() const {}
	void* scratch() const { return (void*)buf; }      // LANES_SCRATCH bytes private to the "warp"
	mutable uint64_t buf[256];
};
constexpr int LANES_SCRATCH = 2048;

CNS_HD inline int popcount32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
	return __popc(x);
#else
	return __builtin_popcount(x);
#endif
}
CNS_HD inline int lowest_bit(uint32_t x)       // x != 0
{
#if defined(__CUDA_ARCH__)
	return __ffs((int)x) - 1;
#else
	return __builtin_ctz(x);
#endif
}
CNS_HD inline int highest_bit(uint32_t x)      // x != 0
{
#if defined(__CUDA_ARCH__)
	return 31 - __clz((int)x);
#else
	return 31 - __builtin_clz(x);
#endif
}
CNS_HD inline uint32_t bit_range(int lo, int hi)   // bits [lo, hi), 0 <= lo <= hi <= 32
{
	return (uint32_t)(((1ull << hi) - 1ull) & ~((1ull << lo) - 1ull));
}

// first position p in [from, to) with pred(p), else to; 32 positions per step
template <class L, class Pred>
CNS_HD inline int find_first(const L& lanes, int from, int to, Pred&& pred)
{
	for (int base = from; base < to; base += L::count) {
		const uint32_t m = lanes.ballot([&](int l) { const int p = base + l; return p < to && pred(p); });
		if (m) return base + lowest_bit(m);
	}
	return to;
}

// check_cov_stats works on one coverage byte per template position (at most 100, so SWAR on 8 bytes never carries).
// The lanes split the aligned 8-byte words of [b, e); lane 0 takes the ragged ends.
template <class L>
CNS_HD inline int count_full(const L& lanes, const uint8_t* cov, int b, int e)   // positions in [b, e) covered >= COV_FULL times
{
	int head = b + (int)((8 - ((uintptr_t)(cov + b) & 7u)) & 7u);
	if (head > e) head = e;
	const int words = (e - head) >> 3, tail = head + 8 * words;
	const uint64_t* w = (const uint64_t*)(cov + head);
	return lanes.sum([&](int lane) {
		int full = 0;
		if (lane == 0) {
			for (int k = b; k < head; ++k) full += cov[k] >= COV_FULL;
			for (int k = tail; k < e; ++k) full += cov[k] >= COV_FULL;
		}
		for (int k = lane; k < words; k += L::count)
			full += popcount64((w[k] + 0x0101010101010101ull * (uint64_t)(128 - COV_FULL)) & 0x8080808080808080ull);
		return full;
	});
}
template <class L>
CNS_HD inline void bump_cov(const L& lanes, uint8_t* cov, int b, int e)
{
	int head = b + (int)((8 - ((uintptr_t)(cov + b) & 7u)) & 7u);
	if (head > e) head = e;
	const int words = (e - head) >> 3, tail = head + 8 * words;
	uint64_t* w = (uint64_t*)(cov + head);
	lanes.each([&](int lane) {
		if (lane == 0) {
			for (int k = b; k < head; ++k) ++cov[k];
			for (int k = tail; k < e; ++k) ++cov[k];
		}
		for (int k = lane; k < words; k += L::count) w[k] += 0x0101010101010101ull;
	});
}

// The accept loop of one read: candidates in trial order, at most 200 looked at, at most 60 accepted, one
// per partner read, mapping-range and coverage gates.  info = 8 ints per task {ok, qstart, qend, sstart,
// send, columns, ...} as written by the extension kernels.  Returns the number accepted; acc[k] = task.
template <class L>
CNS_HD inline int accept_read(const L& lanes, int t0, int t1, const int32_t* info, const int32_t* t_qid, const int32_t* t_qsize,
                              int ssize, double ratio, uint8_t* cov, int32_t* acc, int max_accept, int mode = 0)
{
	if (mode != 0) {
		// M4 input (consensus_one_read_m4_pacbio / _nanopore, mecat_correction.cpp:242-360): the caller has chosen the
		// read's overlaps; every alignment that succeeded is used -- mode 2 (nanopore): if it also spans enough of a read
		int added = 0;
		const int qss = (int)((double)ssize * ratio);
		for (int t = t0; t < t1 && added < max_accept; ++t) {
			const int32_t* o = info + 8 * (int64_t)t;
			if (!o[0]) continue;
			if (mode == 2) {
				const int oq = o[2] - o[1], os = o[4] - o[3];
				const int qqs = (int)((double)t_qsize[t] * ratio);
				if (!(oq >= qqs || os >= qss)) continue;
			}
			if (lanes.leader()) acc[added] = t;
			++added;
		}
		return added;
	}
	int used[MAX_ACCEPT];
	int added = 0, tried = 0;
	const int qss = (int)((double)ssize * ratio);
	for (int t = t0; t < t1 && added < max_accept && tried < MAX_TRIED; ++t) {
		++tried;
		const int qid = t_qid[t];
		bool seen = false;
		for (int k = 0; k < added; ++k) if (used[k] == qid) { seen = true; break; }
		if (seen) continue;
		const int32_t* o = info + 8 * (int64_t)t;
		if (!o[0]) continue;
		const int oq = o[2] - o[1], os = o[4] - o[3];
		const int qqs = (int)((double)t_qsize[t] * ratio);
		if (!(oq >= qqs || os >= qss)) continue;
		// a position can only be COV_FULL deep once COV_FULL alignments have been accepted
		const int full = added >= COV_FULL ? count_full(lanes, cov, o[3], o[4]) : 0;
		if (!(o[4] - o[3] >= full + COV_NEED)) continue;
		bump_cov(lanes, cov, o[3], o[4]);
		used[added] = qid;
		if (lanes.leader()) acc[added] = t;
		++added;
	}
	return added;
}

// ------------------------------------------------------------------------------------------ C4 + C5 fused
// normalize_gaps, meap_add_one_aln and column_index in ONE left-to-right pass with O(1) state, for the kernel
// that owns an accepted alignment.  Observation: the gap push only ever moves a base LEFT into a gap column,
// and it visits columns in ascending order, so when column i is visited every base whose expanded column is
// below i has been emitted already.  The current content of column i is therefore "the next unplaced base of
// that string if it sits exactly on column i, else a gap", and the look-ahead of the reference's inner loops
// (first non-gap character behind i) is simply that next unplaced base.  Two forward scanners stand on the next
// unplaced base of each string, the main loop walks the columns.  Columns are final when emitted, so votes and
// the cursor index are taken on the fly.  Sequential byte streams are read through 16-byte register windows.
struct ByteWindow          // forward-only reader of a byte stream through aligned 16-byte loads
{
	const char* base; int64_t at; unsigned long long lo, hi;
	CNS_HD void open(const char* p) { base = p; at = -1; lo = hi = 0; }
	CNS_HD char get(int64_t i)                 // i never decreases between calls by more than the window
	{
		const uintptr_t addr = (uintptr_t)(base + i);
		const int64_t blk = (int64_t)(addr >> 4);
		if (blk != at) {
			const unsigned long long* w = (const unsigned long long*)(addr & ~(uintptr_t)15);
			lo = w[0]; hi = w[1]; at = blk;
		}
		const int k = (int)(addr & 15u);
		return (char)(((k < 8 ? lo : hi) >> ((k & 7) * 8)) & 255u);
	}
};

struct ByteSink            // forward-only writer, 8 bytes per store; the destination must be 8-byte aligned
{
	char* base; int64_t n; unsigned long long acc;
	CNS_HD void open(char* p) { base = p; n = 0; acc = 0; }
	CNS_HD void put(char c)
	{
		acc |= (unsigned long long)(unsigned char)c << ((n & 7) * 8);
		if ((++n & 7) == 0) { *(unsigned long long*)(base + n - 8) = acc; acc = 0; }
	}
	CNS_HD void close() { if (n & 7) *(unsigned long long*)(base + (n & ~(int64_t)7)) = acc; }   // pads with NULs
};

// Returns the normalised length; nq/nt (8-byte aligned, 2n + 8 bytes) receive the normalised strings followed by
// a NUL, votes/base the read's pile-up, colidx the cursor index; *tend the last template position indexed.
// The loop runs over the ORIGINAL columns; a mismatch column (a, b) stands for the two expanded columns
// ('-', b), (a, '-'), so the base of t sits in half 0 and the base of q in half 1 of such a column.
CNS_HD inline int normalize_vote_index(const char* q0, const char* t0, int n, int soff0, char* nq, char* nt, uint32_t* votes,
                                       char* base, int32_t* colidx, int* tend)
{
	ByteWindow mq, mt, wq, wt;
	mq.open(q0); mt.open(t0); wq.open(q0); wt.open(t0);
	// heads: next unplaced base of q (column jq, character bq; it sits in the LAST half of its column) and of t
	// (column jt, character bt; first half)
	int jq = 0, jt = 0;
	char bq = 0, bt = 0;
	auto seek_q = [&](int from) {
		for (jq = from; jq < n; ++jq) {
			const char c = wq.get(jq);
			if (c != '-') { bq = c; return; }
		}
		bq = 0;
	};
	auto seek_t = [&](int from) {
		for (jt = from; jt < n; ++jt) {
			const char c = wt.get(jt);
			if (c != '-') { bt = c; return; }
		}
		bt = 0;
	};
	seek_q(0); seek_t(0);
	ByteSink oq, ot;
	oq.open(nq); ot.open(nt);
	int soff = soff0, cp = soff0;
	bool in_del_run = false;
	colidx[0] = 0;
	// One expanded column per iteration (a mismatch column takes two), so that the 32 alignments a warp works on all do
	// useful work in every iteration even though their columns differ.
	int i = 0, i0 = 0, h = 0, halves = 1;        // expanded column, original column, half, halves of the original column
	char a = 0, b = 0;
	if (n > 0) { a = mq.get(0); b = mt.get(0); halves = (a != b && a != '-' && b != '-') ? 2 : 1; }
	while (i0 < n) {
		const bool last = i0 == n - 1 && h == halves - 1;
		const bool q_here = jq == i0 && h == halves - 1, t_here = jt == i0 && h == 0;
		char qc = q_here ? bq : '-', tc = t_here ? bt : '-';
		if (!last) {
			if (!t_here && q_here) {
				if (jt < n && bt == qc) { tc = qc; seek_t(jt + 1); }
			} else if (!q_here && t_here) {
				if (jq < n && bq == tc) { qc = tc; seek_q(jq + 1); }
			}
		}
		if (q_here) seek_q(jq + 1);
		if (t_here) seek_t(jt + 1);
		oq.put(qc); ot.put(tc);
		// CnsAln cursor index (column_index)
		if (i >= 1 && tc != '-') { ++cp; colidx[cp - soff0] = i; }
		// meap_add_one_aln
		if (qc == '-' && tc == '-') { }
		else if (in_del_run && tc == '-') { }
		else {
			in_del_run = false;
			if (qc == tc) { vote_add(votes + soff, 1u); base[soff] = tc; ++soff; }
			else if (qc == '-') { vote_add(votes + soff, 1u << 8); ++soff; }
			else { vote_add(votes + soff - 1, 1u << 16); in_del_run = true; }
		}
		++i;
		if (++h == halves) {
			h = 0; ++i0;
			if (i0 < n) { a = mq.get(i0); b = mt.get(i0); halves = (a != b && a != '-' && b != '-') ? 2 : 1; }
		}
	}
	oq.put(0); ot.put(0);
	oq.close(); ot.close();
	*tend = cp;
	return i;
}

// ------------------------------------------------------------------------------------------ C6 (ranges)
struct Range { int start, end; };

// get_effective_ranges, mecat_correction.cpp:119-153.  m (nm <= MAX_ACCEPT entries) is reordered; returns the count in e.
CNS_HD inline int effective_ranges(Range* m, int nm, Range* e, int read_size, double size95)
{
	if (nm == 0) return 0;
	for (int i = 0; i < nm; ++i)
		if (m[i].start <= 500 && read_size - m[i].end <= 500) { e[0].start = 0; e[0].end = read_size; return 1; }
	for (int i = 1; i < nm; ++i) {                       // (start up, end down); equal keys are identical entries
		const Range x = m[i];
		int j = i - 1;
		while (j >= 0 && (m[j].start > x.start || (m[j].start == x.start && m[j].end < x.end))) { m[j + 1] = m[j]; --j; }
		m[j + 1] = x;
	}
	int ne = 0, i = 0, left = m[0].start, right;
	while (i < nm) {
		int j = i + 1;
		while (j < nm && m[j].end <= m[i].end) ++j;
		if (j == nm) {
			right = m[i].end;
			if ((double)(right - left) >= size95) { e[ne].start = left; e[ne].end = right; ++ne; }
			break;
		}
		if (m[i].end - m[j].start < 1000) {
			right = m[i].end < m[j].start ? m[i].end : m[j].start;
			if ((double)(right - left) >= size95) { e[ne].start = left; e[ne].end = right; ++ne; }
			left = m[i].end > m[j].start ? m[i].end : m[j].start;
		}
		i = j;
	}
	return ne;
}

// ------------------------------------------------------------------------------------------ C6 (segment walk)
// One ambiguous region between two anchors of a segment, i.e. one mini-POA.
struct Region
{
	int32_t read;         // batch-local read
	int32_t sb, se;       // template interval [sb, se], se = next anchor (or the segment end)
	int32_t prev_se;      // se of the read's previous region (-1: none): the state CnsAln's cursors are in
	int32_t min_weight;   // int(0.4 * coverage of the anchor)
	int32_t seg;          // batch-wide segment the region belongs to
	int32_t rank;         // ordinal of the anchor sb among the segment's anchors
};

// consensus_worker's run search inside the effective ranges (mecat_correction.cpp:203-239): maximal runs with coverage
// >= min_cov that are at least 0.95 * min_size long, found 32 positions per step.  segs receives up to cap {beg, end}
// pairs; the return value is the number found (callers size cap so that it always fits).  (tests/cns_literal.h has
// the sequential forms of this search and of the anchor walk, written after the reference line by line, which the
// unit tests compare these with.)
template <class L>
CNS_HD inline int find_segments(const L& lanes, const Range* e, int ne, const uint32_t* votes, int min_cov, double size95, int32_t* segs, int cap)
{
	int ns = 0;
	auto covered = [&](int p) { return vote_mat(votes[p]) + vote_ins(votes[p]) >= min_cov; };
	for (int r = 0; r < ne; ++r) {
		const int R = e[r].end;
		int beg = e[r].start;
		while (beg < R) {
			beg = find_first(lanes, beg, R, covered);
			int end = beg + 1;
			if (end < R) end = find_first(lanes, end, R, [&](int p) { return !covered(p); });
			if ((double)(end - beg) >= size95) {
				if (ns < cap && lanes.leader()) { segs[2 * ns] = beg; segs[2 * ns + 1] = end; }
				++ns;
			}
			beg = end;
		}
	}
	return ns;
}

// Anchor walk of one segment, 32 positions per step and without a sequential loop over the anchors: for a lane
// standing on an anchor, the previous anchor is the highest anchor bit below it (or the carry from earlier chunks),
// its interval needs refining iff a problem bit lies in between, and ordinals are popcounts of the bits below.
struct AnchorCarry
{
	int nanchors = 0;         // anchors seen so far in this segment
	int last_anchor = -1;     // position (segment relative) of the last one
	bool pending = false;     // a problem flag since (and including) the last anchor
	int nregions = 0;         // refine intervals closed so far in this segment
	int last_se_abs = -1;     // read coordinate where the read's latest region ended (-1: none yet)
};

// Positions [base, base + 32) of a segment of n positions that starts at read coordinate beg.  flag(p) returns
// classify() of segment position p.  visit(lane, pos, rank, prevpos, closes, ordinal, prev_se_abs) runs on every
// anchor lane: pos / prevpos = this and the previous anchor (segment relative, prevpos -1 for the first), rank =
// ordinal of this anchor, closes = the interval [prevpos, pos) is a region, ordinal = regions of this segment
// closed before this anchor, prev_se_abs = end of the read's latest region before that.
template <class L, class FlagFn, class Visit>
CNS_HD inline void anchor_chunk(const L& lanes, int base, int n, int beg, FlagFn&& flag, AnchorCarry& c, Visit&& visit)
{
	uint32_t A, P;
	lanes.ballot2([&](int l) {
		const int p = base + l;
		if (p >= n) return 0;
		const int f = flag(p);
		return ((f & FMAT) ? 1 : 0) | ((f & (UNDS | FDEL)) ? 2 : 0);
	}, A, P);
	const AnchorCarry c0 = c;
	const uint32_t R = lanes.ballot([&](int l) {
		if (!((A >> l) & 1u)) return false;
		const uint32_t below = A & bit_range(0, l);
		if (below) return (P & bit_range(highest_bit(below), l)) != 0;
		return c0.last_anchor >= 0 && (c0.pending || (P & bit_range(0, l)) != 0);
	});
	lanes.each([&](int l) {
		if (!((A >> l) & 1u)) return;
		const uint32_t below = A & bit_range(0, l), rbelow = R & bit_range(0, l);
		const int prevpos = below ? base + highest_bit(below) : c0.last_anchor;
		const int prev_se = rbelow ? beg + base + highest_bit(rbelow) : c0.last_se_abs;
		visit(l, base + l, c0.nanchors + popcount32(below), prevpos, ((R >> l) & 1u) != 0, c0.nregions + popcount32(rbelow), prev_se);
	});
	if (A) { const int hi = highest_bit(A); c.last_anchor = base + hi; c.pending = (P & bit_range(hi, 32)) != 0; }
	else if (c.last_anchor >= 0) c.pending = c.pending || P != 0;
	c.nanchors += popcount32(A);
	if (R) c.last_se_abs = beg + base + highest_bit(R);
	c.nregions += popcount32(R);
}

// ------------------------------------------------------------------------------------------ C7 (slices)
// An accepted alignment after normalisation, as the mini-POA sees it.
struct KeptAln
{
	const char* q; const char* s;     // normalised strings
	const int32_t* colidx;            // column_index()
	int32_t size, soff, send, tend;   // columns, template interval [soff, send), last position column_index reached
};

struct Slice { int32_t c0, c1, start; };   // columns [c0, c1] and the backbone position of the first one

CNS_HD inline int kept_column(const KeptAln& a, int pos)    // column the cursor rests on once it has reached `pos`
{
	if (pos <= a.soff) return 0;
	if (pos > a.tend) return a.size - 1;
	return a.colidx[pos - a.soff];
}

// CnsAln::retrieve_aln_subseqs for the region [sb, se].  The reference keeps one forward-only cursor per
// alignment; regions are asked for in increasing order and never overlap, so the cursor state before this
// call is a function of the previous region's end alone (prev_se), which makes regions independent.
CNS_HD inline bool kept_slice(const KeptAln& a, int sb, int se, int prev_se, Slice& out)
{
	if (se <= a.soff || sb >= a.send) return false;
	const int before = prev_se > a.soff ? kept_column(a, prev_se) : 0;
	if (before >= a.size - 1) return false;
	out.c0 = kept_column(a, sb);
	out.c1 = kept_column(a, se);
	out.start = (a.soff > sb ? a.soff : sb) - sb + 1;
	return true;
}

// ------------------------------------------------------------------------------------------ C7 (graph)
// AlnGraphBoost on flat arrays.  Adjacency lists are intrusive doubly linked lists threaded through the edges,
// which gives the iteration orders of boost::adjacency_list<vecS, vecS, bidirectionalS> (insertion order,
// order-preserving removal) that the tie-breaks of the best-path search depend on.
// The index type I is int32_t in general, int16_t for graphs below 30 000 nodes / edge slots and int8_t for graphs below
// 120 (the common case: ~8 nodes, ~34 edge slots), which shrinks the scratch of a graph to ~0.5 KB so that a few hundred
// of them fit the shared-memory pool of the graph kernel (cns.cu).
template <class I> struct PoaNodeT
{
	I coverage, weight, bb;                    // bb: _bbMap (node -> backbone node, 0 when never set)
	I in_head, in_tail, in_cnt;
	I out_head, out_tail, out_cnt;
	char base; uint8_t backbone;
};
template <class I> struct PoaEdgeT
{
	I u, v, count, visited;
	I in_next, in_prev;                        // position in v's in-list
	I out_next, out_prev;                      // position in u's out-list
};
static_assert(sizeof(PoaNodeT<int8_t>) == 11 && sizeof(PoaNodeT<int16_t>) == 20 && sizeof(PoaNodeT<int32_t>) == 40, "node layout");
static_assert(sizeof(PoaEdgeT<int8_t>) == 8 && sizeof(PoaEdgeT<int16_t>) == 16 && sizeof(PoaEdgeT<int32_t>) == 32, "edge layout");

// Scratch of one graph with N nodes and E0 edge creations before merging (region_demand): edge slots, queue/stack
// slots, and the bytes of the whole arena laid out as [score float N | nodes N | edges | queue+stack | best edge N].
CNS_HD inline int64_t poa_edge_cap(int64_t nodes, int64_t e0) { return e0 + nodes + 2; }
CNS_HD inline int64_t poa_aux_slots(int64_t nodes) { return 8 * nodes + 64; }
template <class I> CNS_HD inline int64_t poa_arena_bytes(int64_t nodes, int64_t e0)
{
	const int64_t b = 4 * nodes + (int64_t)sizeof(PoaNodeT<I>) * nodes + (int64_t)sizeof(PoaEdgeT<I>) * poa_edge_cap(nodes, e0) +
	                  (int64_t)sizeof(I) * (poa_aux_slots(nodes) + nodes);
	return (b + 3) & ~(int64_t)3;
}
// poa_arena_bytes<int32_t> is linear: 112 nodes + 32 e0 + 320 (the prefix sums of nodes and e0 locate every arena)
constexpr int POA_SMALL_LIMIT = 30000;     // nodes and edge slots below this -> int16_t indices are safe
constexpr int POA_TINY_LIMIT = 120;        // ... and below this int8_t ones (every count in a graph is bounded by its edge slots)
constexpr int POA_NO_BASE = -128;          // "no base processed yet" marker of the merge stack: below every base character, fits int8_t

enum { POA_OK = 0, POA_ERR_EDGES = 1, POA_ERR_QUEUE = 2, POA_ERR_STACK = 3, POA_ERR_EMPTY_LIST = 4, POA_ERR_NODES = 5 };

template <class I>
struct PoaT
{
	typedef PoaNodeT<I> PoaNode;
	typedef PoaEdgeT<I> PoaEdge;
	PoaNode* nd; PoaEdge* ed; I* aux; float* score; I* best_edge;
	int nn, ncap, ecap, nedges, efree, enter, exit_, err;
	int auxcap;

	CNS_HD void fail(int code) { if (!err) err = code; }

	CNS_HD int new_edge_slot()
	{
		if (efree >= 0) { const int e = efree; efree = ed[e].out_next; return e; }
		if (nedges >= ecap) { fail(POA_ERR_EDGES); return -1; }
		return nedges++;
	}
	CNS_HD int add_edge(int u, int v)
	{
		const int e = new_edge_slot();
		if (e < 0) return -1;
		PoaEdge& E = ed[e];
		E.u = u; E.v = v; E.count = 0; E.visited = 0;
		E.out_next = -1; E.out_prev = nd[u].out_tail;
		if (nd[u].out_tail >= 0) ed[nd[u].out_tail].out_next = e; else nd[u].out_head = e;
		nd[u].out_tail = e; ++nd[u].out_cnt;
		E.in_next = -1; E.in_prev = nd[v].in_tail;
		if (nd[v].in_tail >= 0) ed[nd[v].in_tail].in_next = e; else nd[v].in_head = e;
		nd[v].in_tail = e; ++nd[v].in_cnt;
		return e;
	}
	CNS_HD void unlink_out(int e)      // remove e from its source's out-list
	{
		PoaEdge& E = ed[e];
		PoaNode& U = nd[E.u];
		if (E.out_prev >= 0) ed[E.out_prev].out_next = E.out_next; else U.out_head = E.out_next;
		if (E.out_next >= 0) ed[E.out_next].out_prev = E.out_prev; else U.out_tail = E.out_prev;
		--U.out_cnt;
	}
	CNS_HD void unlink_in(int e)       // remove e from its target's in-list
	{
		PoaEdge& E = ed[e];
		PoaNode& V = nd[E.v];
		if (E.in_prev >= 0) ed[E.in_prev].in_next = E.in_next; else V.in_head = E.in_next;
		if (E.in_next >= 0) ed[E.in_next].in_prev = E.in_prev; else V.in_tail = E.in_prev;
		--V.in_cnt;
	}
	CNS_HD void release(int e) { ed[e].out_next = efree; efree = e; }
	CNS_HD void clear_vertex(int n)    // boost::clear_vertex: drop every edge of n
	{
		for (int e = nd[n].out_head; e >= 0;) { const int nx = ed[e].out_next; unlink_in(e); release(e); e = nx; }
		for (int e = nd[n].in_head; e >= 0;) { const int nx = ed[e].in_next; unlink_out(e); release(e); e = nx; }
		nd[n].out_head = nd[n].out_tail = nd[n].in_head = nd[n].in_tail = -1;
		nd[n].out_cnt = nd[n].in_cnt = 0;
	}
	CNS_HD int find_edge(int u, int v) const
	{
		for (int e = nd[u].out_head; e >= 0; e = ed[e].out_next) if (ed[e].v == v) return e;
		return -1;
	}
	CNS_HD void link(int u, int v)     // AlnGraphBoost::addEdge: bump the edge u -> v or create it
	{
		bool have = false;
		for (int e = nd[v].in_head; e >= 0; e = ed[e].in_next) if (ed[e].u == u) { ++ed[e].count; have = true; }
		if (!have) { const int e = add_edge(u, v); if (e >= 0) ++ed[e].count; }
	}
	CNS_HD void init_node(int i)
	{
		PoaNode& n = nd[i];
		n.coverage = 0; n.weight = 0; n.bb = 0;
		n.in_head = n.in_tail = n.out_head = n.out_tail = -1;
		n.in_cnt = n.out_cnt = 0;
		n.base = 'N'; n.backbone = 0;
	}

	// AlnGraphBoost(size_t blen), MECAT_AlnGraphBoost.C:76-97
	CNS_HD void init(char* arena, int node_cap, int edge_cap, int blen)
	{
		ncap = node_cap; ecap = edge_cap; auxcap = (int)poa_aux_slots(node_cap);
		score = (float*)arena;
		nd = (PoaNode*)(arena + 4 * (int64_t)ncap);
		ed = (PoaEdge*)(nd + ncap);
		aux = (I*)(ed + ecap);
		best_edge = aux + auxcap;
		nedges = 0; efree = -1; err = 0;
		nn = blen + 2;
		if (nn > ncap) { fail(POA_ERR_NODES); nn = 0; enter = exit_ = 0; return; }
		for (int i = 0; i < nn; ++i) init_node(i);
		for (int i = 0; i < blen + 1; ++i) add_edge(i, i + 1);
		enter = 0; exit_ = blen + 1;
		nd[enter].base = '^'; nd[enter].backbone = 1;
		for (int i = 1; i <= blen; ++i) { nd[i].backbone = 1; nd[i].weight = 1; nd[i].base = 'N'; nd[i].bb = i; }
		nd[exit_].base = '$'; nd[exit_].backbone = 1;
	}

	// addAln, MECAT_AlnGraphBoost.C:99-152, on columns [c0, c1] of a normalised alignment
	CNS_HD void add_alignment(const char* q, const char* t, int c0, int c1, int start)
	{
		if (err) return;
		int pos = start, prev = enter;
		for (int i = c0; i <= c1; ++i) {
			const char a = q[i], b = t[i];
			if (a == '-' && b == '-') continue;
			const int cur = pos;
			if (a == b) {
				PoaNode& B = nd[nd[cur].bb];
				++B.coverage; B.base = b;
				++nd[cur].weight;
				link(prev, cur);
				++pos; prev = cur;
			} else if (a == '-') {
				PoaNode& B = nd[nd[cur].bb];
				++B.coverage; B.base = b;
				++pos;
			} else {
				if (nn >= ncap) { fail(POA_ERR_NODES); return; }
				const int nv = nn++;
				init_node(nv);
				nd[nv].base = a; ++nd[nv].weight;
				nd[nv].bb = pos;
				link(prev, nv);
				prev = nv;
			}
		}
		link(prev, exit_);
	}

	// smallest base above `last` among list[0..cnt); -1 when none (std::map<char, ...> iteration order)
	CNS_HD int next_base(const I* list, int cnt, int last) const
	{
		int best = -1;
		for (int k = 0; k < cnt; ++k) {
			const int b = (int)(signed char)nd[list[k]].base;
			if (b > last && (best < 0 || b < best)) best = b;
		}
		return best;
	}

	// mergeInNodes, MECAT_AlnGraphBoost.C:219-287.  The recursion on the anchor node is unrolled on an explicit
	// stack in aux[sp0 ...): a frame is the snapshot of the predecessors with a single out-edge, its length and
	// the last base processed.
	CNS_HD void merge_in(int n0, int sp0)
	{
		int sp = sp0;
		auto push_frame = [&](int n) {
			const int begin = sp;
			for (int e = nd[n].in_head; e >= 0; e = ed[e].in_next) {
				const int u = ed[e].u;
				if (nd[u].out_cnt == 1) { if (sp + 3 > auxcap) { fail(POA_ERR_STACK); return; } aux[sp++] = (I)u; }
			}
			if (sp + 2 > auxcap) { fail(POA_ERR_STACK); return; }
			const int entries = sp - begin;
			aux[sp++] = (I)entries;
			aux[sp++] = (I)POA_NO_BASE;    // last base done
		};
		push_frame(n0);
		while (sp > sp0 && !err) {
			const int cnt = aux[sp - 2];
			I* list = aux + (sp - 2 - cnt);
			const int b = next_base(list, cnt, aux[sp - 1]);
			if (b < 0) { sp -= cnt + 2; continue; }
			aux[sp - 1] = (I)b;
			int an = -1, members = 0;
			for (int k = 0; k < cnt; ++k) if ((int)(signed char)nd[list[k]].base == b) { if (an < 0) an = list[k]; ++members; }
			if (members <= 1) continue;
			const int an_out = nd[an].out_head;
			if (an_out < 0) { fail(POA_ERR_EMPTY_LIST); return; }
			for (int k = 0; k < cnt; ++k) {
				const int x = list[k];
				if (x == an || (int)(signed char)nd[x].base != b) continue;
				if (nd[x].out_head < 0) { fail(POA_ERR_EMPTY_LIST); return; }
				ed[an_out].count += ed[nd[x].out_head].count;
				nd[an].weight += nd[x].weight;
			}
			for (int k = 0; k < cnt; ++k) {
				const int x = list[k];
				if (x == an || (int)(signed char)nd[x].base != b) continue;
				for (int ie = nd[x].in_head; ie >= 0; ie = ed[ie].in_next) {
					const int n1 = ed[ie].u;
					const int e = find_edge(n1, an);
					if (e >= 0) ed[e].count += ed[ie].count;
					else { const int ne = add_edge(n1, an); if (ne < 0) return; ed[ne].count = ed[ie].count; ed[ne].visited = ed[ie].visited; }
				}
				clear_vertex(x);
			}
			push_frame(an);
		}
	}

	// mergeOutNodes, MECAT_AlnGraphBoost.C:289-357 (no recursion)
	CNS_HD void merge_out(int n, int sp0)
	{
		int cnt = 0;
		I* list = aux + sp0;
		for (int e = nd[n].out_head; e >= 0; e = ed[e].out_next) {
			const int v = ed[e].v;
			if (nd[v].in_cnt == 1) { if (sp0 + cnt + 1 > auxcap) { fail(POA_ERR_STACK); return; } list[cnt++] = (I)v; }
		}
		int last = POA_NO_BASE;
		for (;;) {
			const int b = next_base(list, cnt, last);
			if (b < 0 || err) break;
			last = b;
			int an = -1, members = 0;
			for (int k = 0; k < cnt; ++k) if ((int)(signed char)nd[list[k]].base == b) { if (an < 0) an = list[k]; ++members; }
			if (members <= 1) continue;
			const int an_in = nd[an].in_head;
			if (an_in < 0) { fail(POA_ERR_EMPTY_LIST); return; }
			for (int k = 0; k < cnt; ++k) {
				const int x = list[k];
				if (x == an || (int)(signed char)nd[x].base != b) continue;
				if (nd[x].in_head < 0) { fail(POA_ERR_EMPTY_LIST); return; }
				ed[an_in].count += ed[nd[x].in_head].count;
				nd[an].weight += nd[x].weight;
			}
			for (int k = 0; k < cnt; ++k) {
				const int x = list[k];
				if (x == an || (int)(signed char)nd[x].base != b) continue;
				for (int oe = nd[x].out_head; oe >= 0; oe = ed[oe].out_next) {
					const int n2 = ed[oe].v;
					const int e = find_edge(an, n2);
					if (e >= 0) ed[e].count += ed[oe].count;
					else { const int ne = add_edge(an, n2); if (ne < 0) return; ed[ne].count = ed[oe].count; ed[ne].visited = ed[oe].visited; }
				}
				clear_vertex(x);
			}
		}
	}

	// mergeNodes, MECAT_AlnGraphBoost.C:199-217.  aux = [queue of qcap ints | merge stack]
	CNS_HD void merge_nodes()
	{
		if (err) return;
		const int qcap = 2 * ncap + 16;
		int head = 0, tail = 0;          // monotone counters into the circular queue aux[0, qcap)
		aux[tail++ % qcap] = (I)enter;
		while (head < tail && !err) {
			const int u = aux[head++ % qcap];
			merge_in(u, qcap);
			merge_out(u, qcap);
			for (int e = nd[u].out_head; e >= 0; e = ed[e].out_next) {
				ed[e].visited = 1;
				const int v = ed[e].v;
				int open = 0;
				for (int ie = nd[v].in_head; ie >= 0; ie = ed[ie].in_next) if (!ed[ie].visited) ++open;
				if (open == 0) {
					if (tail - head >= qcap) { fail(POA_ERR_QUEUE); return; }
					aux[tail++ % qcap] = (I)v;
				}
			}
		}
	}

	// consensus + bestPath, MECAT_AlnGraphBoost.C:417-458,508-592.  Writes the bases of the best path into out
	// (needs nn bytes) and returns, through off/len, the longest run of nodes with weight >= min_weight.
	// aux = [queue | float score per node | best edge per node]
	CNS_HD void consensus(int min_weight, char* out, int& off, int& len)
	{
		off = 0; len = 0;
		if (err) return;
		const int qcap = 2 * ncap + 16;
		for (int i = 0; i < nn; ++i) { score[i] = 0.0f; best_edge[i] = (I)-1; }
		for (int e = 0; e < nedges; ++e) ed[e].visited = 0;
		int head = 0, tail = 0;
		aux[tail++ % qcap] = (I)exit_;
		while (head < tail) {
			const int n = aux[head++ % qcap];
			bool found = false;
			float best = -3.402823466e+38f;
			int best_e = -1;
			for (int e = nd[n].out_head; e >= 0; e = ed[e].out_next) {
				const int v = ed[e].v;
				const float s = score[v];
				float ns;
				if (nd[v].backbone && nd[v].weight == 1) ns = s - 10.0f;
				else ns = (float)ed[e].count - (float)nd[nd[v].bb].coverage * 0.5f + s;
				if (ns > best) { best = ns; best_e = e; found = true; }
			}
			if (found) { score[n] = best; best_edge[n] = (I)best_e; }
			for (int ie = nd[n].in_head; ie >= 0; ie = ed[ie].in_next) {
				ed[ie].visited = 1;
				const int u = ed[ie].u;
				int open = 0;
				for (int oe = nd[u].out_head; oe >= 0; oe = ed[oe].out_next) if (!ed[oe].visited) ++open;
				if (open == 0) {
					if (tail - head >= qcap) { fail(POA_ERR_QUEUE); return; }
					aux[tail++ % qcap] = (I)u;
				}
			}
		}
		int offs = 0, best_offs = 0, length = 0, idx = 0;
		bool met = false;
		for (int p = enter, guard = 0; guard <= nn; ++guard) {
			const PoaNode& N = nd[p];
			if (!(N.base == '^' || N.base == '$')) {
				out[idx] = N.base;
				if (!met && N.weight >= min_weight) { offs = idx; met = true; }
				else if (met && N.weight < min_weight) {
					if (idx - offs > length) { best_offs = offs; length = idx - offs; }
					met = false;
				}
				++idx;
			}
			if (best_edge[p] < 0) break;
			p = ed[best_edge[p]].v;
		}
		if (met && idx - offs > length) { best_offs = offs; length = idx - offs; }
		off = best_offs; len = length;
	}
};

// Node and initial-edge demand of one region: counted from the slices before the graph is built, so that
// every graph gets an arena of exactly its own size.
CNS_HD inline void region_demand(const KeptAln* kept, int nkept, int sb, int se, int prev_se, int& nodes, int& edges0)
{
	const int blen = se - sb + 1;
	int n = blen + 2, e = blen + 1;
	for (int k = 0; k < nkept; ++k) {
		Slice sl;
		if (!kept_slice(kept[k], sb, se, prev_se, sl)) continue;
		const char* q = kept[k].q; const char* s = kept[k].s;
		for (int i = sl.c0; i <= sl.c1; ++i) {
			const char a = q[i], b = s[i];
			if (a == '-') continue;
			++e;                              // a match or an extra base links one edge
			if (a != b) ++n;                  // an extra base also creates a node
		}
		++e;                                  // link to the exit node
	}
	nodes = n; edges0 = e;
}

// meap_cns_one_indel, mecat_correction.cpp:63-78: graph of one region -> best-path bases in out[0..), the kept run
// in off/len.  arena: poa_arena_bytes<I>(node_cap, e0) bytes, 4-byte aligned.  Returns a POA_* status.
template <class I>
CNS_HD inline int region_consensus(const KeptAln* kept, int nkept, int sb, int se, int prev_se, int min_weight,
                                   char* arena, int node_cap, int edge_cap, char* out, int& off, int& len)
{
	PoaT<I> g;
	g.init(arena, node_cap, edge_cap, se - sb + 1);
	for (int k = 0; k < nkept; ++k) {
		Slice sl;
		if (kept_slice(kept[k], sb, se, prev_se, sl)) g.add_alignment(kept[k].q, kept[k].s, sl.c0, sl.c1, sl.start);
	}
	g.merge_nodes();
	g.consensus(min_weight, out, off, len);
	if (g.err) { off = 0; len = 0; }
	return g.err;
}

}  // namespace mbcns
