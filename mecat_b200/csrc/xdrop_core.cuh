// mecat_b200/csrc/xdrop_core.cuh -- per-chain body of the nanopore (`-x 1`) gapped extension.
//
// With `-x 1` the reference extends candidates with XdropAligner instead of the O(nd) diff aligner:
//   XdropAligner::go, align_ex, xdrop_align, script_to_aligned_string   src/common/xdrop_gapalign.cpp:10-439
//   XdropAlignParameters::init(0)                                      src/common/xdrop_gapalign.h:85-103
//   retrieve_next_aln_block, trim_mismatch_end                          src/common/gapalign.cpp:10-67
// a BLAST-style X-drop dynamic programme (match +1, mismatch -1, gap -1 per base, X = 30) over a chain of 500-base
// blocks, with a trace-back per block.  A row of that programme is a sequential scan: the running best score prunes
// cells of the same row to its right, and a pruned cell leaves both gap scores as they were.  One chain -- one
// (candidate, direction) -- is therefore the unit of parallel work here: one thread walks its rows exactly like the
// reference does, all chains of a batch in flight at once.
//
// Per chain scratch (global memory, private to the thread): the score row (8 bytes per subject base of the block), the
// trace-back of the current block packed 4 bits per cell (operation + the two "gap continues" flags), and the first cell /
// first word of every row.  A block needs no more than BLOCK_MAX x (BLOCK_MIN_SIDE + 1) cells (one side of a block is
// below 600 bases, the other below 720: gapalign.cpp:26-34), so the scratch is sized for the worst case and a chain can
// never run out of it.
//
// Columns leave the kernel like those of align.cu (ASCII, walking order, one slot per chain); the string-free form only
// counts them.  Integer code shared by the CUDA backend (xdrop.cu) and the host harness of the CPU test-suite
// (tests/xdrop_host_harness.cpp).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define XD_HD __host__ __device__ __forceinline__
#else
#define XD_HD inline
#endif

namespace mbx {

constexpr int REWARD = 1, PENALTY = -1, GAP_OPEN = 0, GAP_EXTEND = 1, X_DROPOFF = 30;   // xdrop_gapalign.h:85-103
constexpr int BLOCK = 500;
constexpr int NEG = -100000000;                                  // MIN_SCORE, xdrop_gapalign.cpp:8
constexpr int SIDE_MAX = 720;                                    // longest side of a block: min(599 * 1.2, ...) = 718
constexpr int SIDE_MIN_MAX = 600;                                // the shorter side is below 600
constexpr int ROWS = SIDE_MAX + 2;
constexpr int TB_WORDS = ((SIDE_MIN_MAX + 2) * (SIDE_MAX + 2)) / 8 + ROWS;   // 4 bits per cell + one word of slack per row
constexpr int SC_CELLS = SIDE_MAX + 4;

// trace-back nibble: operation in bits 0-1, the flag the reference calls SCRIPT_EXTEND_GAP_A (set when the COLUMN gap
// score continues) in bit 2, SCRIPT_EXTEND_GAP_B (the ROW gap score continues) in bit 3 -- xdrop_gapalign.cpp:96-98,128-138
enum { OP_SUB = 0, OP_GAP_A = 1 /* gap in the query: one subject base */, OP_GAP_B = 2 /* one query base */, F_EXT_A = 4, F_EXT_B = 8 };

struct Cell { int32_t best, gap; };

struct Seq            // bases of one walking direction: base i = two bits at position g0 + i of a packed word array
{
	const uint32_t* arr; uint32_t g0; uint32_t comp; int32_t len;
};
XD_HD int base_at(const Seq& s, int i)
{
	const uint32_t p = s.g0 + (uint32_t)i;
	return (int)(((s.arr[p >> 4] ^ s.comp) >> ((p & 15u) << 1)) & 3u);
}

struct Half           // what one chain produced
{
	int32_t cols, matches, qadv, tadv;
	int32_t last;      // the last column in walking order: bit 0 = it holds a query base, bit 1 = a subject base, bit 2 = equal letters
	int32_t overflow;  // the column slot was too small (cannot happen with slots sized by align_task_columns)
};

struct RingCell { int16_t best, gap; };      // a score-row entry in shared memory

struct Scratch        // private to one thread
{
	Cell* sc;          // SC_CELLS (global memory: the score row of blocks whose band outgrows the ring)
	uint32_t* tb;      // TB_WORDS
	int32_t* row_first;// ROWS
	int32_t* row_word; // ROWS
	RingCell* ring;    // RING entries, `ring_stride` apart (shared memory, interleaved between the threads of a CTA so that
	int ring_stride;   // every lane stays in its own bank); nullptr: no ring
};

// The score row of the X-drop programme only lives between the first cell still inside the drop-off and the sentinel
// behind the last one: ~60 cells on related sequences, a few hundred on unrelated ones (band statistics in DESIGN.md).
// RingRow keeps that window in shared memory (cell b at entry b mod RING, 16-bit scores); a block whose window would not
// fit is redone with the row in global memory.  "Minus infinity" is a property of the row type: the reference's -10^8 for
// the 32-bit row, -16 384 for the 16-bit one.  A value derived from it is the constant plus or minus a few (it is never
// stored: what a cell stores is a real score -- within [-40, 720] -- or the constant itself), real scores never come near
// either constant, and every comparison of the programme is between two real values, a real value and such a derived one,
// or two derived ones at the same small offsets -- so both constants give the same decisions.
constexpr int RING = 128;
struct GlobalRow
{
	Cell* p;
	XD_HD Cell get(int b) const { return p[b]; }
	XD_HD void put(int b, Cell c) const { p[b] = c; }
	XD_HD void put_best(int b, int v) const { p[b].best = v; }
	XD_HD bool room(int, int) const { return true; }
	XD_HD static int neg() { return NEG; }
};
struct RingRow
{
	RingCell* p; int stride;
	XD_HD Cell get(int b) const { const RingCell r = p[(b & (RING - 1)) * stride]; Cell c; c.best = r.best; c.gap = r.gap; return c; }
	XD_HD void put(int b, Cell c) const { RingCell r; r.best = (int16_t)c.best; r.gap = (int16_t)c.gap; p[(b & (RING - 1)) * stride] = r; }
	XD_HD void put_best(int b, int v) const { p[(b & (RING - 1)) * stride].best = (int16_t)v; }
	XD_HD static int neg() { return -16384; }
	XD_HD bool room(int first_b, int b) const { return b - first_b < RING; }      // may cell b be written while first_b is live?
};
constexpr size_t SCRATCH_BYTES = sizeof(Cell) * SC_CELLS + 4 * (size_t)TB_WORDS + 8 * (size_t)ROWS;

struct RowWriter      // 4 bits per cell into the trace-back words of a row; `flush` once per round of up to 4 cells
{
	uint32_t* tb; int w0, n, held; unsigned long long acc; int bad;
	XD_HD void begin(uint32_t* tb_, int w0_) { tb = tb_; w0 = w0_; n = 0; held = 0; acc = 0; }
	XD_HD void put(uint32_t nib) { acc |= (unsigned long long)nib << (held << 2); ++held; }      // at most 11 held between flushes
	XD_HD void flush()
	{
		if (held >= 8) {
			const int w = w0 + (n >> 3);
			if (w < TB_WORDS) tb[w] = (uint32_t)acc; else bad = 1;
			acc >>= 32; held -= 8; n += 8;
		}
	}
	XD_HD int end()      // words used by the row
	{
		flush();
		if (held) {
			const int w = w0 + (n >> 3);
			if (w < TB_WORDS) tb[w] = (uint32_t)acc; else bad = 1;
		}
		return (n + held + 7) >> 3;
	}
};

// 16 bases starting at base i of the walk, base i in the low bits
XD_HD uint32_t bases16(const Seq& s, int i)
{
	const uint32_t p = s.g0 + (uint32_t)i, w = p >> 4, sh = (p & 15u) << 1;
	const unsigned long long x = ((unsigned long long)(s.arr[w + 1] ^ s.comp) << 32) | (s.arr[w] ^ s.comp);
	return (uint32_t)(x >> sh);
}

// xdrop_align without its trace-back walk (xdrop_gapalign.cpp:10-158): fills the trace-back of the block, returns the
// end cell of the best path.  A = query block (M bases from q0), B = subject block (N bases from t0).
// Returns false when the row does not fit `sc` (RingRow only): nothing of the block is valid then.
template <class Row>
XD_HD bool block_dp(const Row& sc, const Seq& Q, int q0, int M, const Seq& T, int t0, int N, const Scratch& S, int& ae, int& be, int& bad)
{
	ae = be = 0;
	if (M <= 0 || N <= 0) return true;
	const int oe = GAP_OPEN + GAP_EXTEND;
	const int xd = X_DROPOFF < oe ? oe : X_DROPOFF;
	const int neg = Row::neg();
	RowWriter W; W.bad = 0;
	int next_word = 0;
	int score = -oe, i;
	{ Cell c0; c0.best = 0; c0.gap = -oe; sc.put(0, c0); }
	S.row_first[0] = 0; S.row_word[0] = 0;
	W.begin(S.tb, 0);
	W.put(OP_SUB);                                   // cell (0, 0): never read
	for (i = 1; i <= N; ++i) {
		if (score < -xd) break;
		{ Cell ci; ci.best = score; ci.gap = score - oe; sc.put(i, ci); }
		score -= GAP_EXTEND;
		W.put(OP_GAP_A);
		W.flush();
	}
	next_word += W.end();
	int b_size = i, best_score = 0, first_b = 0;
	for (int a = 1; a <= M; ++a) {
		const int ac = base_at(Q, q0 + a - 1);
		S.row_first[a] = first_b; S.row_word[a] = next_word;
		W.begin(S.tb, next_word);
		score = neg;
		int gap_row = neg, last_b = first_b, b;
		// four cells per round: their score-row entries and subject bases are loaded together (independent loads in
		// flight instead of one dependent load per cell), then the cells are finished one after the other as before
		for (b = first_b; b < b_size;) {
			const int nb = b_size - b < 4 ? b_size - b : 4;
			Cell pre[4];
			const uint32_t tb4 = bases16(T, t0 + b);          // the subject bases of the round (and 12 more) in one double-word fetch
#if defined(__CUDA_ARCH__)
			#pragma unroll
#endif
			for (int j = 0; j < 4; ++j)
				if (j < nb) pre[j] = sc.get(b + j);
#if defined(__CUDA_ARCH__)
			#pragma unroll
#endif
			for (int j = 0; j < 4; ++j) {
				if (j >= nb) break;
				const int bc = (int)((tb4 >> (2 * j)) & 3u);
				const Cell c = pre[j];
				int gap_col = c.gap;
				const int next = c.best + (ac == bc ? REWARD : PENALTY);
				uint32_t script = OP_SUB;
				if (score < gap_col) { script = OP_GAP_B; score = gap_col; }
				if (score < gap_row) { script = OP_GAP_A; score = gap_row; }
				if (best_score - score > xd) {
					if (first_b == b) ++first_b;
					else sc.put_best(b, neg);
				} else {
					last_b = b;
					if (score > best_score) { best_score = score; ae = a; be = b; }
					Cell o;
					gap_col -= GAP_EXTEND;
					if (gap_col < score - oe) o.gap = score - oe;
					else { o.gap = gap_col; script |= F_EXT_A; }
					gap_row -= GAP_EXTEND;
					if (gap_row < score - oe) gap_row = score - oe;
					else script |= F_EXT_B;
					o.best = score;
					sc.put(b, o);
				}
				score = next;
				W.put(script);
				++b;
			}
			W.flush();
		}
#if defined(XD_STATS)
		{ const int w = b_size - S.row_first[a]; ++g_rows; g_cells += w; if (w > g_maxw) g_maxw = w; if (w > g_blockmax) g_blockmax = w; }
#endif
		if (first_b == b_size) { next_word += W.end(); break; }
		if (last_b < b_size - 1) b_size = last_b + 1;
		else {
			while (gap_row >= best_score - xd && b_size < N) {
				if (!sc.room(first_b, b_size)) return false;
				Cell ce; ce.best = gap_row; ce.gap = gap_row - oe;
				sc.put(b_size, ce);
				gap_row -= GAP_EXTEND;
				W.put(OP_GAP_A);
				W.flush();
				++b_size;
			}
		}
		next_word += W.end();
		if (b_size < N) {
			if (!sc.room(first_b, b_size)) return false;
			Cell cs; cs.best = neg; cs.gap = neg;
			sc.put(b_size, cs);
			++b_size;
		}
	}
	bad |= W.bad;
#if defined(XD_STATS)
	++g_blocks; ++g_hist[g_blockmax / 32 > 15 ? 15 : g_blockmax / 32]; g_blockmax = 0;
#endif
	return true;
}

XD_HD uint32_t tb_get(const Scratch& S, int a, int b)
{
	const int i = b - S.row_first[a];
	return (S.tb[S.row_word[a] + (i >> 3)] >> ((i & 7) << 2)) & 15u;
}

// One step of the trace-back walk (xdrop_gapalign.cpp:166-201): the operation that ends in cell (a, b); moves (a, b).
XD_HD uint32_t tb_step(const Scratch& S, uint32_t script, int& a, int& b)
{
	const uint32_t nib = tb_get(S, a, b);
	const uint32_t op = nib & 3u;
	if (script == OP_GAP_A) script = (nib & F_EXT_A) ? (uint32_t)OP_GAP_A : op;
	else if (script == OP_GAP_B) script = (nib & F_EXT_B) ? (uint32_t)OP_GAP_B : op;
	else script = op;
	if (script == OP_GAP_A) --b;
	else if (script == OP_GAP_B) --a;
	else { --a; --b; }
	return script;
}

// align_ex (xdrop_gapalign.cpp:249-349): the chain of blocks of one direction.  COLS: the columns are written to oq / ot
// (ASCII, walking order, at most cap).
template <bool COLS>
XD_HD void chain(const Seq& Q, const Seq& T, const Scratch& S, char* oq, char* ot, int cap, Half& H)
{
	int qi = 0, ti = 0;
	int cols = 0, mats = 0, qadv = 0, tadv = 0, last = 0, overflow = 0;
	for (;;) {
		const int qleft = Q.len - qi, tleft = T.len - ti;
		int qblk, tblk;
		bool lastblk;
		if (qleft < BLOCK + 100 || tleft < BLOCK + 100) {        // retrieve_next_aln_block, gapalign.cpp:24-34
			const int a = (int)((double)tleft + (double)tleft * 0.2), b = (int)((double)qleft + (double)qleft * 0.2);
			qblk = qleft < a ? qleft : a; tblk = tleft < b ? tleft : b; lastblk = true;
		} else { qblk = tblk = BLOCK; lastblk = false; }
		int ae, be, bad = 0;
		bool done = false;
		if (S.ring) { RingRow rr; rr.p = S.ring; rr.stride = S.ring_stride; done = block_dp(rr, Q, qi, qblk, T, ti, tblk, S, ae, be, bad); }
		if (!done) { bad = 0; GlobalRow gr; gr.p = S.sc; block_dp(gr, Q, qi, qblk, T, ti, tblk, S, ae, be, bad); }
		if (bad) { overflow = 1; break; }
		const bool full = (qblk - ae <= 20 || tblk - be <= 20);
		const bool whole = !full || lastblk;                     // this block's alignment is appended as it is and ends the chain
		// first walk: column count, matches, and trim_mismatch_end's scan from the tail (gapalign.cpp:48-67)
		int n = 0, m = 0, acnt = 0, qcnt = 0, tcnt = 0, mat_all = 0, mat_cut = 0, flag_first = 0, flag_kept = 0;
		bool scanning = !whole;
		{
			int a = ae, b = be;
			uint32_t script = OP_SUB;
			while (a > 0 || b > 0) {
				script = tb_step(S, script, a, b);
				const int eq = script == OP_SUB && base_at(Q, qi + a) == base_at(T, ti + b);
				const int fl = (script != OP_GAP_A ? 1 : 0) | (script != OP_GAP_B ? 2 : 0) | (eq ? 4 : 0);
				if (n == 0) flag_first = fl;
				if (!scanning && n == acnt) flag_kept = fl;
				if (scanning) {
					++acnt; qcnt += fl & 1; tcnt += (fl >> 1) & 1;
					if (eq) { ++m; ++mat_cut; } else m = 0;
					if (m == 4) scanning = false;
				}
				mat_all += eq;
				++n;
			}
		}
		int skip, keep;
		bool stop;
		if (whole) { skip = 0; keep = n; stop = true; flag_kept = flag_first; }
		else if (m == 4 && n - acnt >= 2) { skip = acnt; keep = n - acnt; stop = false; }
		else break;                                              // no 4-match tail: the block is dropped, the chain ends
		if (keep > 0) {
			if (COLS) {
				if (cols + keep > cap) { overflow = 1; break; }
				int a = ae, b = be, j = 0;
				uint32_t script = OP_SUB;
				while (a > 0 || b > 0) {
					script = tb_step(S, script, a, b);
					if (j >= skip) {
						const int at = cols + keep - 1 - (j - skip);
						oq[at] = script == OP_GAP_A ? '-' : "ACGT"[base_at(Q, qi + a)];
						ot[at] = script == OP_GAP_B ? '-' : "ACGT"[base_at(T, ti + b)];
					}
					++j;
				}
			}
			cols += keep;
			mats += mat_all - (whole ? 0 : mat_cut);
			last = flag_kept;
		}
		if (stop) { qadv = qi + ae; tadv = ti + be; break; }
		qi += ae - qcnt; ti += be - tcnt;
		qadv = qi; tadv = ti;
	}
	H.cols = cols; H.matches = mats; H.qadv = qadv; H.tadv = tadv; H.last = last; H.overflow = overflow;
}

// XdropAligner::go (xdrop_gapalign.cpp:351-439) from the two halves: the left part is emitted from its last column but
// one (:392-393), ok = the query span reaches min_aln (:438).  out = {ok, qstart, qend, sstart, send, columns, matches, first}
XD_HD void finish(int qstart, int sstart, const Half& L, const Half& R, int min_aln, int32_t* out)
{
	const int drop = L.cols >= 1 ? 1 : 0;
	const int lq = L.qadv - (drop ? (L.last & 1) : 0), lt = L.tadv - (drop ? ((L.last >> 1) & 1) : 0);
	const int lm = L.matches - (drop ? ((L.last >> 2) & 1) : 0);
	const int qs = qstart - lq, qe = qstart + R.qadv, ss = sstart - lt, se = sstart + R.tadv;
	const int bad = L.overflow | R.overflow;
	out[0] = (qe - qs >= min_aln) && !bad;
	out[1] = qs; out[2] = qe; out[3] = ss; out[4] = se;
	out[5] = L.cols - drop + R.cols; out[6] = lm + R.matches; out[7] = drop;
}

}  // namespace mbx
