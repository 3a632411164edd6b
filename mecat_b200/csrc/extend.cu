// mecat_b200/csrc/extend.cu -- batched O(nd) gapped extension (pw / ref flavour).
//
// Replaces, for a whole batch of candidates at once, the reference's per-candidate
//   DiffAligner::go            src/common/diff_gapalign.cpp:295-349
//   dw_in_one_direction        src/common/diff_gapalign.cpp:221-292
//   retrieve_next_aln_block    src/common/gapalign.cpp:10-45
//   Align                      src/common/diff_gapalign.cpp:107-219
//   GetAlignString + trim_mismatch_end   diff_gapalign.cpp:40-104, gapalign.cpp:48-67
//
// Two kernels.  k_extend_lanes (the default) gives every LANE its own (candidate, direction) chain and walks the
// furthest-reaching recurrence cell by cell, J cells per step; k_extend (below, the first form) gives a chain to a
// whole warp and is kept for the rare rows whose band does not fit a lane's ring, and as an A/B switch.
//
// k_extend (warp per chain): one warp owns one (candidate, direction) chain of <= 720 x 720 blocks.  The two
// block operands are staged 2 bit/base in shared memory in walking order, so a snake
// compares 16 bases per XOR + FFS.  One row of the furthest-reaching recurrence is one warp
// step: lane j owns diagonal min_k + 2j (32 diagonals per pass); the neighbours k-1 / k+1 of
// the previous row are read from a per-warp shared array that is updated in place (a row
// only reads cells of the other parity, exactly like the reference's V array).
//
// No traceback is stored.  What the reference needs from the alignment string of a block is
// (a) its length and (b) where the last run of >= 4 matching columns ends
// (trim_mismatch_end).  A diff alignment has no mismatch columns -- it is snakes separated by
// single indels -- so that run is the tail of the last snake of length >= 4 on the path.
// Every cell therefore carries, next to its furthest x, the packed (x, y, d) of the last
// long snake on its own path ("anchor"); the end cell's anchor gives, in closed form,
//   columns up to the run end = (x + y + d) / 2,   matches = (x + y - d) / 2.
#include "common.cuh"

#include <algorithm>

namespace mb {

namespace {

constexpr int EXT_WARPS = 6;
constexpr int EXT_CTAS_PER_SM = 5;        // 5 x 6 warps x 7.3 KB of per-warp state = 219 KB of shared memory per SM
constexpr int KOFF = 404;                 // even, > max_d of the largest block (0.3 * (599 + 718) = 395)
constexpr int VL_N = KOFF + 4;            // diagonals per parity
constexpr int SEQ_WORDS = 48;             // 719 bases = 45 words (+1 funnel, +2 slack)
constexpr uint32_t NO_ANCHOR = 0xFFFFFFFFu;
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int IDLE = -0x7fffffff - 1;     // x + y of a lane without a cell: below every threshold, and IDLE - 0 does not wrap

struct WarpSmem
{
	uint2 sq[SEQ_WORDS];      // .x = packed word i, .y = word i+1: any 16-base window is one LDS.64 + funnel shift
	uint2 st[SEQ_WORDS];
	uint2 vl[2 * VL_N];       // index k + KOFF: .x = furthest x on diagonal k, .y = packed anchor (one address per row)
};

struct Walk                   // one sequence seen as a forward walk
{
	const uint32_t* arr;
	uint32_t g0;              // array index of walk position 0
	uint32_t comp;            // 0 or 0xFFFFFFFF
	int len;
};

// 16 bases starting at base i of a staged operand, first base in the top two bits (the first mismatch is one CLZ)
__device__ __forceinline__ uint32_t seq16(const uint2* s, int i)
{
	const uint2 w = s[i >> 4];
	return __funnelshift_l(w.y, w.x, i << 1);
}

__device__ __forceinline__ int dtrunc_mul(double a, int b) { return (int)__dmul_rn(a, (double)b); }

}  // namespace

// A chain handed over by k_extend_lanes at the start of a block whose band outgrew the lane's ring.
struct ExtendSpill
{
	uint32_t item;
	int32_t qi, ti, cols, mats, qadv, tadv;
};

// Persistent warps: every warp pulls (candidate, direction) items from a global counter, so a
// long chain never pins three idle warps of its CTA.  spills == nullptr: items 0 .. 2 ntasks - 1 from their start;
// otherwise the nspills chains of the list, each resumed at its recorded block.
template <bool RESUME>
__global__ void __launch_bounds__(EXT_WARPS * 32, EXT_CTAS_PER_SM)
k_extend(const uint32_t* __restrict__ qfwd, const uint32_t* __restrict__ qrev, const int2* __restrict__ qoffsz, int qN,
         const uint32_t* __restrict__ sfwd, const uint32_t* __restrict__ srev, const int2* __restrict__ soffsz, int sN,
         const ExtendTask* __restrict__ tasks, size_t ntasks, ExtendHalf* __restrict__ halves,
         unsigned long long* __restrict__ block_counter, unsigned long long* __restrict__ work_counter,
         const ExtendSpill* __restrict__ spills, unsigned long long nspills)
{
	__shared__ WarpSmem smem[EXT_WARPS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	WarpSmem& S = smem[warp];
	unsigned nblocks = 0;
	unsigned ncells = 0;                                    // < 2^32 per warp and launch
	const unsigned long long nwork = RESUME ? nspills : 2ull * ntasks;

	for (;;) {
		unsigned long long item = 0;
		if (lane == 0) item = atomicAdd(work_counter, 1ull);
		item = __shfl_sync(FULL, item, 0);
		if (item >= nwork) break;
		ExtendSpill sp;
		sp.item = (uint32_t)item; sp.qi = sp.ti = sp.cols = sp.mats = sp.qadv = sp.tadv = 0;
		if (RESUME) { sp = spills[item]; item = sp.item; }
		const ExtendTask t = tasks[item >> 1];
		const int right = (int)(item & 1);

		const int2 qo = qoffsz[t.qread], so = soffsz[t.sread];
		Walk Q, T;
		if (right) {
			if (!t.qstrand) { Q.arr = qfwd; Q.g0 = (uint32_t)(qo.x + t.qstart); Q.comp = 0; }
			else { Q.arr = qrev; Q.g0 = (uint32_t)(qN - qo.x - qo.y + t.qstart); Q.comp = FULL; }
			Q.len = qo.y - t.qstart;
			T.arr = sfwd; T.g0 = (uint32_t)(so.x + t.sstart); T.comp = 0; T.len = so.y - t.sstart;
		} else {
			if (!t.qstrand) { Q.arr = qrev; Q.g0 = (uint32_t)(qN - qo.x - t.qstart); Q.comp = 0; }
			else { Q.arr = qfwd; Q.g0 = (uint32_t)(qo.x + qo.y - t.qstart); Q.comp = FULL; }
			Q.len = t.qstart;
			T.arr = srev; T.g0 = (uint32_t)(sN - so.x - t.sstart); T.comp = 0; T.len = t.sstart;
		}

		int qi = sp.qi, ti = sp.ti;
		int cols = sp.cols, mats = sp.mats, qadv = sp.qadv, tadv = sp.tadv;

		for (;;) {
			// ---- retrieve_next_aln_block
			const int qleft = Q.len - qi, tleft = T.len - ti;
			int qblk, tblk;
			bool last;
			if (qleft < 600 || tleft < 600) {
				int a = (int)__dadd_rn((double)tleft, __dmul_rn((double)tleft, 0.2));
				int b = (int)__dadd_rn((double)qleft, __dmul_rn((double)qleft, 0.2));
				qblk = min(qleft, a);
				tblk = min(tleft, b);
				last = true;
			} else { qblk = tblk = 500; last = false; }
			const int tol = dtrunc_mul(0.3, max(qblk, tblk));
			const int max_d = dtrunc_mul(.3, qblk + tblk);
			++nblocks;

			// ---- stage operands.  The reference zero-fills its V/U arrays per block; a row only ever
			// reads cells the previous row wrote (band edges take the single in-band neighbour), so
			// the one cell that must read as zero is V[1] at d = 0.
			__syncwarp();
			{
				const int qw = (qblk + 15) / 16 + 1, tw = (tblk + 15) / 16 + 1;
				for (int i = lane; i < qw; i += 32) {
					const uint32_t b0 = Q.g0 + (uint32_t)qi + 16u * i;
					S.sq[i] = make_uint2(__brev(ld_bases32(Q.arr, b0) ^ Q.comp), __brev(ld_bases32(Q.arr, b0 + 16u) ^ Q.comp));
				}
				for (int i = lane; i < tw; i += 32) {
					const uint32_t b0 = T.g0 + (uint32_t)ti + 16u * i;
					S.st[i] = make_uint2(__brev(ld_bases32(T.arr, b0) ^ T.comp), __brev(ld_bases32(T.arr, b0 + 16u) ^ T.comp));
				}
				if (lane == 0) S.vl[KOFF + 1] = make_uint2(0u, NO_ANCHOR);
			}
			__syncwarp();

			// ---- Align
			int min_k = 0, max_k = 0, best_m = -1;
			int last_min = 0, last_max = 0, rows = 0;
			bool aligned = false;
			int ex = 0, ey = 0;
			uint32_t ea = NO_ANCHOR;
			const uint2* sq = S.sq;
			const uint2* st = S.st;
			// one furthest-reaching cell: lane-local, reads the other-parity array, writes its own
			auto cell = [&](uint2* own, int j, int n, int k, uint32_t dbits, int& x, uint32_t& anc, bool& hit) {
				const uint2 lf = own[2 * j - 1], rt = own[2 * j + 1];
				if (j == 0 || (j != n - 1 && (int)lf.x < (int)rt.x)) { x = (int)rt.x; anc = rt.y; }
				else { x = (int)lf.x + 1; anc = lf.y; }
				int y = x - k;
				const int x1 = x;
				// x <= qblk and y <= tblk here (a cell that reached an end finished the block); the staged words cover one
				// window past either end.  `room` = matches left on this diagonal before a block end: a window's count is
				// capped by it, and the cell reached an end iff no room is left.
				int room = min(qblk - x, tblk - y);
				for (;;) {
					const uint32_t diff = seq16(sq, x) ^ seq16(st, y);
					const int m = min(__clz((int)diff) >> 1, room);      // matching bases in this 16-base window
					x += m; y += m; room -= m;
					if (m < 16) break;
				}
				hit = room == 0;
				if (x - x1 >= 4) anc = (uint32_t)x | ((uint32_t)y << 10) | dbits;
				own[2 * j] = make_uint2((uint32_t)x, anc);
				return x + y;
			};
			for (int d = 0; d < max_d; ++d) {
				if (max_k - min_k > 2 * tol) break;
				const int n = ((max_k - min_k) >> 1) + 1;
				ncells += (unsigned)n;
				uint2* own = &S.vl[min_k + KOFF];                      // own[2j] <-> diagonal min_k + 2j; neighbours own[2j -+ 1]
				const uint32_t dbits = (uint32_t)d << 20;
				// pass 0: diagonals 0..31 of the band (most rows have no other pass)
				int x0 = 0, u0 = IDLE;
				uint32_t a0 = NO_ANCHOR;
				bool h0 = false, h1 = false, hx = false;
				if (lane < n) u0 = cell(own, lane, n, min_k + 2 * lane, dbits, x0, a0, h0);
				int rowmax = __reduce_max_sync(FULL, u0);
				int x1 = 0, u1 = IDLE;
				uint32_t a1 = NO_ANCHOR;
				if (n > 32) {
					if (lane + 32 < n) u1 = cell(own, lane + 32, n, min_k + 2 * (lane + 32), dbits, x1, a1, h1);
					rowmax = max(rowmax, __reduce_max_sync(FULL, u1));
					for (int base = 64; base < n; base += 32) {
						const int j = base + lane;
						int xx = 0, uu = IDLE;
						uint32_t aa = NO_ANCHOR;
						bool hh = false;
						if (j < n) uu = cell(own, j, n, min_k + 2 * j, dbits, xx, aa, hh);
						hx = hx || hh;
						rowmax = max(rowmax, __reduce_max_sync(FULL, uu));
					}
				}
				__syncwarp();
				if (__any_sync(FULL, h0 || h1 || hx)) {
					// some cell reached a block end: the lowest such diagonal ends the block
					unsigned hm = __ballot_sync(FULL, h0);
					if (hm) {
						const int src = __ffs(hm) - 1;
						ex = __shfl_sync(FULL, x0, src); ey = __shfl_sync(FULL, u0, src) - ex; ea = __shfl_sync(FULL, a0, src);
						aligned = true;
					} else if (n > 32) {
						hm = __ballot_sync(FULL, h1);
						if (hm) {
							const int src = __ffs(hm) - 1;
							ex = __shfl_sync(FULL, x1, src); ey = __shfl_sync(FULL, u1, src) - ex; ea = __shfl_sync(FULL, a1, src);
							aligned = true;
						}
						for (int base = 64; base < n && !aligned; base += 32) {
							const int j = base + lane;
							uint2 c = make_uint2(0u, NO_ANCHOR);
							bool h = false;
							if (j < n) { c = own[2 * j]; const int yy = (int)c.x - (min_k + 2 * j); h = (int)c.x >= qblk || yy >= tblk; }
							hm = __ballot_sync(FULL, h);
							if (hm) {
								const int src = __ffs(hm) - 1;
								ex = __shfl_sync(FULL, (int)c.x, src);
								ey = ex - (min_k + 2 * (base + src));
								ea = __shfl_sync(FULL, c.y, src);
								aligned = true;
							}
						}
					}
					if (aligned) break;
				}
				best_m = max(best_m, rowmax);
				// re-band to the diagonals within `tol` of the best, widened by one
				const int thr = best_m - tol;
				int lo = 0x7fffffff, hi = -0x7fffffff;
				{
					const unsigned km = __ballot_sync(FULL, u0 >= thr);
					if (km) { lo = min_k + 2 * (__ffs(km) - 1); hi = min_k + 2 * (31 - __clz(km)); }
				}
				if (n > 32) {
					const unsigned km = __ballot_sync(FULL, u1 >= thr);
					if (km) { lo = min(lo, min_k + 2 * (32 + __ffs(km) - 1)); hi = max(hi, min_k + 2 * (32 + 31 - __clz(km))); }
					for (int base = 64; base < n; base += 32) {
						const int j = base + lane;
						bool keep = false;
						if (j < n) keep = 2 * (int)own[2 * j].x - (min_k + 2 * j) >= thr;
						const unsigned km2 = __ballot_sync(FULL, keep);
						if (km2) { lo = min(lo, min_k + 2 * (base + __ffs(km2) - 1)); hi = max(hi, min_k + 2 * (base + 31 - __clz(km2))); }
					}
				}
				last_min = min_k; last_max = max_k; ++rows;
				min_k = lo - 1; max_k = hi + 1;
			}
			if (!aligned && rows > 0) {
				// best (x+y) cell: first k of the last completed row that reaches best_m
				const int n = ((last_max - last_min) >> 1) + 1;
				const uint2* own = &S.vl[last_min + KOFF];
				for (int base = 0; base < n; base += 32) {
					const int j = base + lane;
					const int k = last_min + 2 * j;
					uint2 c = make_uint2(0u, NO_ANCHOR);
					bool is = false;
					if (j < n) {
						c = own[2 * j];
						is = 2 * (int)c.x - k == best_m;
					}
					const unsigned bm = __ballot_sync(FULL, is);
					if (bm) {
						const int src = __ffs(bm) - 1;
						const int bx = __shfl_sync(FULL, (int)c.x, src);
						const int bk = __shfl_sync(FULL, k, src);
						const uint32_t ba = __shfl_sync(FULL, c.y, src);
						if (bx > 0) { ex = bx; ey = bx - bk; ea = ba; }
						break;
					}
				}
			}

			// ---- trim_mismatch_end + chain bookkeeping (dw_in_one_direction)
			if (ea == NO_ANCHOR) break;
			const int ax = (int)(ea & 1023u), ay = (int)((ea >> 10) & 1023u), ad = (int)(ea >> 20);
			const int acols = (ax + ay + ad) >> 1, amat = (ax + ay - ad) >> 1;
			if (acols < 6) break;
			const bool full_map = (qblk - ex <= 20) || (tblk - ey <= 20);
			if (last || !full_map) {
				cols += acols; mats += amat; qadv += ax; tadv += ay;
				break;
			}
			cols += acols - 4; mats += amat - 4; qadv += ax - 4; tadv += ay - 4;
			qi += ax - 4; ti += ay - 4;
		}
		if (lane == 0) {
			ExtendHalf h;
			h.cols = cols; h.matches = mats; h.qadv = qadv; h.tadv = tadv;
			halves[item] = h;
		}
	}
	if (lane == 0 && block_counter && nblocks) {
		atomicAdd(block_counter, (unsigned long long)nblocks);
		atomicAdd(block_counter + 5, (unsigned long long)ncells);               // furthest-point cells (whole rows; the reference stops a row at the first end cell)
	}
}

// ------------------------------------------------------------------------------------------------------------------
// k_extend_pairs: a PAIR OF LANES per chain, 16 chains per warp.
//
// The warp-per-chain form spends ~134 warp-instructions on a row of ~30 cells (ncu, profiles/r2_launches_head.csv): a
// second pass for bands wider than 32 cells with most lanes idle, and a per-row tail of reductions, ballots and
// re-banding that a whole warp executes for one chain.  Here a chain belongs to two neighbouring lanes that walk the
// reference's own cell order (k ascending inside a row), two cells at a time, so one warp instruction serves 16
// chains and the per-row tail is shared by all of them:
//   * the row lives in a per-chain ring in shared memory.  Cell (d, k) sits at entry i = (k + d) / 2: its predecessors
//     (d-1, k-1) and (d-1, k+1) are entries i-1 and i of the previous row, so a row is updated IN PLACE -- each cell
//     loads entry i (16-bit furthest x + 32-bit anchor), gets entry i-1 from its neighbour lane by shuffle (the lane
//     that loaded it), and stores entry i; 80 entries per chain;
//   * both block operands are staged per chain as 2-bit words, first base in the top bits (a window of 16 bases is two
//     LDS + one funnel shift, the first mismatch one CLZ); the whole warp stages a chain's block (coalesced loads);
//     layouts are [entry][chain] with the entry parity in the bank number, so the two lanes of a pair (entries i and
//     i+1) and the 16 chains never collide on a bank; a chain needs 856 B, 16 warps (256 chains) fit an SM;
//   * the loop is flat: every iteration runs J cells per lane of the chain's current row (straight-line predicated
//     code, the J dependency chains interleave), then the transitions -- row end (re-banding = two short scans from the
//     tracked first candidate and from the last cell), block end (closed-form anchor bookkeeping), next item -- as
//     divergent tails.  Chains never wait for each other's rows, blocks or items.
//   * a row that would need more than the ring holds hands its chain, from the start of the current block, to the
//     warp-per-chain kernel (ExtendSpill); on CLR reads about one block in ten thousand.
// (A lane-per-chain form of the same loop ran at 8 warps per SM -- 27 KB of rings and operands per warp -- and was
// latency bound: 46 % issue-active, 521 ms against 406 ms for the warp-per-chain kernel; two lanes per chain halve the
// shared memory per warp.)
namespace {

constexpr int PR_WARPS = 16;              // warps per SM, one CTA: 16 x 13 696 B = 219 136 B of shared memory
constexpr int PR_RING = 80;               // ring entries per chain; a row of n cells needs n + 2 J <= PR_RING
constexpr int PR_SEQW = 47;               // words per staged operand: 719 bases = 45 words, the funnel word, the word a window at the very end touches
constexpr int PR_SEQ_WORDS = 2 * PR_SEQW * 16;
constexpr int PR_WARP_WORDS = PR_SEQ_WORDS + PR_RING * 16 + PR_RING * 8;
constexpr int PR_BIG = 0x3fffffff;
enum { PS_ITEM = 0, PS_BLOCK = 1, PS_ROW = 2, PS_DONE = 3 };

// 16 bases starting at base i of a chain's staged operand (s already points at the chain's column), first base on top
__device__ __forceinline__ uint32_t pair_win(const uint32_t* s, int i)
{
	const uint32_t* p = s + ((i >> 4) << 4);
	return __funnelshift_l(p[16], p[0], i << 1);
}

}  // namespace

template <int J>
__global__ void __launch_bounds__(PR_WARPS * 32, 1)
k_extend_pairs(const uint32_t* __restrict__ qfwd, const uint32_t* __restrict__ qrev, const int2* __restrict__ qoffsz, int qN,
               const uint32_t* __restrict__ sfwd, const uint32_t* __restrict__ srev, const int2* __restrict__ soffsz, int sN,
               const ExtendTask* __restrict__ tasks, unsigned long long nitems, ExtendHalf* __restrict__ halves,
               unsigned long long* __restrict__ counters, ExtendSpill* __restrict__ spills)
{
	extern __shared__ uint32_t pr_smem[];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int chain = lane >> 1, sub = lane & 1;
	const unsigned pmask = 3u << (lane & 30);               // the two lanes of this chain
	uint32_t* const wbase = pr_smem + warp * PR_WARP_WORDS;
	uint32_t* const sq = wbase + chain;                     // word w of the chain's query window: sq[16 w]
	uint32_t* const st = sq + PR_SEQW * 16;
	uint32_t* const va = wbase + PR_SEQ_WORDS + chain;      // anchor of ring entry e: va[16 e]
	uint16_t* const vx = (uint16_t*)(wbase + PR_SEQ_WORDS + PR_RING * 16) + 2 * chain;   // furthest x of entry e: vx[32 (e >> 1) + (e & 1)]

	// chain state, identical in both lanes of the pair unless noted
	int state = PS_ITEM;
	unsigned item = 0;
	int qsel = 0, tsel = 0, qlen = 0, tlen = 0;             // walks (see k_extend): which array, first base, complement
	uint32_t qg0 = 0, tg0 = 0, qcomp = 0;
	int qi = 0, ti = 0, cols = 0, mats = 0, qadv = 0, tadv = 0;
	int qblk = 0, tblk = 0, tol = 0, max_d = 0;
	bool last = false;
	int d = 0, min_k = 0, max_k = 0, k = 0, ro = 0, ro_lo = 0, best_m = -1;
	uint32_t dbits = 0;
	// per lane: best own cell (first one reaching the lane's maximum, in row order), first own re-banding candidate,
	// the entry left of the lane's next cell
	int obu = -1, obx = 0, obk = 0, lo_c = PR_BIG, cx = 0;
	uint32_t oba = NO_ANCHOR, ca = NO_ANCHOR;
	unsigned nblocks = 0, ncells = 0;

	for (;;) {
		// ---- next (candidate, direction) for chains without one
		const unsigned need = __ballot_sync(FULL, state == PS_ITEM) & 0x55555555u;
		if (need) {
			unsigned long long base = 0;
			if (lane == 0) base = atomicAdd(counters + 4, (unsigned long long)__popc(need));
			base = __shfl_sync(FULL, base, 0);
			if (state == PS_ITEM) {
				const unsigned long long it = base + (unsigned)__popc(need & ((1u << (lane & 30)) - 1u));
				if (it >= nitems) state = PS_DONE;
				else {
					item = (unsigned)it;
					const ExtendTask t = tasks[it >> 1];
					const int2 qo = qoffsz[t.qread], so = soffsz[t.sread];
					if (it & 1) {
						if (!t.qstrand) { qsel = 0; qg0 = (uint32_t)(qo.x + t.qstart); qcomp = 0; }
						else { qsel = 1; qg0 = (uint32_t)(qN - qo.x - qo.y + t.qstart); qcomp = FULL; }
						qlen = qo.y - t.qstart;
						tsel = 0; tg0 = (uint32_t)(so.x + t.sstart); tlen = so.y - t.sstart;
					} else {
						if (!t.qstrand) { qsel = 1; qg0 = (uint32_t)(qN - qo.x - t.qstart); qcomp = 0; }
						else { qsel = 0; qg0 = (uint32_t)(qo.x + qo.y - t.qstart); qcomp = FULL; }
						qlen = t.qstart;
						tsel = 1; tg0 = (uint32_t)(sN - so.x - t.sstart); tlen = t.sstart;
					}
					qi = ti = 0; cols = mats = qadv = tadv = 0;
					state = PS_BLOCK;
				}
			}
		}
		if (__all_sync(FULL, state == PS_DONE)) break;

		bool endblk = false;
		int ex = 0, ey = 0;
		uint32_t ea = NO_ANCHOR;

		// ---- retrieve_next_aln_block + staging for chains at a block boundary
		const unsigned nbm = __ballot_sync(FULL, state == PS_BLOCK) & 0x55555555u;
		if (nbm) {
			int qw = 0, tw = 0;
			if (state == PS_BLOCK) {
				const int qleft = qlen - qi, tleft = tlen - ti;
				if (qleft < 600 || tleft < 600) {
					const int a = (int)__dadd_rn((double)tleft, __dmul_rn((double)tleft, 0.2));
					const int b = (int)__dadd_rn((double)qleft, __dmul_rn((double)qleft, 0.2));
					qblk = min(qleft, a);
					tblk = min(tleft, b);
					last = true;
				} else { qblk = tblk = 500; last = false; }
				tol = dtrunc_mul(0.3, max(qblk, tblk));
				max_d = dtrunc_mul(.3, qblk + tblk);
				nblocks += (unsigned)(sub == 0);
				qw = (qblk + 15) / 16 + 1; tw = (tblk + 15) / 16 + 1;
			}
			for (unsigned m = nbm; m; m &= m - 1) {
				const int t = __ffs(m) - 1;                     // even lane of the chain
				const int sel = __shfl_sync(FULL, qsel | (tsel << 1), t);
				const uint32_t qb = __shfl_sync(FULL, qg0 + (uint32_t)qi, t), tb = __shfl_sync(FULL, tg0 + (uint32_t)ti, t);
				const uint32_t cp = __shfl_sync(FULL, qcomp, t);
				const int nq = __shfl_sync(FULL, qw, t), nt = __shfl_sync(FULL, tw, t);
				const uint32_t* qa = (sel & 1) ? qrev : qfwd;
				const uint32_t* ta = (sel & 2) ? srev : sfwd;
				uint32_t* dst = wbase + (t >> 1);
				for (int w = lane; w < nq; w += 32) dst[w * 16] = __brev(ld_bases32(qa, qb + 16u * w) ^ cp);
				for (int w = lane; w < nt; w += 32) dst[(PR_SEQW + w) * 16] = __brev(ld_bases32(ta, tb + 16u * w));
			}
			__syncwarp();
			if (state == PS_BLOCK) {
				// the one cell that must read as zero: V[1] at d = 0 (see k_extend)
				if (sub == 0) { vx[0] = 0; va[0] = NO_ANCHOR; }
				d = 0; dbits = 0; min_k = max_k = k = 0; ro = ro_lo = 0;
				best_m = -1; obu = -1; obx = 0; obk = 0; oba = NO_ANCHOR; lo_c = PR_BIG;
				if (max_d > 0) state = PS_ROW;
				else endblk = true;                 // no row at all: nothing aligned, the chain ends here
			}
			__syncwarp();
		}

		// ---- J cells per lane of the chain's current row: in step c the pair takes cells 2c (even lane) and 2c + 1 (odd
		// lane) after position k.  Straight-line predicated code; lanes without a cell compute on position 0 and store nothing.
		const bool rowing = state == PS_ROW;
		__syncwarp();                                       // ring entries the neighbour lane stored in the last step
		int xs[J], ys[J], x1s[J], kcs[J], slots[J];
		uint32_t as[J];
		bool val[J];
		unsigned cont = 0, hits = 0;
#pragma unroll
		for (int c = 0; c < J; ++c) {
			const int kc = k + 2 * (2 * c + sub);
			const bool v = rowing && kc <= max_k;
			int sl = ro + 2 * c + sub;
			if (sl >= PR_RING) sl -= PR_RING;
			const int rx = (int)vx[((sl >> 1) << 5) + (sl & 1)];
			const uint32_t ra = va[sl << 4];
			// entry left of this cell: the even lane's entry for the odd lane, the odd lane's previous entry for the even lane
			const int px = __shfl_xor_sync(FULL, rx, 1);
			const uint32_t pa = __shfl_xor_sync(FULL, ra, 1);
			const int lx = sub ? px : cx;
			const uint32_t la = sub ? pa : ca;
			cx = px; ca = pa;
			const bool take_right = kc == min_k || (kc != max_k && lx < rx);
			int x = take_right ? rx : lx + 1;
			const uint32_t a = take_right ? ra : la;
			int y = x - kc;
			if (!v) { x = 0; y = 0; }
			x1s[c] = x;
			const uint32_t diff = pair_win(sq, x) ^ pair_win(st, y);
			const int room = min(qblk - x, tblk - y);           // matches left on this diagonal before a block end
			const int m = min(__clz((int)diff) >> 1, room);
			x += m; y += m;
			if (v && m == 16) cont |= 1u << c;                   // a whole window matched and the block goes on
			if (v && m == room) hits |= 1u << c;                 // reached a block end
			xs[c] = x; ys[c] = y; as[c] = a; kcs[c] = kc; slots[c] = sl; val[c] = v;
		}
		if (cont) {
			// a snake longer than one window (rare): finish it, 16 bases per step
#pragma unroll
			for (int c = 0; c < J; ++c) {
				if (cont & (1u << c)) {
					int x = xs[c], y = ys[c];
					while (x < qblk && y < tblk) {
						const uint32_t diff = pair_win(sq, x) ^ pair_win(st, y);
						int m = __clz((int)diff) >> 1;
						m = min(m, min(qblk - x, tblk - y));
						x += m; y += m;
						if (m < 16) break;
					}
					xs[c] = x; ys[c] = y;
					if (x >= qblk || y >= tblk) hits |= 1u << c;
				}
			}
		}
		int lane_best = max(best_m, obu);
#pragma unroll
		for (int c = 0; c < J; ++c) {
			const int x = xs[c], y = ys[c];
			const uint32_t a = x - x1s[c] >= 4 ? ((uint32_t)x | ((uint32_t)y << 10) | dbits) : as[c];
			as[c] = a;
			if (val[c]) {                                        // two predicated stores
				vx[((slots[c] >> 1) << 5) + (slots[c] & 1)] = (uint16_t)x;
				va[slots[c] << 4] = a;
			}
			const int u = x + y;
			const bool better = val[c] && u > obu;
			obu = better ? u : obu;
			obx = better ? x : obx;
			oba = better ? a : oba;
			obk = better ? kcs[c] : obk;
			lane_best = max(lane_best, obu);
			// first own cell within `tol` of the best seen so far: the row's final threshold is at least this one
			lo_c = min(lo_c, (val[c] && u >= lane_best - tol) ? kcs[c] : PR_BIG);
		}
		bool hit = false;
		// everything the two lanes of a pair must tell each other at a row end is exchanged here, in converged code with the
		// full mask: a *_sync with a different sub-mask per pair would split the warp into 16 paths
		const int p_obu = __shfl_xor_sync(FULL, obu, 1), p_lo = __shfl_xor_sync(FULL, lo_c, 1);
		if (__ballot_sync(FULL, hits != 0) & pmask) {
			// the first cell of the row (lowest k) that reached a block end finishes the block: the reference stops the row there
			int hk = PR_BIG, hx = 0, hy = 0;
			uint32_t ha = NO_ANCHOR;
#pragma unroll
			for (int c = J - 1; c >= 0; --c)
				if (hits & (1u << c)) { hx = xs[c]; hy = ys[c]; ha = as[c]; hk = kcs[c]; }
			const int ok = __shfl_xor_sync(pmask, hk, 1), ox = __shfl_xor_sync(pmask, hx, 1), oy = __shfl_xor_sync(pmask, hy, 1);
			const uint32_t oa = __shfl_xor_sync(pmask, ha, 1);
			if (ok < hk) { hk = ok; hx = ox; hy = oy; ha = oa; }
			ex = hx; ey = hy; ea = ha;
			hit = true;
			if (sub == 0) ncells += (unsigned)((hk - min_k) >> 1) + 1u;
		}
		k += 4 * J;
		ro += 2 * J;
		if (ro >= PR_RING) ro -= PR_RING;
		__syncwarp();                                       // the scans below read what both lanes just stored

		// ---- row end: re-band to the diagonals within `tol` of the best, widened by one
		if (rowing && !hit && k > max_k) {
			if (sub == 0) ncells += (unsigned)((max_k - min_k) >> 1) + 1u;
			best_m = max(best_m, max(obu, p_obu));
			lo_c = min(lo_c, p_lo);
			const int thr = best_m - tol;
			int hk = max_k;
			int hs = ro_lo + ((max_k - min_k) >> 1);
			if (hs >= PR_RING) hs -= PR_RING;
			while (hk > min_k && 2 * (int)vx[((hs >> 1) << 5) + (hs & 1)] - hk < thr) { hk -= 2; hs = hs ? hs - 1 : PR_RING - 1; }
			int lk = lo_c;
			int ls = ro_lo + ((lo_c - min_k) >> 1);
			if (ls >= PR_RING) ls -= PR_RING;
			while (lk < hk && 2 * (int)vx[((ls >> 1) << 5) + (ls & 1)] - lk < thr) { lk += 2; ls = ls + 1 == PR_RING ? 0 : ls + 1; }
			++d; dbits += 1u << 20;
			min_k = lk - 1; max_k = hk + 1; k = min_k; ro = ro_lo = ls; lo_c = PR_BIG;
			if (d >= max_d || max_k - min_k > 2 * tol) {
				// never aligned: the best (x + y) cell, first one in row order (its row is the last one: the row maximum grows
				// with every row, so equal sums can only meet inside that row and the lower diagonal came first)
				endblk = true;
				const int pu = __shfl_xor_sync(pmask, obu, 1), pk = __shfl_xor_sync(pmask, obk, 1), px = __shfl_xor_sync(pmask, obx, 1);
				const uint32_t pa = __shfl_xor_sync(pmask, oba, 1);
				int bu = obu, bx = obx;
				uint32_t ba = oba;
				if (pu > obu || (pu == obu && pk < obk)) { bu = pu; bx = px; ba = pa; }
				if (bx > 0) { ex = bx; ey = bu - bx; ea = ba; }
			} else if (((max_k - min_k) >> 1) + 1 + 2 * J > PR_RING) {
				// the band outgrew the ring: this block and the rest of the chain go to the warp-per-chain kernel
				if (sub == 0) {
					const unsigned long long si = atomicAdd(counters + 6, 1ull);
					ExtendSpill sp;
					sp.item = item; sp.qi = qi; sp.ti = ti; sp.cols = cols; sp.mats = mats; sp.qadv = qadv; sp.tadv = tadv;
					spills[si] = sp;
					--nblocks;
				}
				state = PS_ITEM;
			}
		}
		endblk = endblk || hit;

		// ---- trim_mismatch_end + chain bookkeeping (dw_in_one_direction), as in k_extend
		if (endblk) {
			bool chain_end = true;
			if (ea != NO_ANCHOR) {
				const int ax = (int)(ea & 1023u), ay = (int)((ea >> 10) & 1023u), ad = (int)(ea >> 20);
				const int acols = (ax + ay + ad) >> 1, amat = (ax + ay - ad) >> 1;
				if (acols >= 6) {
					const bool full_map = (qblk - ex <= 20) || (tblk - ey <= 20);
					if (last || !full_map) { cols += acols; mats += amat; qadv += ax; tadv += ay; }
					else {
						cols += acols - 4; mats += amat - 4; qadv += ax - 4; tadv += ay - 4;
						qi += ax - 4; ti += ay - 4;
						chain_end = false;
					}
				}
			}
			if (chain_end) {
				if (sub == 0) {
					ExtendHalf h;
					h.cols = cols; h.matches = mats; h.qadv = qadv; h.tadv = tadv;
					halves[item] = h;
				}
				state = PS_ITEM;
			} else state = PS_BLOCK;
		}
	}
	nblocks = __reduce_add_sync(FULL, nblocks);
	ncells = __reduce_add_sync(FULL, ncells);
	if (lane == 0) {
		if (nblocks) atomicAdd(counters, (unsigned long long)nblocks);
		if (ncells) atomicAdd(counters + 5, (unsigned long long)ncells);
	}
}

static int extend_mode()
{
	// MECAT_B200_EXTEND = warp (default): everything through the warp-per-chain kernel; pairs1 / pairs2 / pairs3: the
	// lane-pair kernel with that many cells per lane and step.  Measured on BASELINE configs[1] (profiles/README.md,
	// round 2): warp 406 ms, pairs3 436 ms with one launch per tile (541 ms with the 8-chunk pipeline: 256 chains per SM
	// leave a long tail per launch).  The pair kernel needs 15 % fewer warp-instructions but 74 % of its issue slots go to
	// the half-rate integer ALU pipe, so it stays opt-in until its per-iteration tail is leaner.
	static int mode = -1;
	if (mode < 0) {
		mode = 0;
		if (const char* e = getenv("MECAT_B200_EXTEND")) {
			if (!strcmp(e, "pairs1")) mode = 1;
			else if (!strcmp(e, "pairs2")) mode = 2;
			else if (!strcmp(e, "pairs3")) mode = 3;
		}
	}
	return mode;
}

int extend_launch(Ctx* c, const DVolume* q, const DVolume* s, const ExtendTask* d_tasks, size_t ntasks,
                  ExtendHalf* d_halves)
{
	if (!ntasks) return 0;
	unsigned long long* d_counter = c->d_counters;          // [0] block statistics, [4] work queue head, [5] cells, [6] spills
	MB_CUDA(c, cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), c->stream));
	MB_CUDA(c, cudaMemsetAsync(d_counter + 4, 0, 3 * sizeof(unsigned long long), c->stream));
	const size_t items = 2 * ntasks;
	const int mode = extend_mode();
	ExtendSpill* d_spills = nullptr;
	unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	auto wide = [&](const ExtendSpill* spills, unsigned long long nspills) {
		const size_t work = spills ? (size_t)nspills : items;
		const size_t want = (work + EXT_WARPS - 1) / EXT_WARPS;
		const unsigned grid = (unsigned)std::min<size_t>(want, (size_t)c->sm_count * EXT_CTAS_PER_SM);
		KScope ks(c, MECAT_K_EXTEND);
		if (spills)
			k_extend<true><<<grid, EXT_WARPS * 32, 0, c->stream>>>(q->fwd, q->rev, q->offsz, q->num_bases, s->fwd, s->rev, s->offsz,
			                                                       s->num_bases, d_tasks, ntasks, d_halves, d_counter, d_counter + 4,
			                                                       spills, nspills);
		else
			k_extend<false><<<grid, EXT_WARPS * 32, 0, c->stream>>>(q->fwd, q->rev, q->offsz, q->num_bases, s->fwd, s->rev, s->offsz,
			                                                        s->num_bases, d_tasks, ntasks, d_halves, d_counter, d_counter + 4,
			                                                        nullptr, 0);
	};
	auto body = [&]() -> int {
		if (mode == 0) {
			wide(nullptr, 0);
			MB_CUDA(c, cudaGetLastError());
			MB_CUDA(c, cudaMemcpyAsync(h, d_counter, sizeof h, cudaMemcpyDeviceToHost, c->stream));
			MB_CUDA(c, cudaStreamSynchronize(c->stream));
			c->stats.num_extend_cells += (int64_t)h[5];
		} else {
			MB_CUDA(c, c->alloc(&d_spills, items));
			const size_t smem = (size_t)PR_WARPS * PR_WARP_WORDS * sizeof(uint32_t);
			const size_t want = (items + PR_WARPS * 16 - 1) / (PR_WARPS * 16);
			const unsigned grid = (unsigned)std::min<size_t>(want, (size_t)c->sm_count);
			auto launch = [&](auto kern) -> cudaError_t {
				cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
				if (e != cudaSuccess) return e;
				KScope ks(c, MECAT_K_EXTEND);
				kern<<<grid, PR_WARPS * 32, smem, c->stream>>>(q->fwd, q->rev, q->offsz, q->num_bases, s->fwd, s->rev, s->offsz,
				                                               s->num_bases, d_tasks, (unsigned long long)items, d_halves, d_counter,
				                                               d_spills);
				return cudaGetLastError();
			};
			if (mode == 1) MB_CUDA(c, launch(k_extend_pairs<1>));
			else if (mode == 3) MB_CUDA(c, launch(k_extend_pairs<3>));
			else MB_CUDA(c, launch(k_extend_pairs<2>));
			MB_CUDA(c, cudaMemcpyAsync(h, d_counter, sizeof h, cudaMemcpyDeviceToHost, c->stream));
			MB_CUDA(c, cudaStreamSynchronize(c->stream));
			if (h[6]) {
				MB_CUDA(c, cudaMemsetAsync(d_counter + 4, 0, sizeof(unsigned long long), c->stream));
				wide(d_spills, h[6]);
				MB_CUDA(c, cudaGetLastError());
				c->stats.num_extend_spills += (int64_t)h[6];
			}
			c->stats.num_extend_cells += (int64_t)h[5];
		}
		MB_CUDA(c, cudaMemcpyAsync(h, d_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
		MB_CUDA(c, cudaStreamSynchronize(c->stream));
		c->resolve_timers();
		c->stats.num_extend_blocks += (int64_t)h[0];
		return 0;
	};
	const int rc = body();
	c->dfree(d_spills);
	return rc;
}

}  // namespace mb
