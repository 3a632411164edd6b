// mecat_b200/csrc/align.cu -- batched O(nd) gapped extension WITH alignment strings.
//
// Two flavours of the same recurrence (extend.cu has the string-free pw fast path):
//   policy 0  pw / ref : DiffAligner::go + query/target_mapped_string
//                        src/common/diff_gapalign.cpp:40-349, gapalign.cpp:10-67
//                        (consumer: mecat2ref extend_candidate, src/mecat2ref/mecat2ref_aux.cpp:123-183)
//   policy 1  cns      : ns_banded_sw::Align / dw_in_one_direction / dw / GetAlignment
//                        src/mecat2cns/dw.cpp:146-553
//                        (consumer: consensus_one_read_can_pacbio, src/mecat2cns/mecat_correction.cpp:389-450)
//
// One warp owns one (task, direction) chain.  The forward pass is the one of extend.cu; in
// addition every row stores, per 32 diagonals, the ballot of "came from diagonal k+1" (one word)
// and its lowest diagonal.  When a block is accepted the warp walks those words back from the end
// cell to recover the path (one bit per edit), then replays the path forwards: each edit emits one
// gap column, each snake is re-extended 32 bases per ballot and its match columns are written lane
// parallel.  Columns go to a per-(task, direction) slot in walking order; k_aln_assemble reverses
// the left part, appends the right part, applies the flavour's trimming and packs the strings.
#include "common.cuh"
#include "xdrop_core.cuh"

#include <algorithm>

namespace mb {

namespace {

constexpr int AL_WARPS = 4;
// Row budget of a block: max_d is 0.3 (q + t) <= 395 for the pw / ref flavour and 2 err (q + t) for the cns flavour -- 360 at
// err 0.15 (pacbio), 480 at err 0.20 (nanopore consensus, mecat_correction.cpp:487), which gets the WIDE instance.
template <bool WIDE> struct AlDims
{
	static constexpr int KOFF = WIDE ? 484 : 404;      // even, > max_d of any accepted block
	static constexpr int VL_N = KOFF + 4;
	static constexpr int MAXROWS = WIDE ? 480 : 400;
};
constexpr int SEQ_WORDS = 48;
constexpr int DIRW = 12;                  // ballot words per row (band <= 2*180+1 diagonals)
constexpr uint32_t NO_ANCHOR = 0xFFFFFFFFu;
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int IDLE = -0x7fffffff - 1;     // x + y of a lane without a cell (see extend.cu)

template <bool WIDE> struct WarpSmem
{
	static constexpr int VL_N = AlDims<WIDE>::VL_N, MAXROWS = AlDims<WIDE>::MAXROWS;
	uint2 sq[SEQ_WORDS];
	uint2 st[SEQ_WORDS];
	uint2 vl[2 * VL_N];       // index k + KOFF: .x = furthest x on diagonal k, .y = packed anchor (interleaved parities, as in extend.cu)
	uint32_t path[(MAXROWS + 31) / 32 + 1];   // recovered path: bit d = edit d came from diagonal k+1
};

struct Walk { const uint32_t* arr; uint32_t g0; uint32_t comp; int len; };

__device__ __forceinline__ uint32_t seq16(const uint2* s, int i)
{
	const uint2 w = s[i >> 4];
	return __funnelshift_r(w.x, w.y, (i & 15) << 1);
}
__device__ __forceinline__ int base_at(const uint2* s, int i) { return (int)((s[i >> 4].x >> ((i & 15) << 1)) & 3u); }
__device__ __forceinline__ int dtrunc_mul(double a, int b) { return (int)__dmul_rn(a, (double)b); }

}  // namespace

template <int POLICY, bool WIDE>
__global__ void __launch_bounds__(AL_WARPS * 32)
k_align(const uint32_t* __restrict__ qfwd, const uint32_t* __restrict__ qrev, const int2* __restrict__ qoffsz, int qN,
        const uint32_t* __restrict__ sfwd, const uint32_t* __restrict__ srev, const int2* __restrict__ soffsz, int sN,
        const AlignTask* __restrict__ tasks, size_t ntasks, AlnSlot* __restrict__ slots, char* __restrict__ colq,
        char* __restrict__ colt, uint32_t* __restrict__ dir_scratch, short* __restrict__ min_scratch, double err,
        unsigned long long* __restrict__ work_counter)
{
	constexpr int KOFF = AlDims<WIDE>::KOFF, MAXROWS = AlDims<WIDE>::MAXROWS;
	__shared__ WarpSmem<WIDE> smem[AL_WARPS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	WarpSmem<WIDE>& S = smem[warp];
	const size_t gw = (size_t)blockIdx.x * AL_WARPS + warp;
	uint32_t* rowdir = dir_scratch + gw * (size_t)(MAXROWS * DIRW);
	short* rowmin = min_scratch + gw * (size_t)MAXROWS;

	for (;;) {
		unsigned long long item = 0;
		if (lane == 0) item = atomicAdd(work_counter, 1ull);
		item = __shfl_sync(FULL, item, 0);
		if (item >= 2 * ntasks) break;
		const AlignTask t = tasks[item >> 1];
		const int right = (int)(item & 1);
		const int2 qo = qoffsz[t.qread];
		int2 so = soffsz[t.sread];
		if (t.swin_len > 0) { so.x += t.swin_off; so.y = t.swin_len; }    // subject window (mecat2ref)
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
		AlnSlot slot = slots[item];
		char* oq = colq + slot.off;
		char* ot = colt + slot.off;

		int qi = 0, ti = 0;
		int cols = 0, mats = 0, qadv = 0, tadv = 0, overflow = 0;

		for (;;) {
			// ---- block geometry
			const int qleft = Q.len - qi, tleft = T.len - ti;
			int qblk, tblk, tol, max_d;
			bool last;
			if (POLICY == 0) {
				if (qleft < 600 || tleft < 600) {
					int a = (int)__dadd_rn((double)tleft, __dmul_rn((double)tleft, 0.2));
					int b = (int)__dadd_rn((double)qleft, __dmul_rn((double)qleft, 0.2));
					qblk = min(qleft, a); tblk = min(tleft, b); last = true;
				} else { qblk = tblk = 500; last = false; }
				tol = dtrunc_mul(0.3, max(qblk, tblk));
				max_d = dtrunc_mul(.3, qblk + tblk);
			} else {
				const int ext = min(qleft, tleft);
				if (ext > 600) { qblk = tblk = 500; last = false; }
				else { qblk = tblk = ext; last = true; }
				tol = dtrunc_mul(0.3, qblk);
				max_d = (int)__dmul_rn(__dmul_rn(2.0, err), (double)(qblk + tblk));
			}

			__syncwarp();
			{
				const int qw = (qblk + 15) / 16 + 1, tw = (tblk + 15) / 16 + 1;
				for (int i = lane; i < qw; i += 32) {
					const uint32_t b0 = Q.g0 + (uint32_t)qi + 16u * i;
					S.sq[i] = make_uint2(ld_bases32(Q.arr, b0) ^ Q.comp, ld_bases32(Q.arr, b0 + 16u) ^ Q.comp);
				}
				for (int i = lane; i < tw; i += 32) {
					const uint32_t b0 = T.g0 + (uint32_t)ti + 16u * i;
					S.st[i] = make_uint2(ld_bases32(T.arr, b0) ^ T.comp, ld_bases32(T.arr, b0 + 16u) ^ T.comp);
				}
				if (lane == 0) S.vl[KOFF + 1] = make_uint2(0u, NO_ANCHOR);
			}
			__syncwarp();

			// ---- forward pass (see extend.cu), recording the direction ballots of every row
			int min_k = 0, max_k = 0, best_m = -1;
			int last_min = 0, last_max = 0, rows = 0;
			bool aligned = false;
			int ex = 0, ey = 0, ed = 0;
			uint32_t ea = NO_ANCHOR;
			const uint2* sq = S.sq;
			const uint2* st = S.st;
			// one furthest-reaching cell (extend.cu's): lane-local, reads the other parity's neighbours, writes its own slot
			auto cell = [&](uint2* own, int j, int n, int k, uint32_t dbits, int& x, uint32_t& anc, bool& from_right) {
				const uint2 lf = own[2 * j - 1], rt = own[2 * j + 1];
				from_right = (j == 0 || (j != n - 1 && (int)lf.x < (int)rt.x));
				if (from_right) { x = (int)rt.x; anc = rt.y; }
				else { x = (int)lf.x + 1; anc = lf.y; }
				int y = x - k;
				const int x1 = x;
				for (;;) {
					const uint32_t diff = seq16(sq, x) ^ seq16(st, y);
					const int m = __clz(__brev(diff)) >> 1;
					x += m; y += m;
					if (m < 16 || x >= qblk || y >= tblk) break;
				}
				const int over = max(max(x - qblk, y - tblk), 0);
				x -= over; y -= over;
				if (x - x1 >= 4) anc = (uint32_t)x | ((uint32_t)y << 10) | dbits;
				own[2 * j] = make_uint2((uint32_t)x, anc);
				return x + y;
			};
			for (int d = 0; d < max_d; ++d) {
				if (max_k - min_k > 2 * tol) break;
				const int n = ((max_k - min_k) >> 1) + 1;
				uint2* own = &S.vl[min_k + KOFF];
				const uint32_t dbits = (uint32_t)d << 20;
				uint32_t* dirrow = rowdir + d * DIRW;
				// pass 0 and 1 keep their cells in registers (most rows have no further pass)
				int x0 = 0, u0 = IDLE, x1 = 0, u1 = IDLE;
				uint32_t a0 = NO_ANCHOR, a1 = NO_ANCHOR;
				bool fr = false;
				if (lane < n) u0 = cell(own, lane, n, min_k + 2 * lane, dbits, x0, a0, fr);
				unsigned dm = __ballot_sync(FULL, fr);
				if (lane == 0) { rowmin[d] = (short)min_k; dirrow[0] = dm; }
				int rowmax = __reduce_max_sync(FULL, u0);
				if (n > 32) {
					fr = false;
					if (lane + 32 < n) u1 = cell(own, lane + 32, n, min_k + 2 * (lane + 32), dbits, x1, a1, fr);
					dm = __ballot_sync(FULL, fr);
					if (lane == 0) dirrow[1] = dm;
					rowmax = max(rowmax, __reduce_max_sync(FULL, u1));
					for (int base = 64; base < n; base += 32) {
						const int j = base + lane;
						int xx = 0, uu = IDLE;
						uint32_t aa = NO_ANCHOR;
						fr = false;
						if (j < n) uu = cell(own, j, n, min_k + 2 * j, dbits, xx, aa, fr);
						dm = __ballot_sync(FULL, fr);
						if (lane == 0) dirrow[base >> 5] = dm;
						rowmax = max(rowmax, __reduce_max_sync(FULL, uu));
					}
				}
				__syncwarp();
				// x >= qblk needs x + y >= 2 qblk - k, y >= tblk needs x + y >= 2 tblk + k
				if (rowmax >= min(2 * qblk - max_k, 2 * tblk + min_k)) {
					unsigned hm = __ballot_sync(FULL, x0 >= qblk || u0 - x0 >= tblk);
					if (hm) {
						const int src = __ffs(hm) - 1;
						ex = __shfl_sync(FULL, x0, src); ey = __shfl_sync(FULL, u0, src) - ex; ea = __shfl_sync(FULL, a0, src);
						aligned = true;
					} else if (n > 32) {
						hm = __ballot_sync(FULL, x1 >= qblk || u1 - x1 >= tblk);
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
					if (aligned) { ed = d; break; }
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
			if (POLICY == 0 && !aligned && rows > 0) {
				// pw flavour only: best (x+y) cell of the last completed row (diff_gapalign.cpp:197-216)
				const int n = ((last_max - last_min) >> 1) + 1;
				const uint2* own = &S.vl[last_min + KOFF];
				for (int base = 0; base < n; base += 32) {
					const int j = base + lane;
					const int k = last_min + 2 * j;
					uint2 c = make_uint2(0u, NO_ANCHOR);
					bool is = false;
					if (j < n) { c = own[2 * j]; is = 2 * (int)c.x - k == best_m; }
					const unsigned bm = __ballot_sync(FULL, is);
					if (bm) {
						const int src = __ffs(bm) - 1;
						const int bx = __shfl_sync(FULL, (int)c.x, src);
						const int bk = __shfl_sync(FULL, k, src);
						const uint32_t ba = __shfl_sync(FULL, c.y, src);
						if (bx > 0) { ex = bx; ey = bx - bk; ea = ba; ed = rows - 1; }
						break;
					}
				}
			}

			// ---- which prefix of the block's alignment is kept, and does the chain go on
			int kx, ky, kd;          // end cell of the kept path
			int drop;                // trailing match columns of that cell's snake that are not emitted
			bool go_on;
			if (POLICY == 0) {
				if (ea == NO_ANCHOR) break;
				const int ax = (int)(ea & 1023u), ay = (int)((ea >> 10) & 1023u), ad = (int)(ea >> 20);
				if (((ax + ay + ad) >> 1) < 6) break;
				const bool full_map = (qblk - ex <= 20) || (tblk - ey <= 20);
				go_on = !(last || !full_map);
				kx = ax; ky = ay; kd = ad; drop = go_on ? 4 : 0;
			} else {
				if (!aligned) break;
				if (!last) {
					if (ea == NO_ANCHOR) break;
					const int ax = (int)(ea & 1023u), ay = (int)((ea >> 10) & 1023u), ad = (int)(ea >> 20);
					if (ax == 4) break;
					kx = ax; ky = ay; kd = ad; drop = 4; go_on = true;
				} else {
					if (ex == 0) break;
					kx = ex; ky = ey; kd = ed; drop = 0; go_on = false;
				}
			}
			const int kcols = ((kx + ky + kd) >> 1) - drop;

			// ---- walk the direction ballots back from (kd, kx - ky) to row 0
			{
				for (int i = lane; i < (MAXROWS + 31) / 32 + 1; i += 32) S.path[i] = 0u;
				__syncwarp();
				int k = kx - ky;
				for (int dhi = kd; dhi >= 1; dhi -= 32) {
					// lanes fetch the rows dhi, dhi-1, ... (first two ballot words and the row's lowest diagonal)
					const int row = dhi - lane;
					uint32_t w0 = 0, w1 = 0;
					int rmin = 0;
					if (row >= 1) { w0 = rowdir[row * DIRW]; w1 = rowdir[row * DIRW + 1]; rmin = rowmin[row]; }
					uint32_t bits = 0;
					const int steps = min(32, dhi);
					for (int s = 0; s < steps; ++s) {
						const int rm = __shfl_sync(FULL, rmin, s);
						const uint32_t a = __shfl_sync(FULL, w0, s), b = __shfl_sync(FULL, w1, s);
						const int j = (k - rm) >> 1;
						uint32_t w = j < 32 ? a : b;
						if (j >= 64) w = rowdir[(dhi - s) * DIRW + (j >> 5)];     // wide band: fetch the word directly
						const uint32_t bit = (w >> (j & 31)) & 1u;
						bits |= bit << s;
						k += bit ? 1 : -1;
					}
					// bit s of `bits` belongs to row dhi - s
					if (lane == 0) {
						for (int s = 0; s < steps; ++s)
							if ((bits >> s) & 1u) { const int r = dhi - s; S.path[r >> 5] |= 1u << (r & 31); }
					}
				}
				__syncwarp();
			}

			// ---- replay the path forwards, emitting columns
			if (cols + kcols > slot.cap) { overflow = 1; break; }
			{
				int x = 0, y = 0, emitted = 0;
				char* bq = oq + cols;
				char* bt = ot + cols;
				for (int d = 0; d <= kd && emitted < kcols; ++d) {
					if (d > 0) {
						const bool from_right = (S.path[d >> 5] >> (d & 31)) & 1u;
						if (lane == 0) {
							if (from_right) { bq[emitted] = '-'; bt[emitted] = "ACGT"[base_at(st, y)]; }
							else { bq[emitted] = "ACGT"[base_at(sq, x)]; bt[emitted] = '-'; }
						}
						if (from_right) ++y; else ++x;
						++emitted;
					}
					// snake: 32 bases per ballot; the last cell's snake stops `drop` columns early
					for (;;) {
						const int xi = x + lane, yi = y + lane;
						bool eq = false;
						int c = 0;
						if (xi < qblk && yi < tblk) { c = base_at(sq, xi); eq = c == base_at(st, yi); }
						const unsigned m = __ballot_sync(FULL, eq);
						int run = (m == FULL) ? 32 : (__ffs(~m) - 1);
						run = min(run, kcols - emitted);
						if (lane < run) { const char ch = "ACGT"[c]; bq[emitted + lane] = ch; bt[emitted + lane] = ch; }
						x += run; y += run; emitted += run;
						if (run < 32) break;
					}
				}
				const int m_here = ((kx + ky - kd) >> 1) - drop;
				cols += kcols; mats += m_here; qadv += kx - drop; tadv += ky - drop;
			}
			if (!go_on) break;
			qi += kx - 4; ti += ky - 4;
		}
		if (lane == 0) {
			slot.cols = cols; slot.matches = mats; slot.qadv = qadv; slot.tadv = tadv; slot.overflow = overflow;
			slots[item] = slot;
		}
	}
}

// Per task: sizes after merging (and, for the cns flavour, trimming both ends to a 4-match run).
// out[i] = {ok, qstart, qend, sstart, send, columns, matches, first kept merged column}
template <int POLICY>
__global__ void k_aln_sizes(const AlignTask* __restrict__ tasks, const AlnSlot* __restrict__ slots, size_t ntasks, int min_aln,
                            const char* __restrict__ colq, const char* __restrict__ colt, int32_t* __restrict__ out)
{
	const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (i >= ntasks) return;
	const AlignTask t = tasks[i];
	const AlnSlot L = slots[2 * i], R = slots[2 * i + 1];
	if (POLICY == 2) {
		// XdropAligner::go (xdrop_gapalign.cpp:351-439): the left part without its last column, ok by the query span
		if (lane == 0) {
			const mbx::Half hl = {L.cols, L.matches, L.qadv, L.tadv, L.overflow >> 1, L.overflow & 1};
			const mbx::Half hr = {R.cols, R.matches, R.qadv, R.tadv, R.overflow >> 1, R.overflow & 1};
			mbx::finish(t.qstart, t.sstart, hl, hr, min_aln, out + 8 * i);
		}
		return;
	}
	const int n = L.cols + R.cols;
	int ok = (n >= min_aln) && !L.overflow && !R.overflow;
	int qs = t.qstart - L.qadv, qe = t.qstart + R.qadv, ss = t.sstart - L.tadv, se = t.sstart + R.tadv;
	int first = 0, size = n, mats = L.matches + R.matches;
	if (POLICY == 1 && ok) {
		// merged column c: c < L.cols -> left slot column L.cols-1-c, else right slot column c-L.cols
		auto colpair = [&](int c, char& a, char& b) {
			if (c < L.cols) { a = colq[L.off + (L.cols - 1 - c)]; b = colt[L.off + (L.cols - 1 - c)]; }
			else { a = colq[R.off + (c - L.cols)]; b = colt[R.off + (c - L.cols)]; }
		};
		// first run of 4 matching columns from the left (dw.cpp:499-512)
		int run = 0, start = -1, qrb = 0, trb = 0, mcut = 0;      // mcut: matching columns that the trimming drops
		for (int base = 0; base < n && start < 0; base += 32) {
			const int c = base + lane;
			char a = 0, b = 1;
			if (c < n) colpair(c, a, b);
			const unsigned m = __ballot_sync(0xFFFFFFFFu, c < n && a == b);
			const unsigned qm = __ballot_sync(0xFFFFFFFFu, c < n && a != '-');
			const unsigned tm = __ballot_sync(0xFFFFFFFFu, c < n && b != '-');
			for (int s = 0; s < 32 && base + s < n; ++s) {
				run = ((m >> s) & 1u) ? run + 1 : 0;
				qrb += (qm >> s) & 1u; trb += (tm >> s) & 1u; mcut += (m >> s) & 1u;
				if (run == 4) { start = base + s - 3; break; }
			}
		}
		if (start < 0) ok = 0;
		int qre = 0, tre = 0, endc = -1;
		if (ok) {
			qrb -= 4; trb -= 4; mcut -= 4;
			run = 0;
			for (int base = n - 1; base >= 0 && endc < 0; base -= 32) {
				const int c = base - lane;
				char a = 0, b = 1;
				if (c >= 0) colpair(c, a, b);
				const unsigned m = __ballot_sync(0xFFFFFFFFu, c >= 0 && a == b);
				const unsigned qm = __ballot_sync(0xFFFFFFFFu, c >= 0 && a != '-');
				const unsigned tm = __ballot_sync(0xFFFFFFFFu, c >= 0 && b != '-');
				for (int s = 0; s < 32 && base - s >= 0; ++s) {
					run = ((m >> s) & 1u) ? run + 1 : 0;
					qre += (qm >> s) & 1u; tre += (tm >> s) & 1u; mcut += (m >> s) & 1u;
					if (run == 4) { endc = base - s + 3; break; }
				}
			}
			if (endc < 0) ok = 0;
		}
		if (ok) {
			qre -= 4; tre -= 4; mcut -= 4;
			first = start; size = endc + 1 - start;
			mats -= mcut;                 // matches, columns and strings all describe the trimmed alignment
			qs += qrb; qe -= qre; ss += trb; se -= tre;
		}
	}
	if (lane == 0) {
		int32_t* o = out + 8 * i;
		o[0] = ok; o[1] = qs; o[2] = qe; o[3] = ss; o[4] = se; o[5] = ok ? size : 0; o[6] = mats; o[7] = first;
	}
}

// Packs the merged (reversed left + right) columns of every accepted task at its final offset.
__global__ void k_aln_pack(const AlnSlot* __restrict__ slots, size_t ntasks, const int32_t* __restrict__ info,
                           const unsigned long long* __restrict__ outoff, const char* __restrict__ colq,
                           const char* __restrict__ colt, char* __restrict__ packq, char* __restrict__ packt)
{
	const size_t i = blockIdx.x;
	if (i >= ntasks) return;
	const int32_t* o = info + 8 * i;
	if (!o[0]) return;
	const AlnSlot L = slots[2 * i], R = slots[2 * i + 1];
	const int first = o[7], size = o[5];
	char* dq = packq + outoff[i];
	char* dt = packt + outoff[i];
	for (int c = threadIdx.x; c < size; c += blockDim.x) {
		const int m = first + c;
		if (m < L.cols) { dq[c] = colq[L.off + (L.cols - 1 - m)]; dt[c] = colt[L.off + (L.cols - 1 - m)]; }
		else { dq[c] = colq[R.off + (m - L.cols)]; dt[c] = colt[R.off + (m - L.cols)]; }
	}
	if (threadIdx.x == 0) { dq[size] = 0; dt[size] = 0; }
}

// ------------------------------------------------------------------------------------------ host
size_t align_task_columns(const DVolume* q, const DVolume* s, const AlignTask& t)
{
	const int ql = q->h_offsz[2 * t.qread + 1];
	const int sl = t.swin_len > 0 ? t.swin_len : s->h_offsz[2 * t.sread + 1];
	return ((size_t)t.qstart + t.sstart + 8) + ((size_t)(ql - t.qstart) + (sl - t.sstart) + 8);
}

void align_dev_release(Ctx* c, AlignDev* d)
{
	c->dfree(d->d_info); c->dfree(d->d_packq); c->dfree(d->d_packt); c->dfree(d->d_outoff);
	*d = AlignDev();
}

// One arena batch: the tasks' worst-case columns (align_task_columns) must sum to at most Ctx::align_arena.  The
// results stay in device memory (out); align_batch() below copies them to the host, the consensus stage of
// mecat2cns (cns.cu) consumes them in place.
int align_batch_device(Ctx* c, int policy, double err, const DVolume* q, const DVolume* s, const AlignTask* h_tasks, size_t nb,
                       int min_aln, AlignDev* out, std::vector<int32_t>& info)
{
	*out = AlignDev();
	if (!nb) return 0;
	if (policy == 1 && !(err > 0.0 && err <= 0.2)) MB_FAIL(c, "align_batch: error rate %.3f is outside this path (<= 0.20)", err);
	const bool wide = policy == 1 && err > 0.16;
	const int MAXROWS = wide ? AlDims<true>::MAXROWS : AlDims<false>::MAXROWS;
	int per_sm = wide ? 6 : 7;      // 28 warps per SM: the smem (7.3 KB per warp) and register (68) limit; measured 541 / 470 / 430 / 407 ms at 4 / 5 / 6 / 7
	if (const char* e = getenv("MECAT_B200_ALIGN_CTAS")) per_sm = std::max(1, atoi(e));   // tuning hook
	const int grid = c->sm_count * per_sm;
	const size_t nwarps = (size_t)grid * AL_WARPS;
	uint32_t* d_dir = nullptr;
	short* d_min = nullptr;
	AlignTask* d_tasks = nullptr;
	AlnSlot* d_slots = nullptr;
	char *d_colq = nullptr, *d_colt = nullptr;
	auto body = [&]() -> int {
		std::vector<AlnSlot> slots;
		size_t used = 0;
		for (size_t i = 0; i < nb; ++i) {
			const AlignTask& t = h_tasks[i];
			const int ql = q->h_offsz[2 * t.qread + 1];
			const int sl = t.swin_len > 0 ? t.swin_len : s->h_offsz[2 * t.sread + 1];
			const size_t capL = (size_t)t.qstart + t.sstart + 8, capR = (size_t)(ql - t.qstart) + (sl - t.sstart) + 8;
			AlnSlot a; memset(&a, 0, sizeof a);
			a.off = used; a.cap = (int32_t)capL; used += capL; slots.push_back(a);
			a.off = used; a.cap = (int32_t)capR; used += capR; slots.push_back(a);
		}
		if (used > c->align_arena) MB_FAIL(c, "align_batch: %zu tasks need more than the %zu-byte column arena", nb, c->align_arena);
		if (policy != 2) {
			MB_CUDA(c, c->alloc(&d_dir, nwarps * MAXROWS * DIRW));
			MB_CUDA(c, c->alloc(&d_min, nwarps * MAXROWS));
		}
		size_t arena = 256ull << 20;                       // power-of-two sizes: the pool hands the same block back batch after batch
		while (arena < used + 16) arena <<= 1;
		arena = std::min(arena, std::max(c->align_arena, used + 16));
		MB_CUDA(c, c->dmalloc((void**)&d_colq, arena));
		MB_CUDA(c, c->dmalloc((void**)&d_colt, arena));
		MB_CUDA(c, c->alloc(&d_tasks, nb));
		MB_CUDA(c, c->alloc(&d_slots, 2 * nb));
		MB_CUDA(c, c->alloc(&out->d_info, 8 * nb));
		MB_CUDA(c, c->alloc(&out->d_outoff, nb + 1));
		MB_CUDA(c, cudaMemcpyAsync(d_tasks, h_tasks, sizeof(AlignTask) * nb, cudaMemcpyHostToDevice, c->stream));
		MB_CUDA(c, cudaMemcpyAsync(d_slots, slots.data(), sizeof(AlnSlot) * 2 * nb, cudaMemcpyHostToDevice, c->stream));
		MB_CUDA(c, cudaMemsetAsync(c->d_counters + 4, 0, 8, c->stream));
		if (policy == 2) {
			if (xdrop_fill_slots(c, q, s, d_tasks, nb, d_slots, d_colq, d_colt)) return 1;
		} else {
			KScope ks(c, MECAT_K_EXTEND);
			if (policy == 0)
				k_align<0, false><<<grid, AL_WARPS * 32, 0, c->stream>>>(q->fwd, q->rev, q->offsz, q->num_bases, s->fwd, s->rev, s->offsz, s->num_bases,
				                                                  d_tasks, nb, d_slots, d_colq, d_colt, d_dir, d_min, err, c->d_counters + 4);
			else if (wide)
				k_align<1, true><<<grid, AL_WARPS * 32, 0, c->stream>>>(q->fwd, q->rev, q->offsz, q->num_bases, s->fwd, s->rev, s->offsz, s->num_bases,
				                                                        d_tasks, nb, d_slots, d_colq, d_colt, d_dir, d_min, err, c->d_counters + 4);
			else
				k_align<1, false><<<grid, AL_WARPS * 32, 0, c->stream>>>(q->fwd, q->rev, q->offsz, q->num_bases, s->fwd, s->rev, s->offsz, s->num_bases,
				                                                         d_tasks, nb, d_slots, d_colq, d_colt, d_dir, d_min, err, c->d_counters + 4);
		}
		{
			KScope ks(c, MECAT_K_FINAL);
			const unsigned g2 = (unsigned)((nb * 32 + 127) / 128);
			if (policy == 0) k_aln_sizes<0><<<g2, 128, 0, c->stream>>>(d_tasks, d_slots, nb, min_aln, d_colq, d_colt, out->d_info);
			else if (policy == 2) k_aln_sizes<2><<<g2, 128, 0, c->stream>>>(d_tasks, d_slots, nb, min_aln, d_colq, d_colt, out->d_info);
			else k_aln_sizes<1><<<g2, 128, 0, c->stream>>>(d_tasks, d_slots, nb, min_aln, d_colq, d_colt, out->d_info);
		}
		MB_CUDA(c, cudaGetLastError());
		info.resize(8 * nb);
		MB_CUDA(c, cudaMemcpyAsync(info.data(), out->d_info, sizeof(int32_t) * 8 * nb, cudaMemcpyDeviceToHost, c->stream));
		MB_CUDA(c, cudaStreamSynchronize(c->stream));
		std::vector<unsigned long long> outoff(nb + 1);
		size_t total = 0;
		for (size_t i = 0; i < nb; ++i) { outoff[i] = total; if (info[8 * i]) total += (size_t)info[8 * i + 5] + 1; }
		outoff[nb] = total;
		MB_CUDA(c, c->dmalloc((void**)&out->d_packq, total + 16));
		MB_CUDA(c, c->dmalloc((void**)&out->d_packt, total + 16));
		MB_CUDA(c, cudaMemcpyAsync(out->d_outoff, outoff.data(), sizeof(unsigned long long) * (nb + 1), cudaMemcpyHostToDevice, c->stream));
		{
			KScope ks(c, MECAT_K_FINAL);
			k_aln_pack<<<(unsigned)nb, 128, 0, c->stream>>>(d_slots, nb, out->d_info, out->d_outoff, d_colq, d_colt, out->d_packq, out->d_packt);
		}
		MB_CUDA(c, cudaGetLastError());
		MB_CUDA(c, cudaStreamSynchronize(c->stream));      // outoff (host vector) was the source of an async copy
		out->total = total;
		c->stats.h2d_bytes += (int64_t)((sizeof(AlignTask) + 2 * sizeof(AlnSlot) + 8) * nb);
		c->stats.d2h_bytes += (int64_t)(32 * nb);
		return 0;
	};
	int rc = body();
	c->dfree(d_dir); c->dfree(d_min); c->dfree(d_tasks); c->dfree(d_slots); c->dfree(d_colq); c->dfree(d_colt);
	if (rc) align_dev_release(c, out);
	return rc;
}

int align_batch(Ctx* c, int policy, double err, const DVolume* q, const DVolume* s, const AlignTask* h_tasks, size_t ntasks,
                int min_aln, mecat_align_result* h_results, std::vector<char>& qstr, std::vector<char>& sstr, bool want_strings)
{
	qstr.clear(); sstr.clear();
	if (!ntasks) return 0;
	std::vector<int32_t> info;
	size_t done = 0;
	while (done < ntasks) {
		size_t used = 0, nb = 0;
		while (done + nb < ntasks) {
			const size_t need = align_task_columns(q, s, h_tasks[done + nb]);
			if (used + need > c->align_arena) break;
			used += need; ++nb;
		}
		if (nb == 0) MB_FAIL(c, "align_batch: one task needs more than the %zu-byte column arena", c->align_arena);
		AlignDev dev;
		if (align_batch_device(c, policy, err, q, s, h_tasks + done, nb, min_aln, &dev, info)) return 1;
		const size_t base = qstr.size(), total = want_strings ? dev.total : 0;
		qstr.resize(base + total); sstr.resize(base + total);
		cudaError_t e = cudaSuccess;
		if (total) {
			e = cudaMemcpyAsync(qstr.data() + base, dev.d_packq, total, cudaMemcpyDeviceToHost, c->stream);
			if (e == cudaSuccess) e = cudaMemcpyAsync(sstr.data() + base, dev.d_packt, total, cudaMemcpyDeviceToHost, c->stream);
		}
		if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
		align_dev_release(c, &dev);
		if (e != cudaSuccess) MB_FAIL(c, "align_batch: D2H: %s", cudaGetErrorString(e));
		c->resolve_timers();
		c->stats.d2h_bytes += (int64_t)(2 * total);
		size_t at = 0;
		for (size_t i = 0; i < nb; ++i) {
			mecat_align_result& r = h_results[done + i];
			const int32_t* o = &info[8 * i];
			r.ok = o[0]; r.qstart = o[1]; r.qend = o[2]; r.sstart = o[3]; r.send = o[4];
			r.columns = o[5]; r.matches = o[6]; r.pad_ = 0;
			r.str_offset = (o[0] && want_strings) ? (int64_t)(base + at) : -1;
			if (o[0]) at += (size_t)o[5] + 1;
			r.ident = (o[0] && o[5]) ? 100.0 * o[6] / o[5] : 0.0;
		}
		done += nb;
	}
	return 0;
}

}  // namespace mb
