// mecat_b200/csrc/extend.cu -- batched O(nd) gapped extension (pw / ref flavour).
//
// Replaces, for a whole batch of candidates at once, the reference's per-candidate
//   DiffAligner::go            src/common/diff_gapalign.cpp:295-349
//   dw_in_one_direction        src/common/diff_gapalign.cpp:221-292
//   retrieve_next_aln_block    src/common/gapalign.cpp:10-45
//   Align                      src/common/diff_gapalign.cpp:107-219
//   GetAlignString + trim_mismatch_end   diff_gapalign.cpp:40-104, gapalign.cpp:48-67
//
// Mapping: one warp owns one (candidate, direction) chain of <= 720 x 720 blocks.  The two
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

__device__ __forceinline__ uint32_t seq16(const uint2* s, int i)
{
	const uint2 w = s[i >> 4];
	return __funnelshift_r(w.x, w.y, (i & 15) << 1);
}

__device__ __forceinline__ int dtrunc_mul(double a, int b) { return (int)__dmul_rn(a, (double)b); }

}  // namespace

// Persistent warps: every warp pulls (candidate, direction) items from a global counter, so a
// long chain never pins three idle warps of its CTA.
__global__ void __launch_bounds__(EXT_WARPS * 32, EXT_CTAS_PER_SM)
k_extend(const uint32_t* __restrict__ qfwd, const uint32_t* __restrict__ qrev, const int2* __restrict__ qoffsz, int qN,
         const uint32_t* __restrict__ sfwd, const uint32_t* __restrict__ srev, const int2* __restrict__ soffsz, int sN,
         const ExtendTask* __restrict__ tasks, size_t ntasks, ExtendHalf* __restrict__ halves,
         unsigned long long* __restrict__ block_counter, unsigned long long* __restrict__ work_counter)
{
	__shared__ WarpSmem smem[EXT_WARPS];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	WarpSmem& S = smem[warp];
	unsigned nblocks = 0;

	for (;;) {
		unsigned long long item = 0;
		if (lane == 0) item = atomicAdd(work_counter, 1ull);
		item = __shfl_sync(FULL, item, 0);
		if (item >= 2 * ntasks) break;
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

		int qi = 0, ti = 0;
		int cols = 0, mats = 0, qadv = 0, tadv = 0;

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
					S.sq[i] = make_uint2(ld_bases32(Q.arr, b0) ^ Q.comp, ld_bases32(Q.arr, b0 + 16u) ^ Q.comp);
				}
				for (int i = lane; i < tw; i += 32) {
					const uint32_t b0 = T.g0 + (uint32_t)ti + 16u * i;
					S.st[i] = make_uint2(ld_bases32(T.arr, b0) ^ T.comp, ld_bases32(T.arr, b0 + 16u) ^ T.comp);
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
			auto cell = [&](uint2* own, int j, int n, int k, uint32_t dbits, int& x, uint32_t& anc) {
				const uint2 lf = own[2 * j - 1], rt = own[2 * j + 1];
				if (j == 0 || (j != n - 1 && (int)lf.x < (int)rt.x)) { x = (int)rt.x; anc = rt.y; }
				else { x = (int)lf.x + 1; anc = lf.y; }
				int y = x - k;
				const int x1 = x;
				// x <= qblk and y <= tblk here (a cell that reached an end finished the block); the staged
				// words cover one window past either end, and whatever matches there is clamped away below
				for (;;) {
					const uint32_t diff = seq16(sq, x) ^ seq16(st, y);
					const int m = __clz(__brev(diff)) >> 1;      // matching bases in this 16-base window
					x += m; y += m;
					if (m < 16 || x >= qblk || y >= tblk) break;
				}
				const int over = max(max(x - qblk, y - tblk), 0);    // the window may run past a block end
				x -= over; y -= over;
				if (x - x1 >= 4) anc = (uint32_t)x | ((uint32_t)y << 10) | dbits;
				own[2 * j] = make_uint2((uint32_t)x, anc);
				return x + y;
			};
			for (int d = 0; d < max_d; ++d) {
				if (max_k - min_k > 2 * tol) break;
				const int n = ((max_k - min_k) >> 1) + 1;
				uint2* own = &S.vl[min_k + KOFF];                      // own[2j] <-> diagonal min_k + 2j; neighbours own[2j -+ 1]
				const uint32_t dbits = (uint32_t)d << 20;
				// pass 0: diagonals 0..31 of the band (most rows have no other pass)
				int x0 = 0, u0 = IDLE;
				uint32_t a0 = NO_ANCHOR;
				if (lane < n) u0 = cell(own, lane, n, min_k + 2 * lane, dbits, x0, a0);
				int rowmax = __reduce_max_sync(FULL, u0);
				int x1 = 0, u1 = IDLE;
				uint32_t a1 = NO_ANCHOR;
				if (n > 32) {
					if (lane + 32 < n) u1 = cell(own, lane + 32, n, min_k + 2 * (lane + 32), dbits, x1, a1);
					rowmax = max(rowmax, __reduce_max_sync(FULL, u1));
					for (int base = 64; base < n; base += 32) {
						const int j = base + lane;
						int xx = 0, uu = IDLE;
						uint32_t aa = NO_ANCHOR;
						if (j < n) uu = cell(own, j, n, min_k + 2 * j, dbits, xx, aa);
						rowmax = max(rowmax, __reduce_max_sync(FULL, uu));
					}
				}
				__syncwarp();
				// x >= qblk needs x + y >= 2 qblk - k, y >= tblk needs x + y >= 2 tblk + k
				if (rowmax >= min(2 * qblk - max_k, 2 * tblk + min_k)) {
					// some cell may have reached a block end: the lowest such diagonal ends the block
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
	if (lane == 0 && block_counter && nblocks) atomicAdd(block_counter, (unsigned long long)nblocks);
}

int extend_launch(Ctx* c, const DVolume* q, const DVolume* s, const ExtendTask* d_tasks, size_t ntasks,
                  ExtendHalf* d_halves)
{
	if (!ntasks) return 0;
	unsigned long long* d_counter = c->d_counters;          // [0] block statistics, [4] work queue head
	MB_CUDA(c, cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), c->stream));
	MB_CUDA(c, cudaMemsetAsync(d_counter + 4, 0, sizeof(unsigned long long), c->stream));
	const size_t items = 2 * ntasks;
	const size_t want = (items + EXT_WARPS - 1) / EXT_WARPS;
	const unsigned grid = (unsigned)std::min<size_t>(want, (size_t)c->sm_count * EXT_CTAS_PER_SM);
	{
		KScope ks(c, MECAT_K_EXTEND);
		k_extend<<<grid, EXT_WARPS * 32, 0, c->stream>>>(q->fwd, q->rev, q->offsz, q->num_bases, s->fwd, s->rev, s->offsz,
		                                                 s->num_bases, d_tasks, ntasks, d_halves, d_counter, d_counter + 4);
	}
	MB_CUDA(c, cudaGetLastError());
	unsigned long long nb = 0;
	MB_CUDA(c, cudaMemcpyAsync(&nb, d_counter, sizeof nb, cudaMemcpyDeviceToHost, c->stream));
	MB_CUDA(c, cudaStreamSynchronize(c->stream));
	c->resolve_timers();
	c->stats.num_extend_blocks += (int64_t)nb;
	return 0;
}

}  // namespace mb
